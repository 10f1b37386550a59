mkdir -p gpurun_out
rm -f gpurun_out/tma_dbg3.txt
for dbg in 8 10 0; do
  echo "dbg=$dbg" >> gpurun_out/tma_dbg3.txt
  CCB_TMA_DBG=$dbg CCB_PROBE_ALT=1 timeout 100 python tools/tma_probe.py case 1 2>&1 | tail -2 | cut -c1-400 >> gpurun_out/tma_dbg3.txt
done
cat gpurun_out/tma_dbg3.txt
