mkdir -p gpurun_out
python - <<'PY' > gpurun_out/tma_dbg.txt 2>&1
import subprocess, sys, os, json
sys.path.insert(0,'.')
import tools.tma_probe as tp
tp.CASES[:] = [(1, 8, 6, 32, 16, 1, 1, 0, 'fprop'), (1, 8, 4, 32, 16, 3, 1, 1, 'fprop')]
PY
for dbg in 0 1 2 4 3 6 7; do
 for i in 0 1; do
  echo "dbg=$dbg case=$i" >> gpurun_out/tma_dbg.txt
  CCB_TMA_DBG=$dbg CCB_PROBE_ALT=1 timeout 100 python tools/tma_probe.py case $i 2>&1 | tail -2 | cut -c1-400 >> gpurun_out/tma_dbg.txt
 done
done
cat gpurun_out/tma_dbg.txt
