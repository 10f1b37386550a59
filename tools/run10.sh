mkdir -p gpurun_out
timeout 900 python tools/tma_probe.py > gpurun_out/tma_probe.jsonl 2> gpurun_out/tma_probe.err
cat gpurun_out/tma_probe.jsonl | cut -c1-400
