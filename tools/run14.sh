mkdir -p gpurun_out
timeout 1200 python tools/tma_probe.py > gpurun_out/tma_probe2.jsonl 2> gpurun_out/tma_probe2.err
cut -c1-330 gpurun_out/tma_probe2.jsonl
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu14.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu14.log
tail -5 gpurun_out/pytest_gpu14.log
timeout 300 python tools/bench_conv.py --impl 0 > gpurun_out/bench_conv_slab.jsonl 2> gpurun_out/bench_conv_slab.err
tail -1 gpurun_out/bench_conv_slab.jsonl
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench14.json 2> gpurun_out/bench14.err
cut -c1-300 gpurun_out/bench14.json
