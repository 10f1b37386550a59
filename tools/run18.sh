mkdir -p gpurun_out
CCB_PROBE_CASES=9,10,11 timeout 300 python tools/tma_probe.py > gpurun_out/tma_probe4.jsonl 2> gpurun_out/tma_probe4.err
cut -c1-300 gpurun_out/tma_probe4.jsonl
timeout 300 python tools/bench_conv.py --impl 0 > gpurun_out/bench_conv_slab2.jsonl 2> gpurun_out/bench_conv_slab2.err
tail -n 1 gpurun_out/bench_conv_slab2.jsonl; tail -n 3 gpurun_out/bench_conv_slab2.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench17.json 2> gpurun_out/bench17.err
cut -c1-300 gpurun_out/bench17.json; tail -n 3 gpurun_out/bench17.err
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu18.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu18.log
tail -n 4 gpurun_out/pytest_gpu18.log
