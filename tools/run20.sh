mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu20.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu20.log
tail -n 4 gpurun_out/pytest_gpu20.log
timeout 300 python tools/bench_conv.py --impl 0 > gpurun_out/bench_conv_20.jsonl 2> gpurun_out/bench_conv_20.err
tail -n 1 gpurun_out/bench_conv_20.jsonl; tail -n 3 gpurun_out/bench_conv_20.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench20.json 2> gpurun_out/bench20.err
cut -c1-300 gpurun_out/bench20.json; tail -n 3 gpurun_out/bench20.err
