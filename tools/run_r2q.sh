mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tma_family or conv_shapes or big_shapes or net_case" > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r2q_pytest.log
timeout 600 python tools/conv_calls.py cfg3 12 > gpurun_out/r2q_conv_calls_cfg3.txt 2> gpurun_out/r2q_conv_calls.err; echo "calls rc=$?"; head -16 gpurun_out/r2q_conv_calls_cfg3.txt; tail -n 3 gpurun_out/r2q_conv_calls.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-profile --no-cpu-baseline --no-side-configs --no-reference-gpu > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; echo "bench rc=$?"; cut -c1-250 gpurun_out/r2q_bench.json
