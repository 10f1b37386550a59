mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_fullsize.py > gpurun_out/r2e_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/r2e_pytest_gpu.log
timeout 1700 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -s > gpurun_out/r2e_fullsize.log 2>&1; echo "fullsize rc=$?"; grep -E "== parity|FAIL|passed|failed|Error|rel err" gpurun_out/r2e_fullsize.log | head -80
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2e_bench_cfg3.json 2> gpurun_out/r2e_bench_cfg3.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/r2e_bench_cfg3.json; tail -n 3 gpurun_out/r2e_bench_cfg3.err
