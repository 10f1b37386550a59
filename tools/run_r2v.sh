mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "nhwc or tma_family or flow_golden or featwarp or warp_golden or cfg3" > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r2v_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-profile --no-cpu-baseline --no-side-configs --no-reference-gpu > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo "bench rc=$?"; cut -c1-250 gpurun_out/r2v_bench.json
