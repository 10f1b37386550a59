mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2fin_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2fin_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2fin_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/r2fin_smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-side-configs --no-reference-gpu > gpurun_out/r2fin_bench.json 2> gpurun_out/r2fin_bench.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/r2fin_bench.json
