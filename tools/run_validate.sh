mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2x_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r2x_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2x_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/r2x_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2x_bench_cfg3.json 2> gpurun_out/r2x_bench_cfg3.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2x_bench_cfg3.json; tail -n 2 gpurun_out/r2x_bench_cfg3.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2x_launches_cfg3.csv python tools/launch_list.py --cfg cfg3 > gpurun_out/r2x_launches_cfg3.txt 2>&1; echo "ncu rc=$?"
python tools/launch_list.py --summarise gpurun_out/r2x_launches_cfg3.csv > gpurun_out/r2x_launches_cfg3_summary.txt; head -14 gpurun_out/r2x_launches_cfg3_summary.txt
