mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py > gpurun_out/r2b_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 15 gpurun_out/r2b_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench_cfg3.json 2> gpurun_out/r2b_bench_cfg3.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/r2b_bench_cfg3.json; tail -n 5 gpurun_out/r2b_bench_cfg3.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_launches_cfg3.csv python tools/launch_list.py --cfg cfg3 > gpurun_out/r2b_launches_cfg3.txt 2>&1; echo "ncu rc=$?"
python tools/launch_list.py --summarise gpurun_out/r2b_launches_cfg3.csv > gpurun_out/r2b_launches_cfg3_summary.txt; head -60 gpurun_out/r2b_launches_cfg3_summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/r2b_smoke.log
