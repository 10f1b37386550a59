mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py > gpurun_out/r2c_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 8 gpurun_out/r2c_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c_bench_cfg3.json 2> gpurun_out/r2c_bench_cfg3.err; echo "bench rc=$?"; cut -c1-2500 gpurun_out/r2c_bench_cfg3.json; tail -n 5 gpurun_out/r2c_bench_cfg3.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c_launches_cfg3.csv python tools/launch_list.py --cfg cfg3 > gpurun_out/r2c_launches_cfg3.txt 2>&1; echo "ncu rc=$?"
python tools/launch_list.py --summarise gpurun_out/r2c_launches_cfg3.csv > gpurun_out/r2c_launches_cfg3_summary.txt; head -40 gpurun_out/r2c_launches_cfg3_summary.txt
