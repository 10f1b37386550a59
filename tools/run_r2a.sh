mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench_cfg3.json 2> gpurun_out/r2a_bench_cfg3.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/r2a_bench_cfg3.json; tail -n 3 gpurun_out/r2a_bench_cfg3.err
timeout 500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2a_launches_cfg3.csv python tools/launch_list.py --cfg cfg3 > gpurun_out/r2a_launches_cfg3.txt 2>&1; echo "ncu rc=$?"
python tools/launch_list.py --summarise gpurun_out/r2a_launches_cfg3.csv > gpurun_out/r2a_launches_cfg3_summary.txt; head -45 gpurun_out/r2a_launches_cfg3_summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/r2a_smoke.log
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -s > gpurun_out/r2a_fullsize.log 2>&1; echo "fullsize rc=$?"; grep -E "== parity|FAIL|passed|failed|Error" gpurun_out/r2a_fullsize.log | head -60
