mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_fullsize.py > gpurun_out/r2j_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/r2j_pytest_gpu.log
timeout 1700 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -s > gpurun_out/r2j_fullsize.log 2>&1; echo "fullsize rc=$?"; grep -E "== parity|grads |FAIL|passed|failed|Error|miss" gpurun_out/r2j_fullsize.log | head -60
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2j_bench_cfg3.json 2> gpurun_out/r2j_bench_cfg3.err; echo "bench rc=$?"; cat gpurun_out/r2j_bench_cfg3.json; tail -n 3 gpurun_out/r2j_bench_cfg3.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2j_launches_cfg3.csv python tools/launch_list.py --cfg cfg3 > gpurun_out/r2j_launches_cfg3.txt 2>&1; echo "ncu rc=$?"
python tools/launch_list.py --summarise gpurun_out/r2j_launches_cfg3.csv > gpurun_out/r2j_launches_cfg3_summary.txt; head -30 gpurun_out/r2j_launches_cfg3_summary.txt
