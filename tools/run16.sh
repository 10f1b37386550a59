mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_step2.csv python bench.py --steps 1 --warmup 1 --no-graph --no-profile --no-cpu-baseline > gpurun_out/launches_step2.txt 2>&1
wc -l gpurun_out/launches_step2.csv; tail -n 2 gpurun_out/launches_step2.txt | cut -c1-200
