mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/final_smi.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_final.log
tail -n 5 gpurun_out/pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_final.log 2>&1; tail -n 3 gpurun_out/smoke_final.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
cut -c1-400 gpurun_out/bench_final.json; tail -n 2 gpurun_out/bench_final.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err
cut -c1-300 gpurun_out/bench_ref_final.json
# launch list of one eager step of the same command (serialised, cold cache: compare shares)
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 1150 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 1 --no-graph --no-profile --no-cpu-baseline > gpurun_out/launches_final.txt 2>&1
wc -l gpurun_out/launches_final.csv
# full-set capture of the dominant kernels
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_slab_kernel -s 2 -c 1 -o gpurun_out/prof_final_slab -f python tools/tc_one.py --shape 4,32,128,416,32,7,1,3 --bwd --iters 1 > gpurun_out/ncu_final_slab.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_direct_kernel -s 1 -c 1 -o gpurun_out/prof_final_direct -f python tools/tc_one.py --shape 4,16,256,832,16,3,1,1 --impl 0 --iters 1 > gpurun_out/ncu_final_direct.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:photo_fwd_kernel -c 1 -o gpurun_out/prof_final_photo -f python tools/bench_loss.py > gpurun_out/ncu_final_photo.log 2>&1
ls -la gpurun_out/prof_final_*.ncu-rep
