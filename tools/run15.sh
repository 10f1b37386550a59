mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_slab_kernel -s 2 -c 1 -o gpurun_out/prof_slab_f -f python tools/tc_one.py --shape 4,32,128,416,32,7,1,3 --iters 1 > gpurun_out/ncu_slab_f.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_slab_wgrad -s 1 -c 1 -o gpurun_out/prof_slab_w -f python tools/tc_one.py --shape 4,32,128,416,32,7,1,3 --iters 1 --bwd > gpurun_out/ncu_slab_w.log 2>&1
tail -n 3 gpurun_out/ncu_slab_f.log; tail -n 3 gpurun_out/ncu_slab_w.log
ls -la gpurun_out/*.ncu-rep
