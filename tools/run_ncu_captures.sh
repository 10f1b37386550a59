mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:photo_fwd_kernel -c 1 -o gpurun_out/r2r_prof_photo_fwd -f python tools/bench_loss.py --iters 2 > gpurun_out/r2r_ncu_photo_fwd.log 2>&1; echo "ncu1 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:photo_bwd_kernel -c 1 -o gpurun_out/r2r_prof_photo_bwd -f python tools/bench_loss.py --iters 2 > gpurun_out/r2r_ncu_photo_bwd.log 2>&1; echo "ncu2 rc=$?"
timeout 300 ncu --set full --clock-control none -k regex:"corr81_fwd_kernel|corr81_dgrad_kernel|flow_warp_bwd_kernel|conv_tc_kernel|conv_direct_kernel" --launch-skip 300 -c 12 -o gpurun_out/r2r_prof_b2f -f python tools/launch_list.py --cfg cfg2 > gpurun_out/r2r_ncu_b2f.log 2>&1; echo "ncu3 rc=$?"
timeout 200 python tools/bench_loss.py --iters 10 > gpurun_out/r2r_bench_loss.json 2>&1; tail -n 3 gpurun_out/r2r_bench_loss.json | cut -c1-600
ls -la gpurun_out/r2r*.ncu-rep
