mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_nhwc_kernel -s 1 -c 1 -o gpurun_out/r2u_prof_wgrad_nhwc -f python tools/tc_one.py --shape 4,128,64,208,128,3,1,1 --bwd --iters 1 > gpurun_out/r2u_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r2u_prof_wgrad_nhwc.ncu-rep --page raw --csv > gpurun_out/r2u_prof_wgrad_nhwc.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r2u_prof_wgrad_nhwc.csv | grep -v "^---"
ncu -i gpurun_out/r2u_prof_wgrad_nhwc.ncu-rep --page details 2>/dev/null | grep -E "Stall|Eligible|No Eligible|Issued Warp|L1/TEX Hit|Shared|shared|Bank|bank" | head -20
