mkdir -p gpurun_out
timeout 900 python tools/nhwc_probe.py 6 7 8 9 10 11 12 13 14 15 16 > gpurun_out/r2f_nhwc_probe.jsonl 2> gpurun_out/r2f_nhwc_probe.err; echo "probe rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/r2f_nhwc_probe.jsonl'):
    d=json.loads(l)
    if 'time_auto' not in d: print(d); continue
    print(d['case'], 'err', '%.1e'%d['dbg0']['fprop'], '%.1e'%d['dbg0']['dgrad'], ' | '.join('%s f%.0f d%.0f'%(k[5:], v['fprop_us'], v['dgrad_us']) for k,v in d.items() if k.startswith('time_')))
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_nhwc_kernel -s 2 -c 1 -o gpurun_out/r2f_prof_nhwc_128 -f python tools/tc_one.py --shape 4,128,64,208,128,3,1,1 --iters 1 > gpurun_out/r2f_ncu_nhwc_128.log 2>&1; echo "ncu1 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_nhwc_kernel -s 2 -c 1 -o gpurun_out/r2f_prof_nhwc_7x7 -f python tools/tc_one.py --shape 4,32,128,416,32,7,1,3 --iters 1 > gpurun_out/r2f_ncu_nhwc_7x7.log 2>&1; echo "ncu2 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_slab_wgrad_kernel -s 1 -c 1 -o gpurun_out/r2f_prof_wgrad_128 -f python tools/tc_one.py --shape 4,128,64,208,128,3,1,1 --bwd --iters 1 > gpurun_out/r2f_ncu_wgrad_128.log 2>&1; echo "ncu3 rc=$?"
ls -la gpurun_out/*.ncu-rep
