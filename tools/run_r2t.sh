mkdir -p gpurun_out
timeout 900 python tools/nhwc_probe.py 1 2 3 6 7 9 10 11 13 > gpurun_out/r2t_nhwc_probe.jsonl 2> gpurun_out/r2t_nhwc_probe.err; echo "probe rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/r2t_nhwc_probe.jsonl'):
    d=json.loads(l)
    if 'time_auto' not in d: print(str(d)[:600]); continue
    print(d['case'], 'err f %.1e d %.1e w %.1e'%(d['dbg0']['fprop'], d['dbg0']['dgrad'], d['dbg0']['wgrad']), d['dbg0']['status'], ' | '.join('%s f%.0f d%.0f w%.0f'%(k[5:], v['fprop_us'], v['dgrad_us'], v['wgrad_us']) for k,v in d.items() if k.startswith('time_')))
PY
tail -n 5 gpurun_out/r2t_nhwc_probe.err
