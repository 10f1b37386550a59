mkdir -p gpurun_out
CCB_NHWC_MINH=8 CCB_NHWC_WASTE10=30 timeout 600 python tools/nhwc_probe.py 23 24 25 26 > gpurun_out/r2z_probe.jsonl 2> gpurun_out/r2z_probe.err; echo "probe rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/r2z_probe.jsonl'):
    d=json.loads(l)
    if 'time_auto' not in d: print(str(d)[:500]); continue
    print(d['case'], 'err f %.1e d %.1e w %.1e'%(d['dbg0']['fprop'], d['dbg0']['dgrad'], d['dbg0']['wgrad']), d['dbg0']['status'], ' | '.join('%s f%.0f bwd%.0f'%(k[5:], v['fprop_us'], v['dgrad_us']) for k,v in d.items() if k.startswith('time_')))
PY
tail -n 3 gpurun_out/r2z_probe.err
