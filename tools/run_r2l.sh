mkdir -p gpurun_out
timeout 900 python tools/nhwc_probe.py 17 18 19 20 21 22 9 > gpurun_out/r2l_nhwc_probe.jsonl 2> gpurun_out/r2l_nhwc_probe.err; echo "probe rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/r2l_nhwc_probe.jsonl'):
    d=json.loads(l)
    if 'time_auto' not in d: print(d); continue
    print(d['case'], 'err', '%.1e'%d['dbg0']['fprop'], '%.1e'%d['dbg0']['dgrad'], 'thin err', '%.1e'%d['dbg1']['fprop'], '%.1e'%d['dbg1']['dgrad'], d['dbg1']['status'], ' | '.join('%s f%.0f d%.0f'%(k[5:], v['fprop_us'], v['dgrad_us']) for k,v in d.items() if k.startswith('time_')))
PY
timeout 600 python tools/conv_calls.py cfg3 25 > gpurun_out/r2l_conv_calls_cfg3.txt 2> gpurun_out/r2l_conv_calls.err; echo "calls rc=$?"; head -18 gpurun_out/r2l_conv_calls_cfg3.txt; tail -n 3 gpurun_out/r2l_conv_calls.err
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "nhwc or tma_family or tensor_core or cfg3 or net_case" > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/r2l_pytest.log
