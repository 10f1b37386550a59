mkdir -p gpurun_out
CCB_PROBE_CASES=48,49,50,51,52,53,54,55,12,32 timeout 600 python tools/tma_probe.py > gpurun_out/tma_probe7.jsonl 2> gpurun_out/tma_probe7.err
python - <<'PY'
import json
for l in open('gpurun_out/tma_probe7.jsonl'):
    d=json.loads(l)
    if 'tma3' not in d: print('FAIL', l[:300]); continue
    c=d['case']
    flag = 'BAD' if d['tma3']>2e-4 or d['tma1']>3e-3 else 'ok'
    print(flag, c, 'err3 %.1e err1 %.1e'%(d['tma3'],d['tma1']), 'ms: gather3 %.3f tma3 %.3f tma1 %.3f gather1 %.3f'%(d['gather3_ms'],d['tma3_ms'],d['tma1_ms'],d['gather1_ms']))
PY
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu22.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu22.log
tail -n 4 gpurun_out/pytest_gpu22.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench22.json 2> gpurun_out/bench22.err
cut -c1-300 gpurun_out/bench22.json; tail -n 3 gpurun_out/bench22.err
