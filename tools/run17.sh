mkdir -p gpurun_out
timeout 1200 python tools/tma_probe.py > gpurun_out/tma_probe3.jsonl 2> gpurun_out/tma_probe3.err
python - <<'PY'
import json
for l in open('gpurun_out/tma_probe3.jsonl'):
    d=json.loads(l)
    if 'tma3' not in d: print('FAIL', l[:300]); continue
    c=d['case']
    flag = 'BAD' if d['tma3']>2e-4 or d['tma1']>3e-3 else 'ok'
    print(flag, c, 'err3 %.1e err1 %.1e'%(d['tma3'],d['tma1']), 'ms: gather3 %.3f tma3 %.3f tma1 %.3f gather1 %.3f'%(d['gather3_ms'],d['tma3_ms'],d['tma1_ms'],d['gather1_ms']))
PY
CCB_TMA_DBG=32 timeout 200 python tools/tc_one.py --shape 4,32,128,416,32,7,1,3 --iters 1 > gpurun_out/trace_7x7.txt 2>&1
CCB_TMA_DBG=32 timeout 200 python tools/tc_one.py --shape 4,16,256,832,16,3,1,1 --iters 1 > gpurun_out/trace_thin.txt 2>&1
head -52 gpurun_out/trace_7x7.txt | cut -c1-150; head -7 gpurun_out/trace_thin.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu17.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu17.log
tail -n 4 gpurun_out/pytest_gpu17.log
timeout 300 python tools/bench_conv.py --impl 0 > gpurun_out/bench_conv_slab2.jsonl 2> gpurun_out/bench_conv_slab2.err
tail -n 1 gpurun_out/bench_conv_slab2.jsonl
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench17.json 2> gpurun_out/bench17.err
cut -c1-300 gpurun_out/bench17.json
