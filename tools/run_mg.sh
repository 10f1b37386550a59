mkdir -p gpurun_out
export CCB_BENCH_WATCHDOG=100
for feed in inline prefetch; do
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --feed $feed --no-profile > gpurun_out/bench_mg2_$feed.json 2> gpurun_out/bench_mg2_$feed.err
echo "$feed rc=$?"; python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_mg2_$feed.json').read().strip().splitlines()[-1]); print('$feed', d['value'], d['ms_per_step'], d['e2e'])"
done
nvidia-smi topo -m > gpurun_out/mg_topo.txt 2>&1; head -8 gpurun_out/mg_topo.txt
