mkdir -p gpurun_out
export CCB_BENCH_WATCHDOG=300
N=$(nvidia-smi -L | wc -l); echo "gpus=$N"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 3 --no-profile --no-cpu-baseline > gpurun_out/r2o_bench_${N}gpu.json 2> gpurun_out/r2o_bench_${N}gpu.err
rc=$?; echo "rc=$rc"; python -c "
import json;d=json.load(open('gpurun_out/r2o_bench_${N}gpu.json'));print('N=$N', d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['value'], 'launches/step', d['gpu_launches_per_step'], d['clocks'])"; tail -n 3 gpurun_out/r2o_bench_${N}gpu.err | cut -c1-300
