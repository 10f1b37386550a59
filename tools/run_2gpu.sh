mkdir -p gpurun_out
export CCB_BENCH_WATCHDOG=300
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-profile --no-cpu-baseline > gpurun_out/r2y_bench_2gpu.json 2> gpurun_out/r2y_bench_2gpu.err; echo "rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/r2y_bench_2gpu.json'));print('N=2', d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['value'], 'launches/step', d['gpu_launches_per_step'])"; tail -n 2 gpurun_out/r2y_bench_2gpu.err | cut -c1-200
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r2y_ref_2gpu.json 2> gpurun_out/r2y_ref_2gpu.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/r2y_ref_2gpu.json
