mkdir -p gpurun_out
export CCB_BENCH_WATCHDOG=300
run() { tag=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-profile --no-cpu-baseline > gpurun_out/r2n_$tag.json 2> gpurun_out/r2n_$tag.err; echo "$tag rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/r2n_$tag.json'));print('$tag', d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], 'launches/step', d['gpu_launches_per_step'])"; }
run default CCB_X=1
run ctas8 NCCL_MAX_CTAS=8
run ctas16 NCCL_MAX_CTAS=16
run ctas4 NCCL_MAX_CTAS=4
run bucket64 CCB_BUCKET_MB=64
run bucket300 CCB_BUCKET_MB=400
