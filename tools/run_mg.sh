mkdir -p gpurun_out
export CCB_BENCH_WATCHDOG=100
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_mg2.json 2> gpurun_out/bench_mg2.err
rc=$?; echo "rc=$rc"; cut -c1-1200 gpurun_out/bench_mg2.json; tail -n 3 gpurun_out/bench_mg2.err | cut -c1-200
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_mg2_ref.json 2> gpurun_out/bench_mg2_ref.err
echo "ref rc=$?"; cut -c1-200 gpurun_out/bench_mg2_ref.json
