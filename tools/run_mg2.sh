mkdir -p gpurun_out
export CCB_BENCH_WATCHDOG=300
nvidia-smi topo -m | head -8
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2m_bench_2gpu.json 2> gpurun_out/r2m_bench_2gpu.err
rc=$?; echo "rc=$rc"; cut -c1-1500 gpurun_out/r2m_bench_2gpu.json; tail -n 5 gpurun_out/r2m_bench_2gpu.err | cut -c1-300
