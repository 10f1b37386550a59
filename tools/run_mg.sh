mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/mg_smi.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_mg2.json 2> gpurun_out/bench_mg2.err
echo "rc=$?"; cut -c1-600 gpurun_out/bench_mg2.json; tail -n 5 gpurun_out/bench_mg2.err
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mg1.json 2> gpurun_out/bench_mg1.err
cut -c1-300 gpurun_out/bench_mg1.json
