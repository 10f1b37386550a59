mkdir -p gpurun_out
export NCCL_DEBUG=WARN CCB_BENCH_VERBOSE=1 CCB_BENCH_WATCHDOG=70
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_mg2.json 2> gpurun_out/bench_mg2.err
rc=$?; echo "rc=$rc"; cut -c1-700 gpurun_out/bench_mg2.json; grep -n "bench rank\|File \"/.*repo\|Thread\|Current thread" gpurun_out/bench_mg2.err | head -60 | cut -c1-200
if [ $rc -ne 0 ]; then
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-graph > gpurun_out/bench_mg2_nograph.json 2> gpurun_out/bench_mg2_nograph.err
  echo "nograph rc=$?"; cut -c1-700 gpurun_out/bench_mg2_nograph.json; grep -n "bench rank\|File \"/.*repo" gpurun_out/bench_mg2_nograph.err | head -40 | cut -c1-200
fi
