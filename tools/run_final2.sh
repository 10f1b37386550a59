mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_final.log
tail -n 4 gpurun_out/pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_final.log 2>&1; tail -n 2 gpurun_out/smoke_final.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
cut -c1-300 gpurun_out/bench_final.json; tail -n 2 gpurun_out/bench_final.err
timeout 300 python tools/bench_conv.py --impl 0 > gpurun_out/bench_conv_final.jsonl 2> gpurun_out/bench_conv_final.err
tail -n 1 gpurun_out/bench_conv_final.jsonl
