mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "tensor_core or net_case" > gpurun_out/pytest_gpu9.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu9.log
for shp in 4,32,128,416,32,7,1,3 4,16,256,832,1,3,1,1 4,128,32,104,128,3,1,1 4,16,256,832,16,3,1,1; do
  timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/l9_$shp.csv python tools/tc_one.py --shape $shp --impl 0 --bwd --iters 3 > gpurun_out/l9_$shp.txt 2>&1
done
timeout 300 python tools/bench_conv.py --impl 0 > gpurun_out/bench_conv_auto5.jsonl 2> gpurun_out/bench_conv_auto5.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench9.json 2> gpurun_out/bench9.err
grep -n "rel err\|passed\|failed\|Error" gpurun_out/pytest_gpu9.log | head; tail -1 gpurun_out/bench_conv_auto5.jsonl; cut -c1-300 gpurun_out/bench9.json
