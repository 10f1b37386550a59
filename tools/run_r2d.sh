mkdir -p gpurun_out
timeout 1200 python tools/nhwc_probe.py > gpurun_out/r2d_nhwc_probe.jsonl 2> gpurun_out/r2d_nhwc_probe.err; echo "probe rc=$?"; cat gpurun_out/r2d_nhwc_probe.jsonl; tail -n 5 gpurun_out/r2d_nhwc_probe.err
