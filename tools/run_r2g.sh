mkdir -p gpurun_out
timeout 900 python tools/diag_grads.py cfg1 > gpurun_out/r2g_diag_grads.txt 2> gpurun_out/r2g_diag_grads.err; echo "diag rc=$?"; cat gpurun_out/r2g_diag_grads.txt; tail -n 5 gpurun_out/r2g_diag_grads.err
