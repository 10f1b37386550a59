mkdir -p gpurun_out
timeout 900 python tools/diag_convs.py cfg1 > gpurun_out/r2h_diag_convs.txt 2> gpurun_out/r2h_diag_convs.err; echo "diag rc=$?"; cat gpurun_out/r2h_diag_convs.txt; tail -n 5 gpurun_out/r2h_diag_convs.err
