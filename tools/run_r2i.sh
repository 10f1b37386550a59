mkdir -p gpurun_out
timeout 600 python tools/diag_convs.py cfg1 > gpurun_out/r2i_diag_convs.txt 2> gpurun_out/r2i_diag_convs.err; echo "diag rc=$?"; tail -n 4 gpurun_out/r2i_diag_convs.txt
timeout 600 python tools/diag_grads.py cfg1 > gpurun_out/r2i_diag_grads.txt 2> gpurun_out/r2i_diag_grads.err; echo "diag rc=$?"; tail -n 20 gpurun_out/r2i_diag_grads.txt
bash tools/run_r2f.sh
