mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2fin_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2fin_pytest_gpu.log; grep -E "== parity|grads " gpurun_out/r2fin_pytest_gpu.log | head
