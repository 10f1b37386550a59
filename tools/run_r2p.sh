mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_fullsize.py > gpurun_out/r2p_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/r2p_pytest_gpu.log
