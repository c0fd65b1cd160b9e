mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_fullsize_gpu.py tests/test_kernels_gpu.py -m gpu -q --timeout 1200 -s 2>&1 | grep -E "FULLSIZE|passed|failed|FAILED|^E " > gpurun_out/test_fullsize.log; cat gpurun_out/test_fullsize.log | tail -15
