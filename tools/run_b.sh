mkdir -p gpurun_out
T=${1:-r3j}
rm -f gpurun_out/${T}_*
( timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 900 -x -k "pack_grads" 2>&1 | tail -40 ) > gpurun_out/${T}_tests.log 2>&1
