mkdir -p gpurun_out
T=${1:-r3i}
rm -f gpurun_out/${T}_*
( timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_fullsize_gpu.py -m gpu -q --timeout 900 -x -k "inverse or vgg" 2>&1 | tail -3 ) > gpurun_out/${T}_tests.log 2>&1
( timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload vgg16 2>>gpurun_out/${T}_bench.err | tail -1 ) > gpurun_out/${T}_bench_vgg_n1.json
