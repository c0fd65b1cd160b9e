mkdir -p gpurun_out
T=${1:-r3c}
rm -f gpurun_out/${T}_*
( timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 900 -x -k eigh 2>&1 | tail -3 ) > gpurun_out/${T}_tests.log 2>&1
( timeout 600 python tools/parity_stage_debug.py conv1 2>&1 | grep -v Warn ) >> gpurun_out/${T}_stage.log
( timeout 900 python -m pytest tests/test_fullsize_gpu.py -m gpu -q --timeout 900 -s 2>&1 | grep -E "FULLSIZE|passed|failed" ) > gpurun_out/${T}_parity.log 2>&1
( timeout 200 python tools/eigh_probe.py 129 147 513 1000 2304 ) >> gpurun_out/${T}_probe.log 2>&1
( timeout 120 python tools/trd_calib.py 4608 0 2>&1 | grep calib ) >> gpurun_out/${T}_probe.log
( timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/${T}_bench.err | tail -1 ) > gpurun_out/${T}_bench_n1.json
