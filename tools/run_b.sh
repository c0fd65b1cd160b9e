mkdir -p gpurun_out
T=${1:-r3d}
rm -f gpurun_out/${T}_*
( timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 900 -x -k eigh 2>&1 | tail -3 ) > gpurun_out/${T}_tests.log 2>&1
( KFACB_TRD_PROF=1 KFACB_TRD_TEAM=147 timeout 200 python tools/eigh_probe.py 4608 2>&1 | grep "trd profile" | tail -1 ) >> gpurun_out/${T}_probe.log 2>&1
( KFACB_TRD_PROF=1 KFACB_TRD_TEAM=30 timeout 200 python tools/eigh_probe.py 4608 2>&1 | grep "trd profile" | tail -1 ) >> gpurun_out/${T}_probe.log 2>&1
( timeout 200 python tools/eigh_probe.py 4608 2>&1 | tail -2 ) >> gpurun_out/${T}_probe.log 2>&1
( timeout 200 python tools/eigh_probe.py 129 147 513 1000 2304 ) >> gpurun_out/${T}_probe.log 2>&1
( timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/${T}_bench.err | tail -1 ) > gpurun_out/${T}_bench_n1.json
