mkdir -p gpurun_out
rm -f gpurun_out/r2f_*
( KFACB_TRD_PROF=1 KFACB_TRD_TEAM=32 timeout 200 python tools/eigh_probe.py 4608 ) >> gpurun_out/r2f_probe.log 2>&1
( timeout 200 python tools/eigh_probe.py 4608 ) >> gpurun_out/r2f_probe.log 2>&1
( timeout 200 python tools/eigh_probe.py 129 147 513 1000 2304 ) >> gpurun_out/r2f_probe.log 2>&1
( timeout 1200 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 900 -x -k eigh 2>&1 | tail -5 ) > gpurun_out/r2f_tests.log 2>&1
( timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2f_bench.err | tail -1 ) > gpurun_out/r2f_bench_n1.json
