mkdir -p gpurun_out
rm -f gpurun_out/r2d_*
( KFACB_TRD_PROF=1 timeout 200 python tools/eigh_probe.py 4608 ) >> gpurun_out/r2d_probe.log 2>&1
( KFACB_TRD_PROF=1 KFACB_TRD_TEAM=32 timeout 200 python tools/eigh_probe.py 4608 ) >> gpurun_out/r2d_probe.log 2>&1
( timeout 200 python tools/eigh_probe.py 129 147 513 1000 2304 ) >> gpurun_out/r2d_probe.log 2>&1
( timeout 1200 python -m pytest tests/test_kernels_gpu.py tests/test_fullsize_gpu.py -m gpu -q --timeout 900 -s 2>&1 | grep -E "FULLSIZE|passed|failed|FAILED|Error" ) > gpurun_out/r2d_tests.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2d_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2d_ncu.log 2>&1
( timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2d_bench.err | tail -1 ) > gpurun_out/r2d_bench_n1.json
