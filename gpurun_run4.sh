mkdir -p gpurun_out
timeout 120 ./tools/tc_selftest.bin > gpurun_out/selftest.log 2>&1
cat gpurun_out/selftest.log
KFACB_MAX_SWEEPS=1 KFACB_INNER_SWEEPS=1 timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tools/eigh_probe.py 300 > gpurun_out/sanitizer.log 2>&1
grep -v "^$" gpurun_out/sanitizer.log | head -60
