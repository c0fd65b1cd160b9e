mkdir -p gpurun_out; rm -f gpurun_out/probe.log gpurun_out/selftest.log
for v in 0 1 2 3 4; do timeout 60 ./tools/tc_selftest.bin $v >> gpurun_out/selftest.log 2>&1; done
cat gpurun_out/selftest.log
P="timeout 300 python tools/eigh_probe.py"
KFACB_INNER_SWEEPS=1 $P 785 512 1152 >> gpurun_out/probe.log 2>&1
KFACB_INNER_SWEEPS=2 $P 785 512 1152 >> gpurun_out/probe.log 2>&1
KFACB_INNER_SWEEPS=1 KFACB_NS=0 $P 785 512 1152 >> gpurun_out/probe.log 2>&1
KFACB_INNER_SWEEPS=1 PROBE_GRADED=1 $P 513 1000 >> gpurun_out/probe.log 2>&1
KFACB_INNER_SWEEPS=1 $P 4608 >> gpurun_out/probe.log 2>&1
KFACB_INNER_SWEEPS=2 $P 4608 >> gpurun_out/probe.log 2>&1
KFACB_INNER_SWEEPS=1 KFACB_CONV_TOL_E6=10 $P 4608 >> gpurun_out/probe.log 2>&1
KFACB_INNER_SWEEPS=1 $P 4608 4608 4608 2304 2304 2048 2048 1152 1152 1024 1024 >> gpurun_out/probe.log 2>&1
grep -E "^env|^  n=|Error|error" gpurun_out/probe.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/test_all.log 2>&1; echo "tests rc=$?"
tail -12 gpurun_out/test_all.log
