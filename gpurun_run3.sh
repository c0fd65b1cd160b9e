mkdir -p gpurun_out; rm -f gpurun_out/rc.txt gpurun_out/probe.log
P="timeout 300 python tools/eigh_probe.py"
KFACB_JACOBI_SIMT=1 KFACB_INNER_SWEEPS=1 $P 785 512 1152 >> gpurun_out/probe.log 2>&1
KFACB_INNER_SWEEPS=1 $P 785 512 1152 >> gpurun_out/probe.log 2>&1
KFACB_INNER_SWEEPS=1 KFACB_NS=0 $P 785 512 1152 >> gpurun_out/probe.log 2>&1
KFACB_INNER_SWEEPS=2 $P 785 512 1152 >> gpurun_out/probe.log 2>&1
KFACB_INNER_SWEEPS=1 PROBE_GRADED=1 $P 513 1000 >> gpurun_out/probe.log 2>&1
KFACB_INNER_SWEEPS=10 PROBE_GRADED=1 $P 513 1000 >> gpurun_out/probe.log 2>&1
KFACB_INNER_SWEEPS=1 $P 4608 >> gpurun_out/probe.log 2>&1
KFACB_INNER_SWEEPS=2 $P 4608 >> gpurun_out/probe.log 2>&1
KFACB_INNER_SWEEPS=1 KFACB_CONV_TOL_E6=10 $P 4608 >> gpurun_out/probe.log 2>&1
KFACB_JACOBI_SIMT=1 KFACB_INNER_SWEEPS=1 $P 4608 >> gpurun_out/probe.log 2>&1
cat gpurun_out/probe.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/test_all.log 2>&1; echo "tests rc=$?" >> gpurun_out/rc.txt
tail -5 gpurun_out/test_all.log
