mkdir -p gpurun_out
for tol in 300 1000; do
echo "== KFACB_CONV_TOL_E6=$tol"
KFACB_CONV_TOL_E6=$tol timeout 900 python -m pytest tests/test_fullsize_gpu.py -m gpu -q --timeout 600 -s 2>&1 | grep -E "FULLSIZE|passed|failed"
KFACB_CONV_TOL_E6=$tol timeout 300 python tools/eigh_probe.py 4608 2304 1152 785 2>&1 | grep -E "^env|n="
PROBE_GRADED=1 KFACB_CONV_TOL_E6=$tol timeout 300 python tools/eigh_probe.py 1000 513 2>&1 | grep -E "^env|n="
KFACB_CONV_TOL_E6=$tol timeout 600 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('bench',d['value'])"
done
