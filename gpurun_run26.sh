mkdir -p gpurun_out; rm -f gpurun_out/probe.log
timeout 1800 python -m pytest tests -m gpu -q --timeout 1200 -s 2>&1 | grep -E "FULLSIZE|passed|failed|FAILED|^E " > gpurun_out/test_all.log; tail -12 gpurun_out/test_all.log
timeout 300 python tools/eigh_probe.py 4608 2304 1152 2>&1 | grep -E "^env|n=" 
KFACB_CLEANUP_SWEEP=0 timeout 300 python tools/eigh_probe.py 4608 2304 1152 2>&1 | grep -E "^env|n="
timeout 900 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_r50.log 2>&1
python -c "
import json;d=json.loads(open('gpurun_out/bench_r50.log').read().strip().splitlines()[-1]);print(d['value'],d['phases'])"
