mkdir -p gpurun_out; rm -f gpurun_out/probe.log
P="timeout 300 python tools/eigh_probe.py"
timeout 120 python tools/eigh_probe.py 300 785 >> gpurun_out/probe.log 2>&1 || echo "small probe failed/hung"
KFACB_PROFILE=1 $P 4608 >> gpurun_out/probe.log 2>&1
grep -E "^env|^  n=|Error|error|profile" gpurun_out/probe.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/test_all.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/test_all.log
KFACB_PROFILE=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 2>&1 | grep -E "kfacb eigh profile" | tail -2
timeout 900 python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r50.log 2>&1; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/bench_r50.log').read().strip().splitlines()[-1]);print(d['value'],d['phases'],d['e2e'],d['gpu_launches'])"
