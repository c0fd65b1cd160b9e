mkdir -p gpurun_out; rm -f gpurun_out/probe.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/test_all.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/test_all.log
P="timeout 300 python tools/eigh_probe.py"
$P 4608 4608 4608 2304 2304 2048 2048 1152 1152 1024 1024 >> gpurun_out/probe.log 2>&1
KFACB_EIGH_GROUPS=1 $P 4608 4608 4608 2304 2304 2048 2048 1152 1152 1024 1024 >> gpurun_out/probe.log 2>&1
KFACB_EIGH_GROUPS=2 $P 4608 4608 4608 2304 2304 2048 2048 1152 1152 1024 1024 >> gpurun_out/probe.log 2>&1
KFACB_EIGH_GROUPS=4 $P 4608 4608 4608 2304 2304 2048 2048 1152 1152 1024 1024 >> gpurun_out/probe.log 2>&1
grep -E "^env|Error|error" gpurun_out/probe.log
for g in 1 3 4; do KFACB_EIGH_GROUPS=$g timeout 900 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_r50_g$g.log 2>&1; python -c "
import json;d=json.loads(open('gpurun_out/bench_r50_g$g.log').read().strip().splitlines()[-1]);print('groups $g',d['value'],d['phases']['compute_inverse'],d['roofline'] and {k:d['roofline'][k] for k in ('achieved','frac','avg_launch_us','share_of_step')})"; done
