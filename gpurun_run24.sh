mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/test_all.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/test_all.log
timeout 900 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_r50.log 2>&1
python -c "
import json;d=json.loads(open('gpurun_out/bench_r50.log').read().strip().splitlines()[-1]);print(d['value'],d['phases'])"
