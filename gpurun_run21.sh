mkdir -p gpurun_out; rm -f gpurun_out/probe.log
KFACB_PROFILE=1 timeout 300 python tools/eigh_probe.py 4608 2>&1 | grep -E "profile|n=4608|ms=" | tail -3
KFACB_PROFILE=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 2>&1 | grep -E "kfacb eigh profile" | tail -1
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --workload bert_base --e2e-steps 1 > gpurun_out/bench_bert.log 2>&1; echo "bert rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/bench_bert.log').read().strip().splitlines()[-1]);print('bert',d['value'],d['phases'])" || tail -5 gpurun_out/bench_bert.log
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --workload vgg16 --e2e-steps 1 > gpurun_out/bench_vgg.log 2>&1; echo "vgg rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/bench_vgg.log').read().strip().splitlines()[-1]);print('vgg',d['value'],d['phases'])" || tail -5 gpurun_out/bench_vgg.log
