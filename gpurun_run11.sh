mkdir -p gpurun_out; rm -f gpurun_out/probe.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/test_all.log 2>&1; echo "tests rc=$?"
tail -6 gpurun_out/test_all.log
P="timeout 300 python tools/eigh_probe.py"
$P 4608 >> gpurun_out/probe.log 2>&1
KFACB_UPD_REVERSE=0 $P 4608 >> gpurun_out/probe.log 2>&1
KFACB_GEMM_SIMT=1 $P 4608 >> gpurun_out/probe.log 2>&1
grep -E "^env|^  n=|Error|error" gpurun_out/probe.log
timeout 900 python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r50.log 2>&1; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/bench_r50.log').read().strip().splitlines()[-1]);print(d['value'],d['phases'],d['e2e'],d['gpu_launches'])"
KFACB_GEMM_SIMT=1 timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --workload vgg16 > gpurun_out/bench_vgg_simt.log 2>&1
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --workload vgg16 > gpurun_out/bench_vgg.log 2>&1
for f in bench_vgg_simt bench_vgg; do python -c "
import json,sys;d=json.loads(open('gpurun_out/$f.log').read().strip().splitlines()[-1]);print('$f',d['value'],d['phases'])"; done
