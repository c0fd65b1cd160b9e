mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 3 --warmup 2 > gpurun_out/bench_r50_n8.log 2>&1; echo "bench n8 rc=$?"
grep '"metric"' gpurun_out/bench_r50_n8.log | tail -1 > gpurun_out/bench_r50_n8.json; python -c "
import json;d=json.load(open('gpurun_out/bench_r50_n8.json'));print(d['n_gpus'],d['value'],{k:round(v['ms'],2) for k,v in d['phases'].items()},d['e2e']['value'])" || tail -20 gpurun_out/bench_r50_n8.log
