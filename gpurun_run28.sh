mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 1800 python -m pytest tests -m gpu -q --timeout 1200 2>&1 | tail -3
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --workload vgg16 --e2e-steps 1 > gpurun_out/bench_vgg16.log 2>&1; python -c "
import json;d=json.loads(open('gpurun_out/bench_vgg16.log').read().strip().splitlines()[-1]);print('vgg16',round(d['value'],2),{k:round(v['ms'],2) for k,v in d['phases'].items()})"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/bench_r50_n2.log 2>&1; echo "bench n2 rc=$?"
grep '"metric"' gpurun_out/bench_r50_n2.log | tail -1 > gpurun_out/bench_r50_n2.json; python -c "
import json;d=json.load(open('gpurun_out/bench_r50_n2.json'));print(d['n_gpus'],d['value'],{k:round(v['ms'],2) for k,v in d['phases'].items()},d['e2e']['value'],d['clocks'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2>&1 | grep impl | cut -c1-200
