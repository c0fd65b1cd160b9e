mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 600 python -m pytest tests -m gpu -q --timeout 500 -k "world_size_2 or factor_conv or test_eigh" > gpurun_out/test_ws2.log 2>&1; echo "ws2 tests rc=$?"
tail -4 gpurun_out/test_ws2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/bench_r50_n2.log 2>&1; echo "bench n2 rc=$?"
grep '"metric"' gpurun_out/bench_r50_n2.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['n_gpus'],d['value'],d['phases'],d['e2e']['value'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 2 --partition greedy > gpurun_out/bench_r50_n2_greedy.log 2>&1; echo "bench n2 greedy rc=$?"
grep '"metric"' gpurun_out/bench_r50_n2_greedy.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('greedy',d['n_gpus'],d['value'],d['phases'])"
tail -3 gpurun_out/bench_r50_n2.log | cut -c1-300
