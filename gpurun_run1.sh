mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/host.txt; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" >> gpurun_out/host.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x -k "kernels" > gpurun_out/test_kernels.log 2>&1; echo "kernels rc=$?" >> gpurun_out/rc.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "parity" > gpurun_out/test_parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/rc.txt
timeout 300 python bench.py --workload mlp --steps 5 --warmup 3 > gpurun_out/bench_mlp.log 2>&1; echo "bench_mlp rc=$?" >> gpurun_out/rc.txt
tail -5 gpurun_out/test_kernels.log; tail -5 gpurun_out/test_parity.log; cat gpurun_out/rc.txt
