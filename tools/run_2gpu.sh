mkdir -p gpurun_out
T=r2i
rm -f gpurun_out/${T}_*
nvidia-smi -L > gpurun_out/${T}_gpus.txt 2>&1
( timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q --timeout 600 -k "world_size_2 or overlapped" 2>&1 | tail -15 ) > gpurun_out/${T}_tests.log 2>&1
KFACB_EXCHANGE=p2p bash tools/run_multi.sh 2 ${T} p2p:
KFACB_EXCHANGE=nccl bash tools/run_multi.sh 2 ${T} nccl:
