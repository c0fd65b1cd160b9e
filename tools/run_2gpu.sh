mkdir -p gpurun_out
T=r3k
rm -f gpurun_out/${T}_*
nvidia-smi -L > gpurun_out/${T}_gpus.txt 2>&1
( timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_kernels_gpu.py -m gpu -q --timeout 600 -k "world_size_2 or fused or pack_grads" 2>&1 | tail -6 ) > gpurun_out/${T}_tests.log 2>&1
bash tools/run_multi.sh 2 ${T} "r50_rr:" "vgg_rr:--workload vgg16"
