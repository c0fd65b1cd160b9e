#!/bin/bash
# usage: bash tools/run_multi.sh <ngpus> <tag> <name>:<bench args> [<name>:<bench args> ...]   (inside gpurun --gpus N)
# every entry is one torchrun of bench.py on N ranks; default exchange (peer memory) unless KFACB_EXCHANGE is in the env
N=$1; TAG=$2; shift; shift
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt 2>&1
for entry in "$@"; do
  name=${entry%%:*}; extra=${entry#*:}
  ( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline $extra 2>gpurun_out/${TAG}_${name}.err | tail -1 ) > gpurun_out/${TAG}_${name}.json
done
