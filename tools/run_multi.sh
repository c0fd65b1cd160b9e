#!/bin/bash
# usage: bash tools/run_multi.sh <ngpus> <tag> [extra bench args...]   (inside gpurun --gpus N)
N=$1; TAG=$2; shift; shift
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt 2>&1
run() {  # name, env..., -- args
  name=$1; shift
  ( timeout 600 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --steps 5 --warmup 3 $EXTRA 2>gpurun_out/${TAG}_${name}.err | tail -1 ) > gpurun_out/${TAG}_${name}.json
}
EXTRA="$*"
run p2p KFACB_EXCHANGE=p2p
run nccl KFACB_EXCHANGE=nccl
