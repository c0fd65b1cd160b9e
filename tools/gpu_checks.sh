#!/bin/bash
# What is run on the B200 box (through `gpurun -- 'bash tools/gpu_checks.sh'`); results land in gpurun_out/ and the
# summaries that matter are copied to profiles/ (profiles/README.md lists them).  Round-2 pieces:
#   tools/run_full.sh <tag>     whole `-m gpu` suite (writes gpurun_out/r02_parity.txt), bench N=1 with the CPU arm, smoke()
#   tools/run_ncu.sh <tag>      launch list of the bench command + `ncu --set full` of trd_kernel and the GEMM users
#   tools/run_multi.sh N tag .. torchrun bench lines on N GPUs (inside `gpurun --gpus N`)
#   tools/trd_calib.py n team   one point of the tridiagonalisation cost-model calibration
#   tools/eigh_probe.py n ..    eigensolver timing / accuracy (KFACB_TRD_PROF=1: per-team phase profile)
#   tools/parity_stage_debug.py stage-by-stage error of one full-size layer step vs the fp64 oracle
#   tools/overlap_demo.py       torchvision ResNet-50 iteration with / without overlap_factor_capture
set -x
bash tools/run_full.sh r02
bash tools/run_ncu.sh r02
