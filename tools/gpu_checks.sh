#!/bin/bash
# What is run on the B200 box (through `gpurun -- 'bash tools/gpu_checks.sh'`); results land in gpurun_out/
# and the summaries that matter are copied to profiles/.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 1200 -s 2>&1 | grep -E "FULLSIZE|passed|failed|FAILED" | tee gpurun_out/test_all.log
python bench.py --steps 5 --warmup 3 | tail -1 > gpurun_out/bench_r50_n1.json
python bench.py --impl reference --steps 2 --warmup 1 | tail -1 > gpurun_out/bench_reference_arm.json
# single 4608-dim factor (the critical path of a rank at N=8): per-kernel totals
KFACB_PROFILE=1 python tools/eigh_probe.py 4608 > gpurun_out/probe4608.log 2>&1
# launch list of the bench command (first 7000 launches) and a full capture of the three Jacobi kernels
ncu --metrics gpu__time_duration.sum --clock-control none -c 7000 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 > /dev/null
KFACB_MAX_SWEEPS=1 ncu --set full --clock-control none --import-source on \
    -k regex:"bj_h_tc_kernel|bj_update_tc_kernel|bj_solve_kernel" -s 30 -c 3 -o gpurun_out/prof_jacobi \
    python tools/eigh_probe.py 4608 > /dev/null
