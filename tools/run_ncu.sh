#!/bin/bash
# ncu evidence for profiles/ (one GPU): launch list of the bench command + full captures of the dominant kernels.
mkdir -p gpurun_out
T=${1:-r02}
# launch list (cold-cache, serialised; the kernels' SHARE of the step is what must agree with bench.py)
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${T}_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${T}_launches_bench.out 2> gpurun_out/${T}_launches_bench.err
# full capture: the persistent tridiagonalisation kernel of the ResNet-50 step (dominant kernel)
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:trd_kernel -s 1 -c 1 -o gpurun_out/${T}_trd_full -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${T}_trd_full.out 2>&1
ncu -i gpurun_out/${T}_trd_full.ncu-rep --page raw --csv > gpurun_out/${T}_trd_full_raw.csv 2>/dev/null
# full capture: factor SYRK, preconditioning GEMMs, back-transformation GEMM (first instances of each)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"tc_syrk_ema_kernel|bt_upd_kernel|precond|dc_gemm_kernel" -c 12 -o gpurun_out/${T}_gemms_full -f \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${T}_gemms_full.out 2>&1
ncu -i gpurun_out/${T}_gemms_full.ncu-rep --page raw --csv > gpurun_out/${T}_gemms_full_raw.csv 2>/dev/null
ls -la gpurun_out/${T}_*
