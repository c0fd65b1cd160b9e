mkdir -p gpurun_out
KFACB_MAX_SWEEPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bj_solve_kernel" -s 8 -c 1 -o gpurun_out/prof_solve python tools/eigh_probe.py 4608 > gpurun_out/ncu_solve.log 2>&1
ls -la gpurun_out/prof_solve.ncu-rep
