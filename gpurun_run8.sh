mkdir -p gpurun_out
KFACB_MAX_SWEEPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bj_h_tc_kernel|bj_solve_kernel|bj_update_tc_kernel" -s 60 -c 6 -o gpurun_out/prof_jacobi python tools/eigh_probe.py 4608 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
