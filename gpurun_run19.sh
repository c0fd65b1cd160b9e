mkdir -p gpurun_out
KFACB_MAX_SWEEPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bj_h_tc_kernel|bj_update_tc_kernel" -s 16 -c 2 -o gpurun_out/prof_hu python tools/eigh_probe.py 4608 > gpurun_out/ncu_hu.log 2>&1
ls -la gpurun_out/prof_hu.ncu-rep
