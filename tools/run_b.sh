mkdir -p gpurun_out
rm -f gpurun_out/r2b_*.log
( KFACB_TRD_PROF=1 timeout 200 python tools/eigh_probe.py 4608 ) >> gpurun_out/r2b_probe.log 2>&1
( KFACB_TRD_PROF=1 KFACB_TRD_TEAM=32 timeout 200 python tools/eigh_probe.py 4608 ) >> gpurun_out/r2b_probe.log 2>&1
( KFACB_TRD_PROF=1 KFACB_TRD_TEAM=32 timeout 200 python tools/eigh_probe.py 1000 ) >> gpurun_out/r2b_probe.log 2>&1
( KFACB_TRD_PROF=1 KFACB_TRD_TEAM=1 timeout 200 python tools/eigh_probe.py 513 ) >> gpurun_out/r2b_probe.log 2>&1
( KFACB_TDC_DEBUG=1 timeout 200 python tools/eigh_probe.py 129 513 1000 ) >> gpurun_out/r2b_probe.log 2>&1
( KFACB_TDC_DEBUG=1 KFACB_GEMM_SIMT=1 timeout 200 python tools/eigh_probe.py 129 513 1000 ) >> gpurun_out/r2b_probe.log 2>&1
( KFACB_GEMM_SIMT=1 timeout 200 python tools/eigh_probe.py 129 513 1000 ) >> gpurun_out/r2b_probe.log 2>&1
