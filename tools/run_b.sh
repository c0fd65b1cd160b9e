mkdir -p gpurun_out
T=r2h
rm -f gpurun_out/${T}_*
R50="147 $(yes 256 | head -26 | tr '\n' ' ') $(yes 512 | head -19 | tr '\n' ' ') 576 576 576 1000 $(yes 1024 | head -14 | tr '\n' ' ') 1152 1152 1152 1152 2048 2048 2048 2048 2048 2048 2049 2304 2304 2304 2304 2304 2304 4608 4608 4608"
( KFACB_TRD_PROF=1 timeout 300 python tools/eigh_probe.py $R50 2>&1 | grep -v "eig_err" ) > gpurun_out/${T}_r50_prof.log 2>&1
( timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/${T}_bench.err | tail -1 ) > gpurun_out/${T}_bench_n1.json
