mkdir -p gpurun_out
rm -f gpurun_out/r3s_*
bash tools/run_multi.sh 8 r3s "r50_rr:"
