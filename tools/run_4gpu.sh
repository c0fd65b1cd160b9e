mkdir -p gpurun_out
rm -f gpurun_out/r3t_*
bash tools/run_multi.sh 4 r3t "vgg_rr:--workload vgg16" "r50_rr:"
