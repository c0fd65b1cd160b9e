mkdir -p gpurun_out
rm -f gpurun_out/r2t_*
bash tools/run_multi.sh 4 r2t "vgg_rr:--workload vgg16" "r50_rr:" "r50_greedy:--partition greedy"
