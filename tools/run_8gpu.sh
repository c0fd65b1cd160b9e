mkdir -p gpurun_out
rm -f gpurun_out/r2s_*
bash tools/run_multi.sh 8 r2s "r50_rr:" "r50_greedy:--partition greedy" "bert_rr:--workload bert_base" "bert_greedy:--workload bert_base --partition greedy"
