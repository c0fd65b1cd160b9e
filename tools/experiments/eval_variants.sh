#!/bin/bash
# On the B200 box (gpurun -- 'bash tools/experiments/eval_variants.sh'), after tools/experiments/build_variants.py:
# one decomposition of a 4608-dim factor (the critical path of a rank at N=8) and of an N=1-like mix per variant.
mkdir -p gpurun_out
L=gpurun_out/variants.log; : > $L
run() { echo "## $*" >> $L; env KFACB_PROFILE=1 "$@" timeout 120 python tools/eigh_probe.py $NS 2>&1 | grep -E "profile|env=|n=" | tail -4 >> $L; }
B=tools/_bin
for NS in "4608" "4608 2304 2304 2304 1152 1152 1024 1024 576 512 512 256 256 2048 2048"; do
  run BASELINE=1
  [ -f $B/libkfacb_upd_transposed.so ] && run KFACB_PROBE_LIB=$B/libkfacb_upd_transposed.so KFACB_UPD_TRANSPOSED=1
  for k in 3 5 7; do
    [ -f $B/libkfacb_low_precision_sweeps.so ] && run KFACB_PROBE_LIB=$B/libkfacb_low_precision_sweeps.so KFACB_LOW_SWEEPS=$k
  done
done
cat $L
