mkdir -p gpurun_out
T=${1:-r2q}
rm -f gpurun_out/${T}_*
for nt in "256 1" "512 1" "1024 1" "1024 2" "2048 1" "2048 3" "2048 8" "2304 4" "2304 8" "2304 16" "4608 16" "4608 31" "4608 64" "4608 147"; do
  ( timeout 120 python tools/trd_calib.py $nt 2>&1 | grep calib ) >> gpurun_out/${T}_calib.log
done
