mkdir -p gpurun_out
rm -f gpurun_out/rc.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/test_all.log 2>&1; echo "tests rc=$?" >> gpurun_out/rc.txt
for inner in 1 2 3 10; do
  KFACB_INNER_SWEEPS=$inner timeout 300 python tools/eigh_probe.py 4608 >> gpurun_out/probe.log 2>&1
  KFACB_INNER_SWEEPS=$inner timeout 300 python tools/eigh_probe.py 785 512 2304 1152 >> gpurun_out/probe.log 2>&1
done
timeout 900 python bench.py --steps 3 --warmup 1 > gpurun_out/bench_r50.log 2>&1; echo "bench rc=$?" >> gpurun_out/rc.txt
tail -8 gpurun_out/test_all.log; cat gpurun_out/probe.log; cat gpurun_out/rc.txt; tail -c 3000 gpurun_out/bench_r50.log
