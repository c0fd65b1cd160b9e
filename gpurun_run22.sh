mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/test_all.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/test_all.log
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --workload vgg16 --e2e-steps 1 > gpurun_out/bench_vgg.log 2>&1; echo "vgg rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/bench_vgg.log').read().strip().splitlines()[-1]);print('vgg',d['value'],d['phases'])" || tail -5 gpurun_out/bench_vgg.log
KFACB_MAX_SWEEPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bj_h_tc_kernel|bj_update_tc_kernel|bj_solve_kernel" -s 30 -c 3 -o gpurun_out/prof_r01_final python tools/eigh_probe.py 4608 > gpurun_out/ncu_final.log 2>&1
ls -la gpurun_out/prof_r01_final.ncu-rep
