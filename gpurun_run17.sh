mkdir -p gpurun_out; rm -f gpurun_out/probe.log
P="timeout 300 python tools/eigh_probe.py"
timeout 120 python tools/eigh_probe.py 300 785 >> gpurun_out/probe.log 2>&1 || echo "small probe failed/hung"
$P 4608 >> gpurun_out/probe.log 2>&1
KFACB_UPD_TILES=4 $P 4608 >> gpurun_out/probe.log 2>&1
KFACB_UPD_TILES=12 $P 4608 >> gpurun_out/probe.log 2>&1
grep -E "^env|^  n=|Error|error" gpurun_out/probe.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/test_all.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/test_all.log
KFACB_MAX_SWEEPS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 120 --csv --log-file gpurun_out/launches_eigh4608.csv python tools/eigh_probe.py 4608 > gpurun_out/ncu_probe.log 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/launches_eigh4608.csv')))
hdr=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hdr]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
agg=collections.defaultdict(list)
for r in rows[hdr+1:]:
    if len(r)>vi:
        try: agg[r[ki][:60]].append(float(r[vi].replace(',','')))
        except: pass
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1]))[:5]:
    print('%-62s n=%4d mean=%9.1f us total=%9.1f ms'%(k,len(v),sum(v)/len(v)/1e3,sum(v)/1e6))
PY
timeout 900 python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r50.log 2>&1; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/bench_r50.log').read().strip().splitlines()[-1]);print(d['value'],d['phases'],d['e2e'],d['gpu_launches'])"
