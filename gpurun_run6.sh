mkdir -p gpurun_out
KFACB_MAX_SWEEPS=1 KFACB_INNER_SWEEPS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/launches_eigh4608.csv python tools/eigh_probe.py 4608 > gpurun_out/ncu_probe.log 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/launches_eigh4608.csv')))
hdr=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hdr]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); gi=h.index('Grid Size') if 'Grid Size' in h else None
agg=collections.defaultdict(list)
for r in rows[hdr+1:]:
    if len(r)>vi:
        try: agg[r[ki][:60]].append(float(r[vi].replace(',','')))
        except: pass
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])):
    print('%-62s n=%4d mean=%9.1f us total=%9.1f ms'%(k,len(v),sum(v)/len(v)/1e3,sum(v)/1e6))
PY
