mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r50_n1.log 2>&1; echo "bench n1 rc=$?"
tail -1 gpurun_out/bench_r50_n1.log | cut -c1-400
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench ref rc=$?"
tail -1 gpurun_out/bench_ref.log | cut -c1-300
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/launches_bench.csv')))
hdr=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hdr]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
agg=collections.defaultdict(list)
for r in rows[hdr+1:]:
    if len(r)>vi:
        try: agg[r[ki][:70]].append(float(r[vi].replace(',','')))
        except: pass
tot=sum(sum(v) for v in agg.values())
out=[]
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1]))[:25]:
    out.append('%-72s n=%6d mean=%9.1f us total=%9.1f ms share=%5.1f%%'%(k,len(v),sum(v)/len(v)/1e3,sum(v)/1e6,100*sum(v)/tot))
open('gpurun_out/launches_bench_summary.txt','w').write('\n'.join(out)+'\n')
print('\n'.join(out))
PY
