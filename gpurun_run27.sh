mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 1200 -s 2>&1 | grep -E "FULLSIZE|passed|failed|FAILED|^E  +assert" > gpurun_out/test_all.log; tail -8 gpurun_out/test_all.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r50_n1.log 2>&1; echo "bench n1 rc=$?"
tail -1 gpurun_out/bench_r50_n1.log | cut -c1-200
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench ref rc=$?"
for w in vgg16 bert_base mlp; do timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --workload $w --e2e-steps 1 > gpurun_out/bench_$w.log 2>&1; python -c "
import json;d=json.loads(open('gpurun_out/bench_$w.log').read().strip().splitlines()[-1]);print('$w',round(d['value'],2),{k:round(v['ms'],2) for k,v in d['phases'].items()})"; done
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 7000 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu rc=$?"
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
out=['# ncu --metrics gpu__time_duration.sum --clock-control none -c 7000 python bench.py --steps 1 --warmup 1 (first 7000 launches: warm-up step + most of one timed step)']
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1]))[:22]:
    out.append('%-72s n=%6d mean=%9.1f us total=%9.1f ms share=%5.1f%%'%(k,len(v),sum(v)/len(v)/1e3,sum(v)/1e6,100*sum(v)/tot))
open('gpurun_out/launches_bench_summary.txt','w').write('\n'.join(out)+'\n')
print('\n'.join(out[:8]))
PY
