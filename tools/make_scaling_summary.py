"""profiles/r02_scaling_summary.txt from the committed bench lines (profiles/r02_bench_*.json)."""
import glob
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = []
for p in sorted(glob.glob(os.path.join(ROOT, 'profiles', 'r02_bench_n*.json'))):
    d = json.load(open(p))
    ph = {k: round(v['ms'], 2) for k, v in d['phases'].items()}
    chk = d.get('eig_check') or {}
    rows.append((d['config']['workload'].split()[0], d['n_gpus'], d['config'].get('partition'), os.path.basename(p), d['value'],
                 d['e2e']['value'], ph, chk, d['config'].get('rank_load_sum9n3_gflop'), d['gpu_launches'] / max(1, d['steps'])))
rows.sort(key=lambda r: (r[0], r[1], r[2] or ''))
out = ['# round 2: bench.py lines on B200 (steps 5, warm-up 3; value = ms per preconditioner.step(), max over ranks; eig_check after the timed loop)',
       '# final build: resnet50 N=1, N=2 (p2p), N=4 round_robin, N=8 round_robin, vgg16 N=2 / N=4.  Earlier round-2 build (before the',
       '# deflation-tolerance fix, the lighter team barrier, the lower-triangle rank-2k update and the inverse rewrite; the GPU budget',
       '# ran out before they could be repeated): the greedy lines, bert_base, *_earlier_build.', '']
for w, n, part, f, v, e, ph, chk, load, lps in rows:
    c = ('recon %.1e orth %.1e' % (chk['reconstruction'], chk['orthogonality'])) if 'reconstruction' in chk else ('residual %.1e' % chk.get('residual', float('nan')))
    out.append('%-10s N=%d %-11s value %7.1f ms  e2e %7.1f ms  check: %s ok=%s  launches/step(rank0) %d' % (w, n, part, v, e, c, chk.get('ok'), lps))
    out.append('           phases(rank 0) %s' % ph)
    out.append('           per-rank sum 9n^3 [GF] %s   file %s' % (load, f))
open(os.path.join(ROOT, 'profiles', 'r02_scaling_summary.txt'), 'w').write('\n'.join(out) + '\n')
print('\n'.join(out))
