"""GPU probe: time kfacb_eigh_batched on Wishart matrices of the ResNet-50 factor
sizes and report sweeps / accuracy.  Tuning knobs come from the environment
(KFACB_MAX_SWEEPS, KFACB_SOLVE_SPLIT, KFACB_H_KCHUNK, KFACB_UPD_TILES, KFACB_PROFILE).   python tools/eigh_probe.py 4608 [2304 ...]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kfac_pytorch_b200.native import NativeEngine  # noqa: E402


def wishart(n, rows, gen, dev):
    x = torch.randn(rows, n, device=dev, generator=gen)
    a = x.t() @ x / rows
    return (0.05 * torch.eye(n, device=dev) + 0.95 * a).contiguous()


def graded(n, gen, dev):
    q, _ = torch.linalg.qr(torch.randn(n, n, device=dev, generator=gen, dtype=torch.float64))
    m = (q * torch.logspace(-8, 0, n, device=dev, dtype=torch.float64)) @ q.t()
    return (0.5 * (m + m.t())).float().contiguous()


def main():
    dev = torch.device('cuda:0')
    eng = NativeEngine(dev)
    if os.environ.get('KFACB_PROBE_LIB'):        # evaluate a variant library (tools/experiments/build_variants.py)
        from kfac_pytorch_b200.native import load_library
        eng.lib = load_library(os.environ['KFACB_PROBE_LIB'])
    ns = [int(x) for x in sys.argv[1:]] or [4608]
    gen = torch.Generator(device=dev).manual_seed(0)
    if os.environ.get('PROBE_GRADED'):
        mats = [graded(n, gen, dev) for n in ns]
    else:
        mats = [wishart(n, max(1568, 2 * n) if n < 4608 else 1568, gen, dev) for n in ns]
    Qs = [torch.empty_like(m) for m in mats]
    ds = [torch.empty(m.shape[0], device=dev) for m in mats]
    sw = torch.zeros(len(ns), dtype=torch.int32, device=dev)
    for it in range(2):
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        l0 = eng.launch_count()
        e0.record()
        eng.eigh(mats, Qs, ds, 1e-10, sw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    out = ['env=%s ns=%s ms=%.1f launches=%d sweeps=%s' % (
        {k: v for k, v in os.environ.items() if k.startswith(('KFACB_', 'PROBE_'))}, ns, ms,
        eng.launch_count() - l0, sw.tolist())]
    for m, q, d in zip(mats, Qs, ds):
        n = m.shape[0]
        if n <= 2304:
            ref = torch.linalg.eigvalsh(m.double())
            derr = (d.double() - ref).abs().max().item() / ref.abs().max().item()
        else:
            derr = float('nan')
        e = q.t() @ q - torch.eye(n, device=dev)
        orth = e.abs().max().item()
        orth2 = torch.linalg.matrix_norm(e.double(), 2).item() if n <= 2304 else e.norm().item()
        rec = ((q * d) @ q.t() - m).norm().item() / m.norm().item()
        out.append('  n=%d eig_err=%.2e orth_max=%.2e orth_2norm=%.2e recon=%.2e' % (n, derr, orth, orth2, rec))
    print('\n'.join(out), flush=True)


if __name__ == '__main__':
    main()
