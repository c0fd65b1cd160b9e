"""End-to-end effect of overlap_factor_capture on a REAL training iteration: torchvision
ResNet-50 (random init), bs 32, 224^2, forward + backward + preconditioner.step() on one B200, with the factor SYRKs
either accumulated inside step() (reference order, kfac_preconditioner_inv_dp.py:60-90) or launched from the hooks on
the capture stream.  Prints one JSON line per (kfac_update_freq, mode).

    python tools/overlap_demo.py [--iters 12] [--warmup 4]
"""
import argparse
import json
import os
import sys

import torch
import torchvision

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kfac_pytorch_b200 as kfac  # noqa: E402
from kfac_pytorch_b200 import backend  # noqa: E402


def run(overlap: bool, kfac_freq: int, iters: int, warmup: int):
    backend.comm = None
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    model = torchvision.models.resnet50(weights=None).to(dev)
    pre = kfac.get_kfac_module('eigen_dp')(model, lr=0.0125, damping=0.002, kl_clip=0.001, factor_decay=0.95,
                                           fac_update_freq=1, kfac_update_freq=kfac_freq, overlap_factor_capture=overlap)
    gen = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(32, 3, 224, 224, device=dev, generator=gen)
    y = torch.randint(0, 1000, (32,), device=dev, generator=gen)
    crit = torch.nn.CrossEntropyLoss(label_smoothing=0.1)

    def it():
        model.zero_grad(set_to_none=False)
        crit(model(x), y).backward()
        pre.step()

    for _ in range(warmup):
        it()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        it()
    e1.record()
    torch.cuda.synchronize()
    g = torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.grad is not None])
    return e0.elapsed_time(e1) / iters, float(g.double().norm())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=12)
    ap.add_argument('--warmup', type=int, default=4)
    a = ap.parse_args()
    for freq in (1, 10):
        res = {}
        for overlap in (False, True):
            iters = a.iters if freq == 1 else 2 * a.iters + (-2 * a.iters) % freq       # whole update periods
            warm = a.warmup if freq == 1 else freq
            ms, gn = run(overlap, freq, iters, warm)
            res[overlap] = (ms, gn)
        print(json.dumps(dict(model='torchvision resnet50 bs32 224^2 fwd+bwd+step, eigen_dp, fac_update_freq=1',
                              kfac_update_freq=freq, ms_per_iter_capture_in_step=res[False][0],
                              ms_per_iter_overlapped_capture=res[True][0],
                              saved_ms=res[False][0] - res[True][0],
                              final_grad_norm=[res[False][1], res[True][1]])), flush=True)


if __name__ == '__main__':
    main()
