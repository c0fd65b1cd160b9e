"""Drives the vendored, unmodified reference (``oracle/_ref``, see make_ref.py) through its own public API
on a BASELINE workload: ``kfac.get_kfac_module(variant)(model, ...)`` + ``preconditioner.step()`` with the
synthetic activations / output gradients / parameter gradients injected exactly as tests/golden/make_golden.py
does (hook_enabled=False, m_a / m_g / .grad set directly).  Baseline / test infrastructure only.
"""
from __future__ import annotations

import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_DIR = os.path.join(HERE, '_ref')


def available() -> bool:
    return os.path.isdir(os.path.join(REF_DIR, 'kfac'))


def import_reference():
    """``import kfac`` resolving to oracle/_ref/kfac (not to the repo's drop-in alias package)."""
    if not available():
        raise RuntimeError('oracle/_ref is empty: run `python oracle/make_ref.py` where /root/reference exists')
    saved = list(sys.path)
    sys.path = [REF_DIR] + [p for p in sys.path if os.path.abspath(p or '.') != ROOT]
    for k in [k for k in sys.modules if k == 'kfac' or k.startswith('kfac.') or k == 'horovod' or k.startswith('horovod.')]:
        del sys.modules[k]
    try:
        import kfac  # noqa
        import kfac.backend as backend
    finally:
        sys.path = saved
    assert os.path.abspath(kfac.__file__).startswith(REF_DIR), kfac.__file__
    return kfac, backend


class ReferenceStep:
    """One rank (world_size 1) of the reference preconditioner on a layer table of kfac_pytorch_b200.workloads."""

    def __init__(self, workload: str, device: str = 'cpu', seed: int = 1234):
        import torch.distributed as dist
        from kfac_pytorch_b200.workloads import WORKLOADS, LayerStack, io_shapes
        kfac, backend = import_reference()
        if not dist.is_initialized():
            os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
            os.environ.setdefault('MASTER_PORT', '29577')
            dist.init_process_group('gloo', rank=0, world_size=1)
        backend.comm = None
        backend.init('Torch')
        table_fn, variant, batch, hyper = WORKLOADS[workload]
        self.layers = table_fn()
        self.variant, self.batch = variant, batch
        self.dev = torch.device(device)
        model = LayerStack([ls.module for ls in self.layers]).to(self.dev)
        self.pre = kfac.get_kfac_module(variant)(model, lr=hyper['lr'], damping=hyper['damping'],
                                                 kl_clip=hyper['kl_clip'], factor_decay=0.95,
                                                 fac_update_freq=1, kfac_update_freq=1, hook_enabled=False)
        self.mods = self.pre.modules
        gen = torch.Generator().manual_seed(seed)
        self.a, self.g, self.w, self.b = [], [], [], []
        for ls, m in zip(self.layers, self.mods):
            sa, sg = io_shapes(ls, batch)
            self.a.append(torch.randn(*sa, generator=gen).to(self.dev))
            self.g.append((torch.randn(*sg, generator=gen) / batch).to(self.dev))
            self.w.append(torch.randn(m.weight.shape, generator=gen).to(self.dev))
            self.b.append(None if m.bias is None else torch.randn(m.bias.shape, generator=gen).to(self.dev))

    def step(self) -> float:
        """Inject + ``preconditioner.step()``; returns the seconds of the step() call alone."""
        for m, a, g, w, b in zip(self.mods, self.a, self.g, self.w, self.b):
            self.pre.m_a[m] = a
            self.pre.m_g[m] = g
            m.weight.grad = w.clone()
            if b is not None:
                m.bias.grad = b.clone()
        if self.dev.type == 'cuda':
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        self.pre.step()
        if self.dev.type == 'cuda':
            torch.cuda.synchronize()
        return time.perf_counter() - t0
