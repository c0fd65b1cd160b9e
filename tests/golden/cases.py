"""Golden-case definitions shared by ``make_golden.py`` (which drives the real
reference), ``tests/test_oracle_golden.py`` (oracle vs reference fixtures)
and the ``-m gpu`` parity tests (CUDA path vs fixtures + oracle).

Every case is fully determined by seeds: model structure -> layer specs,
per-step injected tensors come from ``torch.Generator`` streams, so nothing
but the reference's OUTPUTS has to be stored in the fixtures.

Injection mode: ``hook_enabled=False``; before each
``step()`` the harness sets ``m_a[mod]``, ``m_g[mod]`` and the ``.grad`` of
weight/bias directly.  The hook path itself is covered by the separate
``hooks_small`` case whose captured tensors are stored in the fixture.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import torch
import torch.nn as nn


@dataclass
class Case:
    name: str
    variant: str                     # 'eigen_dp' | 'inverse_dp'
    build: callable                  # () -> nn.Module (modules named Linear/Conv2d)
    batch: int
    steps: int
    hw: Optional[Tuple[int, int]] = None      # input H,W of the first conv
    seq: Optional[int] = None                 # sequence length for 3-D linear inputs
    lr: float = 0.1
    damping: float = 0.003
    kl_clip: Optional[float] = 0.001
    factor_decay: float = 0.95
    fac_update_freq: int = 1
    kfac_update_freq: int = 1
    world_size: int = 1
    seed: int = 1234
    lr_schedule: Optional[List[float]] = None  # per-step lr written to param_groups[0]
    extra: dict = field(default_factory=dict)   # variant-specific constructor kwargs (e.g. distribute_layer_factors)


class Flatten(nn.Module):
    def forward(self, x):
        return x.reshape(x.shape[0], -1)


def mlp_cfg1():
    return nn.Sequential(nn.Linear(784, 512), nn.ReLU(), nn.Linear(512, 10))


def mlp_small():
    return nn.Sequential(nn.Linear(20, 16), nn.ReLU(), nn.Linear(16, 12, bias=False),
                         nn.ReLU(), nn.Linear(12, 10))


def convnet_bias():
    # VGG-like: 3x3 pad 1 convs WITH bias (odd dA = 9*Cin + 1) + classifier
    return nn.Sequential(nn.Conv2d(3, 16, 3, padding=1), nn.ReLU(),
                         nn.Conv2d(16, 24, 3, padding=1), nn.ReLU(),
                         nn.MaxPool2d(2), Flatten(), nn.Linear(24 * 4 * 4, 10))


def convnet_resnetish():
    # the conv shape classes of ResNet-50: 7x7 s2 p3, 1x1, 3x3 s1 p1, 3x3 s2 p1,
    # 1x1 s2 p0 (all bias-free) + fc with bias
    return nn.Sequential(nn.Conv2d(3, 8, 7, stride=2, padding=3, bias=False), nn.ReLU(),
                         nn.Conv2d(8, 16, 1, bias=False), nn.ReLU(),
                         nn.Conv2d(16, 16, 3, padding=1, bias=False), nn.ReLU(),
                         nn.Conv2d(16, 24, 3, stride=2, padding=1, bias=False), nn.ReLU(),
                         nn.Conv2d(24, 32, 1, stride=2, bias=False), nn.ReLU(),
                         nn.AdaptiveAvgPool2d(1), Flatten(), nn.Linear(32, 10))


def seq_linears():
    # BERT-like: Linear layers fed (B, T, D) tensors, all with bias
    return nn.Sequential(nn.Linear(24, 24), nn.Linear(24, 48), nn.GELU(),
                         nn.Linear(48, 24), nn.Linear(24, 2))


def hooks_small():
    return nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.ReLU(),
                         nn.Conv2d(8, 8, 3, stride=2, padding=1, bias=False), nn.ReLU(),
                         Flatten(), nn.Linear(8 * 4 * 4, 10))


CASES = [
    Case('mlp_cfg1_eigen', 'eigen_dp', mlp_cfg1, batch=32, steps=3),
    Case('mlp_cfg1_inverse', 'inverse_dp', mlp_cfg1, batch=32, steps=3),
    Case('mlp_small_eigen_freq2', 'eigen_dp', mlp_small, batch=8, steps=5,
         fac_update_freq=2, kfac_update_freq=2, lr_schedule=[0.1, 0.1, 0.05, 0.05, 0.01]),
    Case('convnet_bias_inverse', 'inverse_dp', convnet_bias, batch=6, steps=3, hw=(8, 8),
         damping=0.03),
    Case('convnet_bias_eigen', 'eigen_dp', convnet_bias, batch=6, steps=2, hw=(8, 8)),
    Case('convnet_resnetish_eigen', 'eigen_dp', convnet_resnetish, batch=4, steps=3,
         hw=(32, 32), damping=0.002, lr=0.0125),
    Case('seq_linears_eigen_noclip', 'eigen_dp', seq_linears, batch=4, steps=2, seq=6,
         kl_clip=None, damping=0.03),
    Case('mlp_small_eigen_ws2', 'eigen_dp', mlp_small, batch=8, steps=3, world_size=2),
    Case('convnet_bias_inverse_ws2', 'inverse_dp', convnet_bias, batch=6, steps=2, hw=(8, 8),
         damping=0.03, world_size=2),
    # model-parallel ("MPD") variants: full factor replicas + all-reduce
    Case('mlp_small_inverse_mpd_ws2', 'inverse', mlp_small, batch=8, steps=3, world_size=2),
    Case('mlp_small_eigen_mpd_ws2', 'eigen', mlp_small, batch=8, steps=3, world_size=2),
    Case('convnet_bias_eigen_mpd', 'eigen', convnet_bias, batch=6, steps=2, hw=(8, 8)),
    Case('mlp_small_eigen_mpd_factorwise_ws2', 'eigen', mlp_small, batch=8, steps=2, world_size=2,
         extra=dict(distribute_layer_factors=True)),
    Case('mlp_small_inverse_mpd_comminv_ws2', 'inverse', mlp_small, batch=8, steps=2, world_size=2,
         extra=dict(communicate_inverse_or_not=True)),
    Case('convnet_bias_inverse_mpd_ws2', 'inverse', convnet_bias, batch=6, steps=2, hw=(8, 8),
         damping=0.03, world_size=2),
]
CASES_BY_NAME = {c.name: c for c in CASES}


def layer_input_hw(model: nn.Module, hw):
    """Input (H, W) seen by each Conv2d of a Sequential, by shape propagation
    with a meta-free dry run on zeros (CPU, tiny)."""
    sizes = {}
    if hw is None:
        return sizes
    hooks = []

    def mk(m):
        def pre(mod, inp):
            sizes[mod] = tuple(inp[0].shape[-2:])
        return pre
    for m in model.modules():
        if m.__class__.__name__ == 'Conv2d':
            hooks.append(m.register_forward_pre_hook(mk(m)))
    first = [m for m in model.modules() if m.__class__.__name__ == 'Conv2d'][0]
    with torch.no_grad():
        model(torch.zeros(1, first.in_channels, *hw))
    for h in hooks:
        h.remove()
    return sizes


SUBSAMPLE_LIMIT = 4096
SUBSAMPLE_TARGET = 1536


def digest_index(numel: int):
    """Indices kept in a fixture: everything up to 4096 elements, otherwise
    ~1536 evenly strided samples (odd stride so that it walks across rows)."""
    if numel <= SUBSAMPLE_LIMIT:
        return slice(None)
    stride = -(-numel // SUBSAMPLE_TARGET) | 1
    return slice(0, None, stride)


def digest(t: torch.Tensor):
    """Flatten + subsample (see ``digest_index``); dtype is preserved
    (float32 reference run / float64 truth run)."""
    f = t.detach().reshape(-1).cpu()
    return f[digest_index(f.numel())].numpy().copy()
