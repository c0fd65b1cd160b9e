"""Shared test harness: replays a golden ``Case`` through the oracle (CPU) and
compares key-by-key with the fixtures produced by the real reference."""
from __future__ import annotations

import os

import numpy as np
import torch

from oracle.kfac_oracle import (OracleDPKFAC, specs_from_model, synth_grads,
                                synth_layer_inputs)
from tests.golden.cases import CASES_BY_NAME, digest, layer_input_hw

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_fixture(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + '.npz')))


def case_inputs(case, specs, mods, hw_of, step, dtype=torch.float32):
    """Per-rank LayerInputs + (rank independent) grads of one step, identical
    to what ``make_golden.py`` injected into the reference."""
    per_rank = []
    for rank in range(case.world_size):
        gen = torch.Generator().manual_seed(case.seed + 1000 * step + rank)
        per_rank.append([synth_layer_inputs(spec, case.batch, hw_of.get(mod), case.seq, gen, dtype)
                         for spec, mod in zip(specs, mods)])
    ggen = torch.Generator().manual_seed(case.seed + 1000 * step + 500)
    wg, bg = [], []
    for spec in specs:
        w, b = synth_grads(spec, ggen, dtype)
        wg.append(w); bg.append(b)
    return per_rank, wg, bg


def build_case(case):
    torch.manual_seed(0)
    model = case.build()
    specs, mods = specs_from_model(model)
    hw_of = layer_input_hw(model, case.hw)
    return model, specs, mods, hw_of


def run_oracle_case(case, dtype=torch.float32):
    """Returns a dict with the same keys as the fixture (minus the tag prefix),
    one copy per simulated rank."""
    model, specs, mods, hw_of = build_case(case)
    orc = OracleDPKFAC(specs, variant=case.variant, lr=case.lr, damping=case.damping,
                       fac_update_freq=case.fac_update_freq,
                       kfac_update_freq=case.kfac_update_freq, kl_clip=case.kl_clip,
                       factor_decay=case.factor_decay, world_size=case.world_size)
    out = {}
    for s in range(case.steps):
        lr = case.lr_schedule[s] if case.lr_schedule is not None else None
        per_rank, wg, bg = case_inputs(case, specs, mods, hw_of, s, dtype)
        ow, ob = orc.step(per_rank, wg, bg, lr=lr)
        for rank in range(case.world_size):
            for l, spec in enumerate(specs):
                pre = f'r{rank}_s{s}_l{l}_'
                out[pre + 'wgrad'] = digest(ow[l])
                if ob[l] is not None:
                    out[pre + 'bgrad'] = digest(ob[l])
                out[pre + 'precon'] = digest(orc.m_precon[l])
                if orc.owners[l] == rank:
                    out[pre + 'A'] = digest(orc.m_A[l])
                    out[pre + 'G'] = digest(orc.m_G[l])
                    if case.variant == 'eigen_dp':
                        out[pre + 'dA'] = orc.m_dA[l].numpy().copy()
                        out[pre + 'dG'] = orc.m_dG[l].numpy().copy()
                    else:
                        out[pre + 'invA'] = digest(orc.m_inv_A[l])
                        out[pre + 'invG'] = digest(orc.m_inv_G[l])
    return out


def rel_err(x, ref):
    x = np.asarray(x, dtype=np.float64); ref = np.asarray(ref, dtype=np.float64)
    den = np.linalg.norm(ref)
    return float(np.linalg.norm(x - ref) / den) if den > 0 else float(np.linalg.norm(x))


def noise_by_kind(fx):
    """Worst fp32-vs-fp64 distance of the REFERENCE ITSELF per quantity kind
    (wgrad, bgrad, precon, A, G, dA, dG, invA, invG) over all layers, steps and
    ranks of a fixture: the noise floor any fp32 implementation of this path
    lives on (DESIGN.md 'parity definition')."""
    worst = {}
    for k in fx:
        if k.startswith('ref32_'):
            kind = k.rsplit('_', 1)[-1]
            worst[kind] = max(worst.get(kind, 0.0), rel_err(fx[k], fx['ref64_' + k[6:]]))
    return worst


def tolerance(kind, noise, floor=2e-5, mult=3.0):
    """Parity bar for an fp32 result against the fp64 truth: north_star's 1e-4
    relative, widened only where the reference's own fp32 run is already
    further than 1e-4/3 from the truth (then 3x that noise)."""
    return max(mult * noise.get(kind, 0.0), floor)
