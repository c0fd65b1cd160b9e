"""Shared test harness: replays a golden ``Case`` through the oracle (CPU) and
compares key-by-key with the fixtures produced by the real reference."""
from __future__ import annotations

import os

import numpy as np
import torch

from oracle.kfac_oracle import (OracleDPKFAC, OracleMPDKFAC, specs_from_model, synth_grads,
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
    mpd = not case.variant.endswith('_dp')
    cls = OracleMPDKFAC if mpd else OracleDPKFAC
    orc = cls(specs, variant=case.variant, lr=case.lr, damping=case.damping,
              fac_update_freq=case.fac_update_freq,
              kfac_update_freq=case.kfac_update_freq, kl_clip=case.kl_clip,
              factor_decay=case.factor_decay, world_size=case.world_size,
              **{k: v for k, v in case.extra.items() if k == 'distribute_layer_factors'})
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
                owner = orc.owners[l] == rank
                if mpd or owner:
                    out[pre + 'A'] = digest(orc.m_A[rank][l] if mpd else orc.m_A[l])
                    out[pre + 'G'] = digest(orc.m_G[rank][l] if mpd else orc.m_G[l])
                if case.variant.startswith('eigen') and (mpd or owner):
                    out[pre + 'dA'] = orc.m_dA[l].numpy().copy()
                    out[pre + 'dG'] = orc.m_dG[l].numpy().copy()
                if case.variant.startswith('inverse') and (owner or case.extra.get('communicate_inverse_or_not')):
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


def tolerance(kind, noise, floor=1e-4, mult=3.0):
    """Parity bar for an fp32 result against the fp64 truth: north_star's 1e-4
    relative (``floor``), widened only where the reference's own fp32 run is
    already further than floor/3 from the truth (then 3x that noise)."""
    return max(mult * noise.get(kind, 0.0), floor)


# ---------------------------------------------------------------------------
# replay of a golden case through the PRODUCT package (kfac_pytorch_b200)
# ---------------------------------------------------------------------------
def run_package_case(case, device, rank=0, engine=None, partition='round_robin'):
    """Drives ``kfac.get_kfac_module(variant)`` exactly the way
    ``make_golden.py`` drove the reference (injection mode) and returns the
    fixture-keyed digests of this rank."""
    import kfac_pytorch_b200 as kfac
    from kfac_pytorch_b200 import backend
    model, specs, mods, hw_of = build_case(case)
    model.to(device)
    KFAC = kfac.get_kfac_module(case.variant)
    pre = KFAC(model, lr=case.lr, damping=case.damping, fac_update_freq=case.fac_update_freq,
               kfac_update_freq=case.kfac_update_freq, kl_clip=case.kl_clip,
               factor_decay=case.factor_decay, hook_enabled=False, engine=engine,
               partition=partition, **case.extra)
    assert pre.modules == mods
    mpd = not case.variant.endswith('_dp')
    if pre.module_ranks is None:
        pre.schedule_module_ranks()
    out = {}
    for s in range(case.steps):
        if case.lr_schedule is not None:
            pre.param_groups[0]['lr'] = case.lr_schedule[s]
        per_rank, wg, bg = case_inputs(case, specs, mods, hw_of, s)
        for l, mod in enumerate(mods):
            if pre.steps % pre.param_groups[0]['fac_update_freq'] == 0:
                ra, rg = pre.module_ranks[mod]
                if mpd or rank == ra:
                    pre.m_a[mod] = per_rank[rank][l].a.to(device)
                if mpd or rank == rg:
                    pre.m_g[mod] = per_rank[rank][l].g.to(device)
            mod.weight.grad = wg[l].to(device)
            if bg[l] is not None:
                mod.bias.grad = bg[l].to(device)
        pre.step()
        for l, mod in enumerate(mods):
            p = f'r{rank}_s{s}_l{l}_'
            out[p + 'wgrad'] = digest(mod.weight.grad)
            if mod.bias is not None:
                out[p + 'bgrad'] = digest(mod.bias.grad)
            out[p + 'precon'] = digest(pre.m_precon_grad[mod])
            owner = pre.module_ranks[mod][0] == rank
            if mpd or owner:
                out[p + 'A'] = digest(pre.m_A[mod])
                out[p + 'G'] = digest(pre.m_G[mod])
            if case.variant.startswith('eigen') and (mpd or owner):
                out[p + 'dA'] = pre.m_dA[mod].cpu().numpy().copy()
                out[p + 'dG'] = pre.m_dG[mod].cpu().numpy().copy()
            if case.variant.startswith('inverse') and (owner or case.extra.get('communicate_inverse_or_not')):
                out[p + 'invA'] = digest(pre.m_inv_A[mod])
                out[p + 'invG'] = digest(pre.m_inv_G[mod])
    return out, pre


def compare_with_fixture(case, out, mult=3.0, floor=1e-4):
    """Every produced quantity must sit within the noise-calibrated tolerance of
    the fp64 truth recorded from the reference.  Returns the worst ratio
    err/tolerance per kind for reporting."""
    fx = load_fixture(case.name)
    noise = noise_by_kind(fx)
    report = {}
    for k, v in out.items():
        kind = k.rsplit('_', 1)[-1]
        ref64 = fx['ref64_' + k]
        assert ref64.shape == v.shape, (k, ref64.shape, v.shape)
        err = rel_err(v, ref64)
        tol = tolerance(kind, noise, floor=floor, mult=mult)
        assert err <= tol, f'{case.name}:{k}: rel err {err:.3e} > tol {tol:.3e} (reference fp32 noise {noise.get(kind, 0):.3e})'
        report[kind] = max(report.get(kind, 0.0), err)
    return report
