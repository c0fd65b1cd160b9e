"""TEST-ONLY engine: implements the ``NativeEngine`` interface with the CPU
oracle so that the host orchestration of ``kfac_pytorch_b200.preconditioner``
(gating, ownership, slab exchange, scheduler interplay) can be exercised
without a GPU (``-m "not gpu"`` tests, gloo world_size 2).  Never imported by
the package."""
from __future__ import annotations

import torch

from oracle import kfac_oracle as orc
from kfac_pytorch_b200.native import KFACB_CONV2D


def spec_of(d) -> orc.LayerSpec:
    if d.kind == KFACB_CONV2D:
        return orc.LayerSpec('conv2d', in_channels=d.in_c, out_features=d.out_c,
                             has_bias=bool(d.has_bias), kernel_size=(d.kh, d.kw),
                             stride=(d.sh, d.sw), padding=(d.ph, d.pw))
    return orc.LayerSpec('linear', in_features=d.in_c, out_features=d.out_c,
                         has_bias=bool(d.has_bias))


class OracleEngine:
    def factor_update(self, layers, a, g, A, G, alpha, first_step):
        for d, ai, gi, Ai, Gi in zip(layers, a, g, A, G):
            s = spec_of(d)
            if ai is not None:
                cur = torch.eye(Ai.shape[0], dtype=Ai.dtype) if first_step else Ai
                Ai.copy_(orc.running_average(orc.factor_A(ai, s), cur, alpha))
            if gi is not None:
                cur = torch.eye(Gi.shape[0], dtype=Gi.dtype) if first_step else Gi
                Gi.copy_(orc.running_average(orc.factor_G(gi, s, True), cur, alpha))

    def eigh(self, mats, Qs, ds, eps_floor, sweeps_out=None):
        for M, Q, d in zip(mats, Qs, ds):
            dd, QQ = orc.eig_floored(M, eps_floor)
            Q.copy_(QQ); d.copy_(dd)

    def pi_damped_inverse(self, A, G, invA, invG, damping):
        for Ai, Gi, iA, iG in zip(A, G, invA, invG):
            if Ai is None:
                continue
            Ad, Gd, a_inv, g_inv = orc.pi_damped_inverses(Ai, Gi, damping)
            Ai.copy_(Ad); Gi.copy_(Gd); iA.copy_(a_inv); iG.copy_(g_inv)

    @staticmethod
    def _gm(d, w, b):
        if w.dim() == 1:      # fused gradient-reduce path: flat views [W | b] of the reduced send slab
            w = w.reshape(d.dG, -1)
            b = None if b is None else b.reshape(-1)
        return orc.grad_matrix(spec_of(d), w, b)

    def precond_eigen(self, layers, QA, QG, dA, dG, wgrad, bgrad, out, damping, peer_deltas=None):
        for d, qa, qg, da, dg, w, b, o in zip(layers, QA, QG, dA, dG, wgrad, bgrad, out):
            gm = self._gm(d, w, b)
            o.copy_(orc.precondition_eigen(gm, qa, da, qg, dg, damping))

    def precond_inverse(self, layers, invA, invG, wgrad, bgrad, out, peer_deltas=None):
        for d, ia, ig, w, b, o in zip(layers, invA, invG, wgrad, bgrad, out):
            gm = self._gm(d, w, b)
            o.copy_(orc.precondition_inverse(gm, ia, ig))

    def finalize(self, layers, precon, wgrad, bgrad, lr, kl_clip, force_nu_one, nu_out):
        vs, gs, splits = [], [], []
        for d, v, w, b in zip(layers, precon, wgrad, bgrad):
            vw, vb = orc.split_preconditioned(spec_of(d), v, w.shape, None if b is None else b.shape)
            splits.append((vw, vb))
            vs.append(vw); gs.append(w)
            if vb is not None:
                vs.append(vb); gs.append(b)
        nu = 1.0 if force_nu_one else orc.kl_clip_scale(vs, gs, lr, kl_clip)
        nu_out.fill_(nu)
        for (vw, vb), w, b in zip(splits, wgrad, bgrad):
            w.copy_(vw * nu)
            if b is not None:
                b.copy_(vb * nu)

    # ---- fused gradient-reduce path (kfacb_pack_grads / kfacb_finalize_dot / kfacb_finalize_write) ----
    def pack_grads(self, layers, wgrad, bgrad, dst):
        for d, w, b, o in zip(layers, wgrad, bgrad, dst):
            nw = w.numel()
            o[:nw].copy_(w.reshape(-1))
            if b is not None:
                o[nw:].copy_(b.reshape(-1))

    def finalize_dot(self, layers, precon, wgrad, bgrad, vg_sum_out):
        tot = 0.0
        for d, v, w, b in zip(layers, precon, wgrad, bgrad):
            g = torch.cat([w.reshape(v.shape[0], -1), b.reshape(-1, 1)], 1) if b is not None else w.reshape(v.shape)
            tot += float((v.double() * g.double()).sum())
        vg_sum_out.fill_(tot)

    def finalize_write(self, layers, precon, wgrad, bgrad, vg_sum, lr, kl_clip, force_nu_one, nu_out):
        nu = 1.0
        if kl_clip and kl_clip > 0 and not force_nu_one and vg_sum is not None:
            nu = min(1.0, (kl_clip / abs(float(vg_sum[0]) * lr * lr)) ** 0.5)
        nu_out.fill_(nu)
        for d, v, w, b in zip(layers, precon, wgrad, bgrad):
            vw, vb = orc.split_preconditioned(spec_of(d), v, w.shape, None if b is None else b.shape)
            w.copy_(vw * nu)
            if b is not None:
                b.copy_(vb * nu)
