"""CPU oracle for the DP-KFAC ``preconditioner.step()`` hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The shipped package
(``kfac_pytorch_b200``) never imports anything from ``oracle/``.

It is a *restatement* (not a copy) of the reference algorithm in plain
torch-CPU tensor arithmetic over explicit layer specs -- no nn.Module hooks,
no optimizer, no communication backend.  Every function cites the reference
lines it follows (paths relative to the reference repo root,
lzhangbv/kfac_pytorch @ c45acb8).

Pinning: the reference repo ships no golden vectors for this path
(SURVEY.md section 8c), so the oracle is pinned against outputs of the
reference itself, generated in the authoring container by
``tests/golden/make_golden.py`` (which imports the unmodified reference
package) and committed under ``tests/golden/*.npz``.
``tests/test_oracle_golden.py`` checks oracle == reference on every fixture.

All functions are dtype generic: run them in float32 for "the contract" and
in float64 for "the truth" (the reference's own fp32-vs-fp64 noise on this
path is 1e-5..3e-4).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------
# layer description
# --------------------------------------------------------------------------
@dataclass
class LayerSpec:
    """Shape description of one preconditioned layer (nn.Linear / nn.Conv2d).

    Mirrors what the reference reads off the module object in
    kfac/utils.py:85-103,118-140 and kfac/kfac_preconditioner_inv.py:145-154.
    """
    kind: str                       # 'linear' | 'conv2d'
    in_features: int = 0            # linear
    out_features: int = 0           # linear (or out channels for conv)
    has_bias: bool = True
    in_channels: int = 0            # conv
    kernel_size: Tuple[int, int] = (1, 1)
    stride: Tuple[int, int] = (1, 1)
    padding: Tuple[int, int] = (0, 0)

    @property
    def dA(self) -> int:
        if self.kind == 'linear':
            d = self.in_features
        else:
            d = self.in_channels * self.kernel_size[0] * self.kernel_size[1]
        return d + (1 if self.has_bias else 0)

    @property
    def dG(self) -> int:
        return self.out_features

    def weight_shape(self) -> Tuple[int, ...]:
        if self.kind == 'linear':
            return (self.out_features, self.in_features)
        return (self.out_features, self.in_channels,
                self.kernel_size[0], self.kernel_size[1])


def specs_from_model(model, exclude_vocabulary_size=None) -> Tuple[List[LayerSpec], list]:
    """Walk ``model.modules()`` the way kfac/kfac_preconditioner_base.py:132-147
    does (class *name* match on {'Linear','Conv2d'}, optional vocabulary-size
    exclusion) and return the layer specs plus the matched modules."""
    specs, mods = [], []
    for m in model.modules():
        name = m.__class__.__name__
        if name == 'Linear':
            if exclude_vocabulary_size is not None and m.out_features == exclude_vocabulary_size:
                continue
            specs.append(LayerSpec('linear', in_features=m.in_features,
                                   out_features=m.out_features,
                                   has_bias=m.bias is not None))
            mods.append(m)
        elif name == 'Conv2d':
            specs.append(LayerSpec('conv2d', in_channels=m.in_channels,
                                   out_features=m.out_channels,
                                   has_bias=m.bias is not None,
                                   kernel_size=tuple(m.kernel_size),
                                   stride=tuple(m.stride),
                                   padding=tuple(m.padding)))
            mods.append(m)
    return specs, mods


# --------------------------------------------------------------------------
# factor statistics  (kfac/utils.py:33-140)
# --------------------------------------------------------------------------
def conv_patches(x: torch.Tensor, spec: LayerSpec) -> torch.Tensor:
    """im2col rows.  Follows kfac/utils.py:33-54: zero padding (p0 on H, p1 on
    W), stride honoured, dilation/groups ignored, feature order (c, kh, kw).
    Restated with F.unfold, whose channel-major (c, kh, kw) ordering is the
    same ordering ``weight.view(out, -1)`` uses.  Returns (B*oh*ow, C*kh*kw)
    and the spatial size oh*ow."""
    cols = F.unfold(x, kernel_size=spec.kernel_size, padding=spec.padding,
                    stride=spec.stride)               # (B, C*kh*kw, oh*ow)
    spatial = cols.shape[-1]
    rows = cols.transpose(1, 2).reshape(-1, cols.shape[1])
    return rows, spatial


def factor_A(a: torch.Tensor, spec: LayerSpec) -> torch.Tensor:
    """Activation factor of one mini-batch.

    conv2d  (kfac/utils.py:85-94):  rows = patches, ones column appended when
            the layer has a bias, *everything* divided by the spatial size S,
            then  rows^T (rows / B)  ->  [P 1]^T [P 1] / (B S^2).
    linear  (kfac/utils.py:96-103): inputs with >2 dims are averaged over
            dim 1 (the sequence axis); ones column when bias;  a^T (a / B).
    """
    if spec.kind == 'conv2d':
        batch = a.shape[0]
        rows, spatial = conv_patches(a, spec)
        if spec.has_bias:
            rows = torch.cat([rows, rows.new_ones(rows.shape[0], 1)], dim=1)
        rows = rows / spatial
        return rows.t() @ (rows / batch)
    if spec.kind == 'linear':
        if a.dim() > 2:
            a = a.mean(dim=1)
        batch = a.shape[0]
        if spec.has_bias:
            a = torch.cat([a, a.new_ones(a.shape[0], 1)], dim=1)
        return a.t() @ (a / batch)
    raise NotImplementedError(spec.kind)        # kfac/utils.py:82


def factor_G(g: torch.Tensor, spec: LayerSpec, batch_averaged: bool = True) -> torch.Tensor:
    """Output-gradient factor of one mini-batch.

    conv2d  (kfac/utils.py:118-129): NCHW -> (B*S, C) rows; rows *= B when the
            loss was batch-averaged; rows *= S;  rows^T (rows / (B*S)).
    linear  (kfac/utils.py:131-140): >2-dim grads are averaged over dim 1;
            g^T (g * B) when batch-averaged else g^T (g / B).
    """
    if spec.kind == 'conv2d':
        batch = g.shape[0]
        spatial = g.shape[2] * g.shape[3]
        rows = g.permute(0, 2, 3, 1).reshape(-1, g.shape[1])
        if batch_averaged:
            rows = rows * batch
        rows = rows * spatial
        return rows.t() @ (rows / rows.shape[0])
    if spec.kind == 'linear':
        if g.dim() > 2:
            g = g.mean(dim=1)
        batch = g.shape[0]
        if batch_averaged:
            return g.t() @ (g * batch)
        return g.t() @ (g / batch)
    raise NotImplementedError(spec.kind)        # kfac/utils.py:115


def running_average(new: torch.Tensor, current: torch.Tensor, alpha: float) -> torch.Tensor:
    """kfac/utils.py:66-71:  current <- (1-alpha)*current + alpha*new.
    NOTE the weighting: with factor_decay = 0.95 the NEW sample gets 0.95."""
    return current * (1 - alpha) + new * alpha


# --------------------------------------------------------------------------
# inverses  (kfac/utils.py:11-30, kfac_preconditioner_inv_dp.py:98-123,
#            kfac_preconditioner_eigen_dp.py:62-75)
# --------------------------------------------------------------------------
def eig_floored(M: torch.Tensor, eps: float = 1e-10):
    """kfac/utils.py:22-25 + kfac_preconditioner_eigen_dp.py:68-75:
    ascending eigenvalues, eigenvectors in columns, d <- d * (d > eps)."""
    d, Q = torch.linalg.eigh(M)
    d = d * (d > eps).to(d.dtype)
    return d, Q.contiguous()


def pi_damped_inverses(A: torch.Tensor, G: torch.Tensor, damping: float):
    """kfac_preconditioner_inv_dp.py:98-123 (+ inv.py:106-107, utils.py:11-16).

    pi = sqrt((tr A / dA) / (tr G / dG)); sqrt(damping)*pi is added to the
    diagonal of the *stored* running factor A (and sqrt(damping)/pi to G) --
    the damped matrices are returned so the caller can persist them, exactly
    like the in-place ``add_`` of the reference -- then Cholesky inverse."""
    pi = torch.sqrt((A.trace() / A.shape[0]) / (G.trace() / G.shape[0]))
    A_d = A + torch.diag(A.new_full((A.shape[0],), 1.0) * ((damping ** 0.5) * pi))
    G_d = G + torch.diag(G.new_full((G.shape[0],), 1.0) * ((damping ** 0.5) / pi))
    inv_A = torch.cholesky_inverse(torch.linalg.cholesky(A_d)).contiguous()
    inv_G = torch.cholesky_inverse(torch.linalg.cholesky(G_d)).contiguous()
    return A_d, G_d, inv_A, inv_G


# --------------------------------------------------------------------------
# preconditioning  (kfac_preconditioner_inv.py:145-154, eigen_dp.py:78-93,
#                   inv_dp.py:126-138)
# --------------------------------------------------------------------------
def grad_matrix(spec: LayerSpec, wgrad: torch.Tensor, bgrad: Optional[torch.Tensor]) -> torch.Tensor:
    """kfac_preconditioner_inv.py:145-154: [W.view(out,-1) | b]  (dG x dA)."""
    gm = wgrad.reshape(wgrad.shape[0], -1)
    if spec.has_bias:
        gm = torch.cat([gm, bgrad.reshape(-1, 1)], dim=1)
    return gm


def precondition_eigen(gm, QA, dA, QG, dG, damping: float) -> torch.Tensor:
    """kfac_preconditioner_eigen_dp.py:87-89."""
    v1 = QG.t() @ gm @ QA
    v2 = v1 / (dG.unsqueeze(1) * dA.unsqueeze(0) + damping)
    return QG @ v2 @ QA.t()


def precondition_inverse(gm, inv_A, inv_G) -> torch.Tensor:
    """kfac_preconditioner_inv_dp.py:135."""
    return inv_G @ gm @ inv_A


def split_preconditioned(spec: LayerSpec, v: torch.Tensor, wshape, bshape):
    """kfac_preconditioner_inv.py:178-186."""
    if spec.has_bias:
        return v[:, :-1].reshape(wshape), v[:, -1:].reshape(bshape)
    return v.reshape(wshape), None


def kl_clip_scale(v_list, g_list, lr: float, kl_clip: Optional[float]) -> float:
    """kfac_preconditioner_inv.py:190-212: vg_sum = sum_tensors sum(v*g*lr^2)
    accumulated in a Python float from per-tensor ``.sum().item()``;
    nu = min(1, sqrt(kl_clip/|vg_sum|)).  kl_clip None/<=0 disables
    (kfac_preconditioner_base.py:91) -> nu = 1."""
    if kl_clip is None or kl_clip <= 0:
        return 1.0
    vg_sum = 0.0
    for v, g in zip(v_list, g_list):
        vg_sum += (v * g * lr ** 2).sum().item()
    return min(1.0, math.sqrt(kl_clip / abs(vg_sum)))


# --------------------------------------------------------------------------
# whole step, any world size, in one process
# --------------------------------------------------------------------------
@dataclass
class LayerInputs:
    """Per-step tensors of one layer on ONE rank."""
    a: Optional[torch.Tensor]       # layer input  (None on non-owners)
    g: Optional[torch.Tensor]       # grad wrt layer output (None on non-owners)


class OracleDPKFAC:
    """Restatement of ``KFAC.step()`` for the two DP variants
    (kfac_preconditioner_base.py:185-230 driving inv_dp.py / eigen_dp.py),
    simulating ``world_size`` ranks inside one process.

    ``step(per_rank_inputs, wgrads, bgrads)``:
      per_rank_inputs[r][l] -> LayerInputs seen by rank r for layer l (only
      the owner's entry is read, as in inv_dp.py:60-72); wgrads/bgrads are
      the (already rank-averaged) parameter gradients, identical on every
      rank (README.md:51).  Returns the new (weight_grad, bias_grad) lists.
    """

    def __init__(self, specs: Sequence[LayerSpec], variant: str = 'eigen_dp',
                 lr: float = 0.1, damping: float = 0.001, fac_update_freq: int = 1,
                 kfac_update_freq: int = 1, kl_clip: Optional[float] = 0.001,
                 factor_decay: float = 0.95, world_size: int = 1,
                 owners: Optional[Sequence[int]] = None):
        assert variant in ('eigen_dp', 'inverse_dp')
        self.specs = list(specs)
        self.variant = variant
        self.lr, self.damping = lr, damping
        self.fac_update_freq, self.kfac_update_freq = fac_update_freq, kfac_update_freq
        self.kl_clip = kl_clip if (kl_clip is not None and kl_clip > 0) else None
        self.factor_decay = factor_decay
        self.world_size = world_size
        # kfac_preconditioner_inv.py:62-77: round robin, rank_g == rank_a
        self.owners = list(owners) if owners is not None else \
            [i % world_size for i in range(len(self.specs))]
        self.eps = 1e-10                         # kfac_preconditioner_base.py:115
        self.steps = 0
        n = len(self.specs)
        self.m_A: List[Optional[torch.Tensor]] = [None] * n
        self.m_G: List[Optional[torch.Tensor]] = [None] * n
        self.m_QA = [None] * n; self.m_QG = [None] * n
        self.m_dA = [None] * n; self.m_dG = [None] * n
        self.m_inv_A = [None] * n; self.m_inv_G = [None] * n
        self.m_precon = [None] * n
        self.last_nu = 1.0

    # (1) factors -- inv_dp.py:75-90
    def _update_factors(self, per_rank_inputs):
        for l, spec in enumerate(self.specs):
            li = per_rank_inputs[self.owners[l]][l]
            A_new = factor_A(li.a, spec)
            if self.steps == 0:
                self.m_A[l] = torch.eye(A_new.shape[0], dtype=A_new.dtype)
            self.m_A[l] = running_average(A_new, self.m_A[l], self.factor_decay)
            G_new = factor_G(li.g, spec, batch_averaged=True)
            if self.steps == 0:
                self.m_G[l] = torch.eye(G_new.shape[0], dtype=G_new.dtype)
            self.m_G[l] = running_average(G_new, self.m_G[l], self.factor_decay)

    # (2) inverses -- eigen_dp.py:62-75 / inv_dp.py:98-123
    def _update_inverses(self):
        for l in range(len(self.specs)):
            if self.variant == 'eigen_dp':
                self.m_dA[l], self.m_QA[l] = eig_floored(self.m_A[l], self.eps)
                self.m_dG[l], self.m_QG[l] = eig_floored(self.m_G[l], self.eps)
            else:
                (self.m_A[l], self.m_G[l],
                 self.m_inv_A[l], self.m_inv_G[l]) = pi_damped_inverses(
                    self.m_A[l], self.m_G[l], self.damping)

    def step(self, per_rank_inputs, wgrads, bgrads, lr=None, damping=None):
        if lr is not None:
            self.lr = lr
        if damping is not None:
            self.damping = damping
        if self.steps % self.fac_update_freq == 0:
            self._update_factors(per_rank_inputs)
        if self.steps % self.kfac_update_freq == 0:
            self._update_inverses()
        v_all, g_all, out_w, out_b = [], [], [], []
        for l, spec in enumerate(self.specs):
            gm = grad_matrix(spec, wgrads[l], bgrads[l])
            if self.variant == 'eigen_dp':
                v = precondition_eigen(gm, self.m_QA[l], self.m_dA[l],
                                       self.m_QG[l], self.m_dG[l], self.damping)
            else:
                v = precondition_inverse(gm, self.m_inv_A[l], self.m_inv_G[l])
            self.m_precon[l] = v
            vw, vb = split_preconditioned(
                spec, v, wgrads[l].shape, None if bgrads[l] is None else bgrads[l].shape)
            v_all.append(vw); g_all.append(wgrads[l])
            if vb is not None:
                v_all.append(vb); g_all.append(bgrads[l])
            out_w.append(vw); out_b.append(vb)
        nu = kl_clip_scale(v_all, g_all, self.lr, self.kl_clip)
        self.last_nu = nu
        out_w = [w * nu for w in out_w]
        out_b = [None if b is None else b * nu for b in out_b]
        self.steps += 1
        return out_w, out_b


class OracleMPDKFAC:
    """Restatement of the two model-parallel ("MPD") variants -- ``inverse``
    (kfac_preconditioner_inv.py) and ``eigen`` (kfac_preconditioner_eigen.py) -- simulating ``world_size``
    ranks in one process.  Unlike the DP variants every rank keeps its OWN copy of every factor:

      (1) each rank accumulates A/G from its local batch into its copy (identity init + EMA, inv.py:80-91);
      (2) the copies are averaged across ranks in place (all-reduce, inv.py:94-103);
      (3) ``inverse``: pi from the averaged factors, then ONLY the owner adds sqrt(damping)*pi to ITS copy (the
          in-place add of inv.py:122-128 runs under the rank guard, so the replicas drift apart until the next
          average) and inverts; ``eigen``: the owner of each factor (rank_a for A, rank_g for G) runs eigh + floor
          (eigen.py:98-119);
      (4) ``eigen`` broadcasts (Q, d) and every rank preconditions (eigen.py:122-144); ``inverse`` either
          broadcasts the inverses (communicate_inverse_or_not=True) or lets the owner precondition and broadcasts
          the result (default, inv.py:156-175) -- all ranks end with the owner's numbers either way;
      (5) KL clip as in the DP variants (inv.py:188-217).
    """

    def __init__(self, specs: Sequence[LayerSpec], variant: str = 'eigen', lr: float = 0.1,
                 damping: float = 0.001, fac_update_freq: int = 1, kfac_update_freq: int = 1,
                 kl_clip: Optional[float] = 0.001, factor_decay: float = 0.95, world_size: int = 1,
                 distribute_layer_factors: Optional[bool] = None):
        assert variant in ('eigen', 'inverse')
        self.specs = list(specs)
        self.variant = variant
        self.lr, self.damping = lr, damping
        self.fac_update_freq, self.kfac_update_freq = fac_update_freq, kfac_update_freq
        self.kl_clip = kl_clip if (kl_clip is not None and kl_clip > 0) else None
        self.factor_decay = factor_decay
        self.world_size = world_size
        if distribute_layer_factors is None:                       # eigen.py:66-71
            distribute_layer_factors = variant == 'eigen' and world_size > len(self.specs)
        self.owner_a, self.owner_g = [], []
        it = 0
        for _ in self.specs:                                       # eigen.py:75-94 / inv.py:62-77
            ra = it % world_size
            if distribute_layer_factors and variant == 'eigen':
                it += 1
                rg = it % world_size
            else:
                rg = ra
            self.owner_a.append(ra); self.owner_g.append(rg)
            it += 1
        self.owners = self.owner_a
        self.eps = 1e-10
        self.steps = 0
        n = len(self.specs)
        self.m_A = [[None] * n for _ in range(world_size)]          # [rank][layer]
        self.m_G = [[None] * n for _ in range(world_size)]
        self.m_QA = [None] * n; self.m_QG = [None] * n
        self.m_dA = [None] * n; self.m_dG = [None] * n
        self.m_inv_A = [None] * n; self.m_inv_G = [None] * n
        self.m_precon = [None] * n
        self.last_nu = 1.0

    def step(self, per_rank_inputs, wgrads, bgrads, lr=None, damping=None):
        if lr is not None:
            self.lr = lr
        if damping is not None:
            self.damping = damping
        P = self.world_size
        if self.steps % self.fac_update_freq == 0:
            for l, spec in enumerate(self.specs):
                for r in range(P):
                    li = per_rank_inputs[r][l]
                    A_new = factor_A(li.a, spec)
                    G_new = factor_G(li.g, spec, batch_averaged=True)
                    if self.steps == 0:
                        self.m_A[r][l] = torch.eye(A_new.shape[0], dtype=A_new.dtype)
                        self.m_G[r][l] = torch.eye(G_new.shape[0], dtype=G_new.dtype)
                    self.m_A[r][l] = running_average(A_new, self.m_A[r][l], self.factor_decay)
                    self.m_G[r][l] = running_average(G_new, self.m_G[r][l], self.factor_decay)
                if P > 1:                                           # sum over ranks, then divide (backend.py:136-139)
                    A_avg = sum(self.m_A[r][l] for r in range(P)) / P
                    G_avg = sum(self.m_G[r][l] for r in range(P)) / P
                    for r in range(P):
                        self.m_A[r][l] = A_avg.clone()
                        self.m_G[r][l] = G_avg.clone()
        if self.steps % self.kfac_update_freq == 0:
            for l in range(len(self.specs)):
                ra, rg = self.owner_a[l], self.owner_g[l]
                if self.variant == 'eigen':
                    self.m_dA[l], self.m_QA[l] = eig_floored(self.m_A[ra][l], self.eps)
                    self.m_dG[l], self.m_QG[l] = eig_floored(self.m_G[rg][l], self.eps)
                else:
                    A, G = self.m_A[ra][l], self.m_G[ra][l]
                    (self.m_A[ra][l], self.m_G[ra][l],
                     self.m_inv_A[l], self.m_inv_G[l]) = pi_damped_inverses(A, G, self.damping)
        v_all, g_all, out_w, out_b = [], [], [], []
        for l, spec in enumerate(self.specs):
            gm = grad_matrix(spec, wgrads[l], bgrads[l])
            if self.variant == 'eigen':
                v = precondition_eigen(gm, self.m_QA[l], self.m_dA[l], self.m_QG[l], self.m_dG[l], self.damping)
            else:
                v = precondition_inverse(gm, self.m_inv_A[l], self.m_inv_G[l])
            self.m_precon[l] = v
            vw, vb = split_preconditioned(
                spec, v, wgrads[l].shape, None if bgrads[l] is None else bgrads[l].shape)
            v_all.append(vw); g_all.append(wgrads[l])
            if vb is not None:
                v_all.append(vb); g_all.append(bgrads[l])
            out_w.append(vw); out_b.append(vb)
        nu = kl_clip_scale(v_all, g_all, self.lr, self.kl_clip)
        self.last_nu = nu
        out_w = [w * nu for w in out_w]
        out_b = [None if b is None else b * nu for b in out_b]
        self.steps += 1
        return out_w, out_b


# --------------------------------------------------------------------------
# seeded synthetic inputs shared by the golden generator, the tests and bench
# --------------------------------------------------------------------------
def conv_out_hw(spec: LayerSpec, h: int, w: int) -> Tuple[int, int]:
    oh = (h + 2 * spec.padding[0] - spec.kernel_size[0]) // spec.stride[0] + 1
    ow = (w + 2 * spec.padding[1] - spec.kernel_size[1]) // spec.stride[1] + 1
    return oh, ow


def synth_layer_inputs(spec: LayerSpec, batch: int, hw: Optional[Tuple[int, int]],
                       seq: Optional[int], gen: torch.Generator,
                       dtype=torch.float32) -> LayerInputs:
    """a ~ N(0,1) with the layer's input shape, g ~ N(0,1)/batch with its
    output shape (BASELINE.md section 2 'injection mode').  Values are always
    DRAWN in float32 (so the fp32 and fp64 runs see the same numbers) and then
    cast to ``dtype``."""
    f32 = torch.float32
    if spec.kind == 'conv2d':
        h, w = hw
        oh, ow = conv_out_hw(spec, h, w)
        a = torch.randn(batch, spec.in_channels, h, w, generator=gen, dtype=f32)
        g = torch.randn(batch, spec.out_features, oh, ow, generator=gen, dtype=f32) / batch
    else:
        if seq is None:
            a = torch.randn(batch, spec.in_features, generator=gen, dtype=f32)
            g = torch.randn(batch, spec.out_features, generator=gen, dtype=f32) / batch
        else:
            a = torch.randn(batch, seq, spec.in_features, generator=gen, dtype=f32)
            g = torch.randn(batch, seq, spec.out_features, generator=gen, dtype=f32) / batch
    return LayerInputs(a.to(dtype), g.to(dtype))


def synth_grads(spec: LayerSpec, gen: torch.Generator, dtype=torch.float32):
    w = torch.randn(*spec.weight_shape(), generator=gen, dtype=torch.float32).to(dtype)
    b = (torch.randn(spec.out_features, generator=gen, dtype=torch.float32).to(dtype)
         if spec.has_bias else None)
    return w, b
