"""Host side of the B200-native DP-KFAC preconditioner.

Mirrors the reference's Python surface for the distributed-preconditioning
variants (``eigen_dp`` / ``inverse_dp``):

  reference                                           here
  --------------------------------------------------  ---------------------------
  kfac/kfac_preconditioner_base.py:13-230  KFAC       KFAC (orchestration, hooks)
  kfac/kfac_preconditioner_inv.py:62-77,145-217       schedule_module_ranks, _communicate_pred,
                                                      _update_grad_in_place
  kfac/kfac_preconditioner_inv_dp.py:60-138           owner-only hooks / factors / inverse / pred
  kfac/kfac_preconditioner_eigen_dp.py:62-93          eigen variant
  kfac/kfac_preconditioner_base.py:233-301            KFACParamScheduler (scheduler.py)

All tensor math is delegated to ``NativeEngine`` (ctypes -> libkfacb.so, CUDA
sm_100a).  Nothing here falls back to PyTorch math: without the library or a
GPU, ``step()`` raises.  The per-layer Python loops of the reference collapse
into one batched native call per phase, the per-tensor ``.item()`` syncs of the
KL-clip (kfac_preconditioner_inv.py:198,200) are gone (nu stays on the device),
and the L per-layer broadcasts become one broadcast per owner rank over a
contiguous slab.
"""
from __future__ import annotations

import logging
from typing import Dict, List, Optional

import torch
import torch.optim as optim

from . import backend
from .native import KFACB_CONV2D, KFACB_LINEAR, LayerDesc

logger = logging.getLogger()


def _static_desc(module) -> LayerDesc:
    """Descriptor with the batch-independent fields (enough for dA/dG)."""
    d = LayerDesc()
    if module.__class__.__name__ == 'Conv2d':
        d.kind = KFACB_CONV2D
        d.in_c, d.out_c = module.in_channels, module.out_channels
        d.kh, d.kw = module.kernel_size
        d.sh, d.sw = module.stride
        d.ph, d.pw = module.padding if not isinstance(module.padding, str) else (0, 0)
        d.in_h, d.in_w = d.kh, d.kw
        d.out_h, d.out_w = 1, 1
    else:
        d.kind = KFACB_LINEAR
        d.in_c, d.out_c = module.in_features, module.out_features
        d.kh = d.kw = d.sh = d.sw = 1
        d.in_h = d.in_w = d.out_h = d.out_w = 1
    d.has_bias = 1 if module.bias is not None else 0
    d.batch, d.seq = 1, 1
    return d


def _full_desc(module, a: Optional[torch.Tensor], g: Optional[torch.Tensor]) -> LayerDesc:
    """Descriptor for factor accumulation, from the captured tensors' shapes
    (what ComputeA/ComputeG read off ``a``, ``g`` and ``layer``; kfac/utils.py:85-140)."""
    d = _static_desc(module)
    if d.kind == KFACB_CONV2D:
        if isinstance(module.padding, str):
            raise NotImplementedError('KFAC does not support string padding on Conv2d')
        if a is not None:
            if a.dim() != 4:
                raise ValueError('Conv2d input must be (B,C,H,W)')
            d.batch, d.in_h, d.in_w = a.shape[0], a.shape[2], a.shape[3]
            d.out_h = (d.in_h + 2 * d.ph - d.kh) // d.sh + 1
            d.out_w = (d.in_w + 2 * d.pw - d.kw) // d.sw + 1
            if g is not None and tuple(g.shape) != (d.batch, d.out_c, d.out_h, d.out_w):
                raise ValueError('Conv2d grad_output shape %s does not match geometry' % (tuple(g.shape),))
        elif g is not None:
            d.batch, d.out_h, d.out_w = g.shape[0], g.shape[2], g.shape[3]
            d.in_h = (d.out_h - 1) * d.sh + d.kh - 2 * d.ph
            d.in_w = (d.out_w - 1) * d.sw + d.kw - 2 * d.pw
    else:
        t = a if a is not None else g
        if t.dim() == 2:
            d.batch, d.seq = t.shape[0], 1
        elif t.dim() == 3:
            d.batch, d.seq = t.shape[0], t.shape[1]       # mean over dim 1 (kfac/utils.py:98-99)
        else:
            raise ValueError('Linear inputs must be (B,D) or (B,T,D)')
    return d


def _prep(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


class KFAC(optim.Optimizer):
    """Distributed-preconditioning K-FAC gradient preconditioner (B200 native).

    Same constructor, ``param_groups[0]`` keys, hooks, public state dicts and
    ``step(closure=None, epoch=None)`` contract as the reference classes
    (kfac/kfac_preconditioner_eigen_dp.py:35-58, kfac_preconditioner_inv_dp.py:34-57).
    ``variant`` is fixed by the subclass.  ``partition`` ('round_robin' =
    reference default, or 'greedy' = longest-processing-time on dA^3+dG^3) is
    keyword-only and not part of the reference signature.
    """
    variant = None   # 'eigen_dp' | 'inverse_dp'

    def __init__(self, model, lr=0.1, damping=0.001, fac_update_freq=1, kfac_update_freq=1,
                 kl_clip=0.001, factor_decay=0.95, exclude_vocabulary_size=None,
                 hook_enabled=True, exclude_parts='', *, partition='round_robin', engine=None):
        defaults = dict(lr=lr, damping=damping, fac_update_freq=fac_update_freq,
                        kfac_update_freq=kfac_update_freq)
        super().__init__(model.parameters(), defaults)
        if self.variant not in ('eigen_dp', 'inverse_dp'):
            raise NotImplementedError('use KFAC_EIGEN_DP / KFAC_INV_DP (or kfac.get_kfac_module)')
        backend.ensure_initialized()
        self.fac_update_freq = fac_update_freq
        self.kfac_update_freq = kfac_update_freq
        self.communicate_inverse_or_not = False          # DP variants always exchange preconditioned grads
        self.kl_clip = kl_clip if (kl_clip is not None and kl_clip > 0) else None
        self.factor_decay = factor_decay
        self.exclude_vocabulary_size = exclude_vocabulary_size
        self.hook_enabled = hook_enabled
        self.exclude_communicate_inverse = exclude_parts.find('CommunicateInverse') >= 0
        self.exclude_compute_inverse = exclude_parts.find('ComputeInverse') >= 0
        self.exclude_communicate_factor = exclude_parts.find('CommunicateFactor') >= 0
        self.exclude_compute_factor = exclude_parts.find('ComputeFactor') >= 0
        self.partition = partition

        self.modules: List[torch.nn.Module] = []
        self.module_names: List[str] = []
        self._register_module_hooks(model)

        self.m_a, self.m_g = {}, {}
        self.m_A, self.m_G = {}, {}
        self.m_inv_A, self.m_inv_G = {}, {}
        self.m_QA, self.m_QG = {}, {}
        self.m_dA, self.m_dG = {}, {}
        self.m_precon_grad = {}
        self.module_ranks = None
        self.eps = 1e-10
        self.steps = 0

        self._engine = engine
        self._static = [_static_desc(m) for m in self.modules]
        self._slab = None
        self._slab_segments = None
        self._nu = None
        self.schedule_module_ranks()       # DP variants schedule eagerly (inv_dp.py:57)

    # ------------------------------------------------------------------ hooks
    def set_hook_enabled(self, mode=True):
        self.hook_enabled = mode

    def _forward_hook_event(self, module, input):
        """Owner-only capture of the layer input (kfac_preconditioner_inv_dp.py:60-65)."""
        if self.hook_enabled and torch.is_grad_enabled() and self.steps % self.fac_update_freq == 0:
            rank_a, _ = self.module_ranks[module]
            if backend.comm.rank() == rank_a:
                self.m_a[module] = input[0].data

    def _backward_hook_event(self, module, grad_input, grad_output):
        """Owner-only capture of the output gradient (kfac_preconditioner_inv_dp.py:67-72)."""
        if self.hook_enabled and self.steps % self.fac_update_freq == 0:
            _, rank_g = self.module_ranks[module]
            if backend.comm.rank() == rank_g:
                self.m_g[module] = grad_output[0].data

    def _register_module_hooks(self, model):
        """Class-NAME match on {'Linear','Conv2d'} in model.modules() order
        (kfac_preconditioner_base.py:132-149)."""
        supported = {'Linear', 'Conv2d'}
        name_idx = 0
        for module in model.modules():
            classname = module.__class__.__name__
            if classname in supported:
                if (self.exclude_vocabulary_size is not None and classname == 'Linear'
                        and module.out_features == self.exclude_vocabulary_size):
                    continue
                self.modules.append(module)
                module.register_forward_pre_hook(self._forward_hook_event)
                module.register_full_backward_hook(self._backward_hook_event)
                self.module_names.append('module_name_%s_%d' % (classname, name_idx))
                name_idx += 1
        if backend.comm.rank() == 0:
            logger.info("#register modules: %s", len(self.modules))

    # ------------------------------------------------------------- scheduling
    def schedule_module_ranks(self):
        """Round robin with rank_g == rank_a (kfac_preconditioner_inv.py:62-77),
        or greedy LPT when ``partition='greedy'``."""
        if self.module_ranks is not None:
            return self.module_ranks
        size = backend.comm.size()
        ranks = {}
        if self.partition == 'greedy' and size > 1:
            load = [0.0] * size
            order = sorted(range(len(self.modules)),
                           key=lambda i: -(float(self._static[i].dA) ** 3 + float(self._static[i].dG) ** 3))
            for i in order:
                r = min(range(size), key=lambda k: (load[k], k))
                load[r] += float(self._static[i].dA) ** 3 + float(self._static[i].dG) ** 3
                ranks[self.modules[i]] = (r, r)
            ranks = {m: ranks[m] for m in self.modules}
        elif self.partition in ('round_robin', 'greedy'):
            for i, module in enumerate(self.modules):
                ranks[module] = (i % size, i % size)
        else:
            raise ValueError('unknown partition %r' % (self.partition,))
        self.module_ranks = ranks
        if backend.comm.rank() == 0:
            logger.info('module_ranks: %s', ranks.values())
        return ranks

    def _owned(self) -> List[int]:
        me = backend.comm.rank()
        return [i for i, m in enumerate(self.modules) if self.module_ranks[m][0] == me]

    # ----------------------------------------------------------------- engine
    def _get_engine(self):
        if self._engine is None:
            from .native import NativeEngine
            dev = self.modules[0].weight.device if self.modules else None
            if dev is None or dev.type != 'cuda':
                raise RuntimeError('kfac_pytorch_b200: the model must live on a CUDA (sm_100a) device; '
                                   'preconditioner.step() has no CPU fallback')
            self._engine = NativeEngine(dev)
        return self._engine

    def _device(self):
        return self.modules[0].weight.device

    # ---------------------------------------------------------------- phase 1
    def _compute_factors(self):
        """A/G accumulation + running average on the owner only, from the owner's
        local mini-batch (kfac_preconditioner_inv_dp.py:75-90)."""
        eng = self._get_engine()
        owned = self._owned()
        if not owned:
            return
        descs, a_l, g_l, A_l, G_l = [], [], [], [], []
        for i in owned:
            m = self.modules[i]
            a, g = _prep(self.m_a[m]), _prep(self.m_g[m])
            d = _full_desc(m, a, g)
            if m not in self.m_A:
                self.m_A[m] = torch.empty(d.dA, d.dA, dtype=torch.float32, device=a.device)
                self.m_G[m] = torch.empty(d.dG, d.dG, dtype=torch.float32, device=a.device)
            descs.append(d); a_l.append(a); g_l.append(g)
            A_l.append(self.m_A[m]); G_l.append(self.m_G[m])
        eng.factor_update(descs, a_l, g_l, A_l, G_l, self.factor_decay, self.steps == 0)

    def _communicate_factors(self):
        """No factor communication in the DP variants (kfac_preconditioner_inv_dp.py:93-95)."""
        pass

    # ---------------------------------------------------------------- phase 2
    def _compute_inverse(self):
        eng = self._get_engine()
        owned = self._owned()
        if not owned:
            return
        if self.variant == 'eigen_dp':
            # eigh + floor on the owner (kfac_preconditioner_eigen_dp.py:62-75)
            mats, Qs, ds = [], [], []
            for i in owned:
                m = self.modules[i]
                A, G = self.m_A[m], self.m_G[m]
                if m not in self.m_QA:
                    self.m_QA[m] = torch.empty_like(A); self.m_dA[m] = A.new_empty(A.shape[0])
                    self.m_QG[m] = torch.empty_like(G); self.m_dG[m] = G.new_empty(G.shape[0])
                mats += [A, G]; Qs += [self.m_QA[m], self.m_QG[m]]; ds += [self.m_dA[m], self.m_dG[m]]
            eng.eigh(mats, Qs, ds, self.eps)
        else:
            # pi-damped (in place!) Cholesky-equivalent inverse (kfac_preconditioner_inv_dp.py:98-123)
            A_l, G_l, iA, iG = [], [], [], []
            for i in owned:
                m = self.modules[i]
                A, G = self.m_A[m], self.m_G[m]
                if m not in self.m_inv_A:
                    self.m_inv_A[m] = torch.empty_like(A)
                    self.m_inv_G[m] = torch.empty_like(G)
                A_l.append(A); G_l.append(G); iA.append(self.m_inv_A[m]); iG.append(self.m_inv_G[m])
            eng.pi_damped_inverse(A_l, G_l, iA, iG, self.damping)

    # ---------------------------------------------------------------- phase 3
    def _ensure_slab(self):
        """One flat buffer for all preconditioned gradients, laid out so that each
        owner rank's layers are contiguous: the exchange is one broadcast per rank."""
        if self._slab is not None:
            return
        dev = self._device()
        size = backend.comm.size()
        order = sorted(range(len(self.modules)),
                       key=lambda i: (self.module_ranks[self.modules[i]][0], i))
        offs, seg, off = {}, [[None, None] for _ in range(size)], 0
        for i in order:
            r = self.module_ranks[self.modules[i]][0]
            n = self._static[i].dA * self._static[i].dG
            if seg[r][0] is None:
                seg[r][0] = off
            offs[i] = off
            off += n
            seg[r][1] = off
        self._slab = torch.zeros(max(off, 1), dtype=torch.float32, device=dev)
        self._slab_segments = seg
        for i, m in enumerate(self.modules):
            d = self._static[i]
            self.m_precon_grad[m] = self._slab[offs[i]:offs[i] + d.dA * d.dG].view(d.dG, d.dA)
        self._nu = torch.ones(1, dtype=torch.float32, device=dev)

    def _grads(self, idx):
        w, b = [], []
        for i in idx:
            m = self.modules[i]
            if m.weight.grad is None or (m.bias is not None and m.bias.grad is None):
                raise RuntimeError('%s has no .grad: call loss.backward() before preconditioner.step()'
                                   % self.module_names[i])
            wg = m.weight.grad.data
            bg = m.bias.grad.data if m.bias is not None else None
            if wg.dtype != torch.float32 or not wg.is_contiguous() or \
                    (bg is not None and (bg.dtype != torch.float32 or not bg.is_contiguous())):
                raise RuntimeError('%s: gradients must be contiguous float32' % self.module_names[i])
            w.append(wg); b.append(bg)
        return w, b

    def _compute_pred(self):
        """Owner-only two-sided product (eigen_dp.py:78-93 / inv_dp.py:126-138)."""
        assert not self.communicate_inverse_or_not
        eng = self._get_engine()
        self._ensure_slab()
        owned = self._owned()
        if not owned:
            return
        descs = [self._static[i] for i in owned]
        w, b = self._grads(owned)
        out = [self.m_precon_grad[self.modules[i]] for i in owned]
        mods = [self.modules[i] for i in owned]
        if self.variant == 'eigen_dp':
            eng.precond_eigen(descs, [self.m_QA[m] for m in mods], [self.m_QG[m] for m in mods],
                              [self.m_dA[m] for m in mods], [self.m_dG[m] for m in mods],
                              w, b, out, self.damping)
        else:
            eng.precond_inverse(descs, [self.m_inv_A[m] for m in mods],
                                [self.m_inv_G[m] for m in mods], w, b, out)

    def _communicate_pred(self):
        """Owner -> everyone exchange of the preconditioned gradients.  The reference
        issues one broadcast per layer (kfac_preconditioner_inv.py:164-175); the slab
        layout makes it one broadcast per owner rank."""
        handles = []
        for r, (beg, end) in enumerate(self._slab_segments):
            if beg is None:
                continue
            handles.append(backend.comm.broadcast_async_(self._slab[beg:end], r))
        for h in handles:
            backend.comm.synchronize(h)

    # ---------------------------------------------------------------- phase 4
    def _update_grad_in_place(self):
        """KL clip + rewrite of weight/bias grads (kfac_preconditioner_inv.py:188-217),
        with nu computed and applied on the device."""
        eng = self._get_engine()
        idx = list(range(len(self.modules)))
        w, b = self._grads(idx)
        eng.finalize(self._static, [self.m_precon_grad[m] for m in self.modules], w, b,
                     float(self.lr), self.kl_clip,
                     self.kl_clip is not None and self.exclude_communicate_inverse, self._nu)

    # ------------------------------------------------------------------- step
    @torch.no_grad()
    def step(self, closure=None, epoch=None):
        """One K-FAC step; same phase order and gating as
        kfac_preconditioner_base.py:185-230."""
        group = self.param_groups[0]
        self.lr = group['lr']
        self.damping = group['damping']
        self.fac_update_freq = group['fac_update_freq']
        self.kfac_update_freq = group['kfac_update_freq']

        if self.module_ranks is None:
            self.schedule_module_ranks()

        if self.steps % self.fac_update_freq == 0:
            if not self.exclude_compute_factor:
                self._compute_factors()
            if not self.exclude_communicate_factor:
                if backend.comm.size() > 1:
                    self._communicate_factors()

        if self.steps % self.kfac_update_freq == 0:
            if not self.exclude_compute_inverse:
                self._compute_inverse()

        if not self.exclude_compute_inverse:
            self._compute_pred()

        if not self.exclude_communicate_inverse and not self.communicate_inverse_or_not:
            if backend.comm.size() > 1:
                self._ensure_slab()
                self._communicate_pred()

        if not self.exclude_compute_inverse:
            self._update_grad_in_place()

        self.steps += 1
        self.m_a, self.m_g = {}, {}


class KFAC_EIGEN_DP(KFAC):
    """``eigen_dp``: kfac/kfac_preconditioner_eigen_dp.py."""
    variant = 'eigen_dp'


class KFAC_INV_DP(KFAC):
    """``inverse_dp``: kfac/kfac_preconditioner_inv_dp.py."""
    variant = 'inverse_dp'


class _MPDUnsupported(KFAC):
    """The model-parallel ('MPD') variants ``inverse`` / ``eigen`` of the reference
    (kfac_preconditioner_inv.py / kfac_preconditioner_eigen.py: factor all-reduce +
    inverse broadcast) are outside the hot path this package accelerates
    (SURVEY.md section 8f rank 3).  The keys exist so ``kfac_mappers.keys()`` matches."""
    variant = None

    def __init__(self, *args, **kwargs):
        raise NotImplementedError(
            "kfac_pytorch_b200 implements the distributed-preconditioning variants "
            "'eigen_dp' and 'inverse_dp' only; '%s' (MPD-KFAC) is not provided" % self._name)


class KFAC_INV(_MPDUnsupported):
    _name = 'inverse'


class KFAC_EIGEN(_MPDUnsupported):
    _name = 'eigen'
