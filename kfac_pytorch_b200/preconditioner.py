"""Host side of the B200-native K-FAC preconditioner.

Mirrors the reference's Python surface.  The hot path this package is built
for is the distributed-preconditioning pair (``eigen_dp`` / ``inverse_dp``);
the model-parallel variants (``eigen`` / ``inverse``: full factor replicas +
all-reduce + distributed inverse, kfac_preconditioner_eigen.py /
kfac_preconditioner_inv.py) run on the same kernels with a different
communication schedule.

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
        # the reference's ComputeA/_get_grad assume dense convolutions (kfac/utils.py:33-54,85-94): a grouped conv
        # fails there with a matmul shape error; here it would index past the gradient buffer, so refuse it loudly
        if getattr(module, 'groups', 1) != 1:
            raise NotImplementedError('KFAC: Conv2d with groups=%d is not supported' % module.groups)
        if tuple(getattr(module, 'dilation', (1, 1))) != (1, 1):
            raise NotImplementedError('KFAC: dilated Conv2d is not supported')
        if getattr(module, 'padding_mode', 'zeros') != 'zeros':
            raise NotImplementedError('KFAC: Conv2d padding_mode=%r is not supported' % module.padding_mode)
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
    variant = None   # 'eigen_dp' | 'inverse_dp' (distributed preconditioning) | 'eigen' | 'inverse' (MPD)

    def __init__(self, model, lr=0.1, damping=0.001, fac_update_freq=1, kfac_update_freq=1,
                 kl_clip=0.001, factor_decay=0.95, exclude_vocabulary_size=None,
                 hook_enabled=True, exclude_parts='', *, partition='round_robin', engine=None,
                 communicate_inverse_or_not=False, distribute_layer_factors=None,
                 overlap_factor_capture=False, fuse_grad_reduce=False):
        defaults = dict(lr=lr, damping=damping, fac_update_freq=fac_update_freq,
                        kfac_update_freq=kfac_update_freq)
        super().__init__(model.parameters(), defaults)
        if self.variant not in ('eigen_dp', 'inverse_dp', 'eigen', 'inverse'):
            raise NotImplementedError('use kfac.get_kfac_module(...) / the KFAC_* classes')
        backend.ensure_initialized()
        self.is_dp = self.variant.endswith('_dp')
        self.uses_eigen = self.variant.startswith('eigen')
        self.fac_update_freq = fac_update_freq
        self.kfac_update_freq = kfac_update_freq
        # DP variants always exchange preconditioned grads (inv_dp.py:49); MPD eigen always exchanges the
        # eigenbases (eigen.py:53); MPD inverse takes the flag from its constructor (inv.py:42)
        self.communicate_inverse_or_not = (False if self.is_dp else
                                           (True if self.variant == 'eigen' else bool(communicate_inverse_or_not)))
        self.kl_clip = kl_clip if (kl_clip is not None and kl_clip > 0) else None
        self.factor_decay = factor_decay
        self.exclude_vocabulary_size = exclude_vocabulary_size
        self.hook_enabled = hook_enabled
        self.exclude_communicate_inverse = exclude_parts.find('CommunicateInverse') >= 0
        self.exclude_compute_inverse = exclude_parts.find('ComputeInverse') >= 0
        self.exclude_communicate_factor = exclude_parts.find('CommunicateFactor') >= 0
        self.exclude_compute_factor = exclude_parts.find('ComputeFactor') >= 0
        self.partition = partition
        # launch each layer's factor accumulation from its own forward / backward hook on a side
        # stream, so the SYRKs overlap the rest of the forward / backward pass and the captured activations are
        # released as soon as their kernel has run (keyword-only; DP variants; not part of the reference signature)
        self.overlap_factor_capture = bool(overlap_factor_capture)
        # the caller does NOT all-reduce the gradients; step() reduces each layer's gradient to its
        # owner only (averaged), the owner preconditions it, and the existing exchange distributes the result --
        # half the gradient bytes on the wire of all-reduce + exchange.  The KL-clip sum is formed from the owners'
        # partials with one 8-byte all-reduce (keyword-only; DP variants; not part of the reference signature)
        self.fuse_grad_reduce = bool(fuse_grad_reduce)
        if (self.overlap_factor_capture or self.fuse_grad_reduce) and not self.is_dp:
            raise ValueError('overlap_factor_capture / fuse_grad_reduce are implemented for the distributed-preconditioning '
                             'variants (eigen_dp, inverse_dp) only, not for %r' % self.variant)
        self._gslab = None
        self._vg = None
        self._cap_stream = None
        self._cap_done = {'a': set(), 'g': set()}

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
        self._ex = None
        self._ex_fused = False
        self._slab_segments = None
        self._nu = None
        # factor-wise distribution of the MPD eigen variant (eigen.py:66-71)
        if distribute_layer_factors is None:
            self.distribute_layer_factors = (self.variant == 'eigen' and
                                             backend.comm.size() > len(self.modules))
        else:
            self.distribute_layer_factors = bool(distribute_layer_factors) and self.variant == 'eigen'
        if self.is_dp:
            self.schedule_module_ranks()   # DP variants schedule eagerly (inv_dp.py:57)

    # ------------------------------------------------------------------ hooks
    def set_hook_enabled(self, mode=True):
        self.hook_enabled = mode

    def _forward_hook_event(self, module, input):
        """Owner-only capture of the layer input (kfac_preconditioner_inv_dp.py:60-65)."""
        if self.hook_enabled and torch.is_grad_enabled() and self.steps % self.fac_update_freq == 0:
            if not self.is_dp:                     # MPD: every rank keeps its local batch (base.py:122-125)
                self.m_a[module] = input[0].data
                return
            rank_a, _ = self.module_ranks[module]
            if backend.comm.rank() == rank_a:
                if self._overlap_active():
                    self._capture_now(module, 'a', input[0].data)
                else:
                    self.m_a[module] = input[0].data

    def _backward_hook_event(self, module, grad_input, grad_output):
        """Owner-only capture of the output gradient (kfac_preconditioner_inv_dp.py:67-72)."""
        if self.hook_enabled and self.steps % self.fac_update_freq == 0:
            if not self.is_dp:                     # MPD (base.py:127-130)
                self.m_g[module] = grad_output[0].data
                return
            _, rank_g = self.module_ranks[module]
            if backend.comm.rank() == rank_g:
                if self._overlap_active():
                    self._capture_now(module, 'g', grad_output[0].data)
                else:
                    self.m_g[module] = grad_output[0].data

    # -------------------------------------------- overlapped capture (8f rank 1)
    def _overlap_active(self) -> bool:
        return (self.overlap_factor_capture and self.is_dp and not self.exclude_compute_factor
                and self.modules[0].weight.device.type == 'cuda')

    def _capture_now(self, module, which: str, t: torch.Tensor):
        """Factor accumulation of ONE layer side ('a' -> A at forward time, 'g' -> G at backward time) launched
        from the hook on the capture stream: same kernel, same EMA / identity-init flags as ``_compute_factors``
        (kfac_preconditioner_inv_dp.py:60-90), only earlier and off the main stream.  The captured tensor is not
        retained: ``record_stream`` keeps its memory alive until the kernel has consumed it."""
        if module in self._cap_done[which]:
            raise RuntimeError('overlap_factor_capture=True needs exactly one forward/backward per step(): %s was '
                               'captured twice (gradient accumulation re-applies the running average; use the '
                               'default capture mode for that)' % self.module_names[self.modules.index(module)])
        eng = self._get_engine()
        dev = t.device
        if self._cap_stream is None:
            self._cap_stream = torch.cuda.Stream(device=dev)
        t = _prep(t)
        d = _full_desc(module, t if which == 'a' else None, t if which == 'g' else None)
        store, n = (self.m_A, d.dA) if which == 'a' else (self.m_G, d.dG)
        if module not in store:
            mk = torch.empty if self.steps == 0 else (lambda k, _k, **kw: torch.eye(k, **kw))
            store[module] = mk(n, n, dtype=torch.float32, device=dev)
        side = self._cap_stream
        side.wait_stream(torch.cuda.current_stream(dev))      # the tensor (and a fresh factor buffer) are ready
        with torch.cuda.stream(side):
            eng.factor_update([d], [t if which == 'a' else None], [t if which == 'g' else None],
                              [store[module] if which == 'a' else None], [store[module] if which == 'g' else None],
                              self.factor_decay, self.steps == 0, workspace_key='factor_capture')
        t.record_stream(side)
        self._cap_done[which].add(module)

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
            it = 0
            for module in self.modules:
                rank_a = it % size
                if self.distribute_layer_factors:  # factor-wise: G on the next rank (eigen.py:82-90)
                    it += 1
                    rank_g = it % size
                else:
                    rank_g = rank_a
                ranks[module] = (rank_a, rank_g)
                it += 1
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
        owned = self._owned() if self.is_dp else list(range(len(self.modules)))   # MPD: full replicas (inv.py:80-91)
        if self._cap_stream is not None and (self._cap_done['a'] or self._cap_done['g']):
            # overlapped capture: the hooks already launched these layers' SYRKs; join the capture stream
            torch.cuda.current_stream(self._device()).wait_stream(self._cap_stream)
            done = self._cap_done
            missing = [i for i in owned if (self.modules[i] in done['a']) != (self.modules[i] in done['g'])]
            if missing:
                raise RuntimeError('overlap_factor_capture: layer %s saw only one of its forward / backward hooks '
                                   'this step' % self.module_names[missing[0]])
            owned = [i for i in owned if self.modules[i] not in done['a']]
            self._cap_done = {'a': set(), 'g': set()}
        if not owned:
            return
        descs, a_l, g_l, A_l, G_l = [], [], [], [], []
        for i in owned:
            m = self.modules[i]
            if m not in self.m_a or m not in self.m_g:
                raise RuntimeError('%s: no captured input / output gradient for this step (run forward + backward with '
                                   'the hooks enabled, or inject m_a / m_g)' % self.module_names[i])
            a, g = _prep(self.m_a[m]), _prep(self.m_g[m])
            d = _full_desc(m, a, g)
            if m not in self.m_A:
                # step 0: the kernel overwrites with the identity-initialised EMA (first_step).  Later (a layer whose
                # factors were not restored from a checkpoint): start from the identity like step 0 did, never from
                # uninitialised memory
                mk = torch.empty if self.steps == 0 else (lambda n, _n, **kw: torch.eye(n, **kw))
                self.m_A[m] = mk(d.dA, d.dA, dtype=torch.float32, device=a.device)
                self.m_G[m] = mk(d.dG, d.dG, dtype=torch.float32, device=a.device)
            descs.append(d); a_l.append(a); g_l.append(g)
            A_l.append(self.m_A[m]); G_l.append(self.m_G[m])
        eng.factor_update(descs, a_l, g_l, A_l, G_l, self.factor_decay, self.steps == 0)

    def _communicate_factors(self):
        """DP: no factor communication (kfac_preconditioner_inv_dp.py:93-95).
        MPD: average-all-reduce of every factor (kfac_preconditioner_inv.py:94-103)."""
        if self.is_dp:
            return
        handles = []
        for m in self.modules:
            handles.append(backend.comm.allreduce_async_(self.m_A[m], op=backend.comm.Average))
            handles.append(backend.comm.allreduce_async_(self.m_G[m], op=backend.comm.Average))
        for h in handles:
            backend.comm.synchronize(h)

    # ---------------------------------------------------------------- phase 2
    def _compute_inverse(self):
        eng = self._get_engine()
        me = backend.comm.rank()
        if self.uses_eigen:
            # eigh + floor on the owner of each factor (eigen_dp.py:62-75; MPD eigen.py:98-119, where the A and
            # G factor of a layer may live on different ranks and non-owners start from zeros)
            mats, Qs, ds = [], [], []
            for i, m in enumerate(self.modules):
                rank_a, rank_g = self.module_ranks[m]
                own_a, own_g = me == rank_a, me == rank_g
                if self.is_dp and not own_a:
                    continue
                A, G = self.m_A.get(m), self.m_G.get(m)
                if m not in self.m_QA:
                    da, dg = self._static[i].dA, self._static[i].dG
                    dev = self._device()
                    mk = torch.empty if self.is_dp else torch.zeros
                    self.m_QA[m] = mk(da, da, dtype=torch.float32, device=dev)
                    self.m_dA[m] = mk(da, dtype=torch.float32, device=dev)
                    self.m_QG[m] = mk(dg, dg, dtype=torch.float32, device=dev)
                    self.m_dG[m] = mk(dg, dtype=torch.float32, device=dev)
                if own_a:
                    mats.append(A); Qs.append(self.m_QA[m]); ds.append(self.m_dA[m])
                if own_g:
                    mats.append(G); Qs.append(self.m_QG[m]); ds.append(self.m_dG[m])
            if mats:
                eng.eigh(mats, Qs, ds, self.eps)
        else:
            # pi-damped (in place, on the owner's copy only!) Cholesky-equivalent inverse
            # (inv_dp.py:98-123; MPD inv.py:109-129, where non-owners keep zero inverses)
            A_l, G_l, iA, iG = [], [], [], []
            for i, m in enumerate(self.modules):
                rank_a, rank_g = self.module_ranks[m]
                assert rank_a == rank_g
                own = me == rank_a
                if self.is_dp and not own:
                    continue
                if m not in self.m_inv_A:
                    mk = torch.empty_like if self.is_dp else torch.zeros_like
                    self.m_inv_A[m] = mk(self.m_A[m])
                    self.m_inv_G[m] = mk(self.m_G[m])
                if own:
                    A_l.append(self.m_A[m]); G_l.append(self.m_G[m])
                    iA.append(self.m_inv_A[m]); iG.append(self.m_inv_G[m])
            if A_l:
                eng.pi_damped_inverse(A_l, G_l, iA, iG, self.damping)

    def _communicate_inverse(self):
        """MPD only: broadcast the inverses / eigenbases from their owners
        (kfac_preconditioner_inv.py:132-142, kfac_preconditioner_eigen.py:122-134)."""
        handles = []
        for m in self.modules:
            rank_a, rank_g = self.module_ranks[m]
            if self.uses_eigen:
                handles.append(backend.comm.broadcast_async_(self.m_QA[m], rank_a))
                handles.append(backend.comm.broadcast_async_(self.m_dA[m], rank_a))
                handles.append(backend.comm.broadcast_async_(self.m_QG[m], rank_g))
                handles.append(backend.comm.broadcast_async_(self.m_dG[m], rank_g))
            else:
                handles.append(backend.comm.broadcast_async_(self.m_inv_A[m], rank_a))
                handles.append(backend.comm.broadcast_async_(self.m_inv_G[m], rank_g))
        for h in handles:
            backend.comm.synchronize(h)

    # ---------------------------------------------------------------- phase 3
    def _exchange_mode(self) -> str:
        """'p2p' (default for the DP variants when every rank sits on this node: symmetric slab over CUDA IPC peer
        memory, transfer fused into the last preconditioning GEMM), or 'nccl' (one broadcast per owner rank over
        the same slab layout).  KFACB_EXCHANGE overrides."""
        import os
        mode = os.environ.get('KFACB_EXCHANGE', '').lower()
        if mode in ('p2p', 'nccl', 'p2p_push'):
            return mode
        size = backend.comm.size()
        local = int(os.environ.get('LOCAL_WORLD_SIZE', '0') or 0)
        if self.is_dp and size > 1 and local == size and self._device().type == 'cuda':
            return 'p2p'
        return 'nccl'

    def _ensure_slab(self):
        """One flat buffer for all preconditioned gradients, laid out so that each
        owner rank's layers are contiguous: the exchange is one broadcast per rank
        (NCCL mode) or lands in every peer's copy while it is produced (P2P mode)."""
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
            off += (n + 3) & ~3                 # 16-byte aligned layers (vector copies of kfacb_exchange_push)
            seg[r][1] = off
        self._ex = None
        mode = self._exchange_mode() if (size > 1 and self.is_dp) else 'nccl'
        if mode.startswith('p2p'):
            import torch.distributed as dist
            from .native import Exchange
            ex = Exchange(self._get_engine(), backend.comm.rank(), size, max(off, 4))
            handles = [None] * size
            dist.all_gather_object(handles, ex.handle())
            ex.connect(handles)
            self._ex, self._ex_fused = ex, mode == 'p2p'
            self._slab = ex.slab
        else:
            self._slab = torch.zeros(max(off, 1), dtype=torch.float32, device=dev)
        self._slab_segments = seg
        self._slab_offsets = offs
        for i, m in enumerate(self.modules):
            d = self._static[i]
            self.m_precon_grad[m] = self._slab[offs[i]:offs[i] + d.dA * d.dG].view(d.dG, d.dA)
        self._nu = torch.ones(1, dtype=torch.float32, device=dev)

    def _use_exchange(self) -> bool:
        # the ablation flag exclude_parts='CommunicateInverse' must really skip the exchange (kfac_preconditioner_base.py:214-217)
        return (getattr(self, '_ex', None) is not None and not self.exclude_communicate_inverse
                and not self.communicate_inverse_or_not)

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
            st = self._static[i]
            if wg.numel() != st.dG * (st.dA - st.has_bias) or (bg is not None and bg.numel() != st.dG):
                raise RuntimeError('%s: gradient has %d elements, the layer geometry says %d'
                                   % (self.module_names[i], wg.numel(), st.dG * (st.dA - st.has_bias)))
            w.append(wg); b.append(bg)
        return w, b

    # ------------------------------------------------ fused gradient reduce (8f rank 2)
    def _fused_reduce_active(self) -> bool:
        return self.fuse_grad_reduce and self.is_dp and backend.comm.size() > 1

    def _reduce_grads_to_owners(self):
        """LOCAL (un-averaged) gradients -> averaged gradient of every layer on its owner only: one pack kernel into a
        send slab laid out like the exchange slab (owner-contiguous, per layer [W | b]), then one reduce per owner
        segment.  Replaces the caller's all-reduce (examples/pytorch_imagenet_resnet.py:371-383)."""
        eng = self._get_engine()
        self._ensure_slab()
        if self._gslab is None:
            self._gslab = torch.empty(self._slab.numel(), dtype=torch.float32, device=self._device())
        idx = list(range(len(self.modules)))
        w, b = self._grads(idx)
        dst = [self._gslab[self._slab_offsets[i]:self._slab_offsets[i] + self._static[i].dA * self._static[i].dG]
               for i in idx]
        eng.pack_grads(self._static, w, b, dst)
        handles = []
        for r, (beg, end) in enumerate(self._slab_segments):
            if beg is not None:
                handles.append(backend.comm.reduce_async_(self._gslab[beg:end], r, backend.comm.Average))
        for h in handles:
            backend.comm.synchronize(h)

    def _owner_grads(self, owned):
        """Views of the reduced send slab as (wgrad, bgrad) of the owned layers."""
        w, b = [], []
        for i in owned:
            st = self._static[i]
            seg = self._gslab[self._slab_offsets[i]:self._slab_offsets[i] + st.dA * st.dG]
            nw = st.dG * (st.dA - st.has_bias)
            w.append(seg[:nw])
            b.append(seg[nw:] if st.has_bias else None)
        return w, b

    def _update_grad_fused(self):
        """KL clip from the owners' partial sums (8-byte all-reduce) + rewrite of every gradient from the exchanged
        slab (kfac_preconditioner_inv.py:188-217; the rewrite only needs nu and v)."""
        eng = self._get_engine()
        owned = self._owned()
        if self._vg is None:
            self._vg = torch.zeros(1, dtype=torch.float64, device=self._device())
        use_clip = self.kl_clip is not None and not self.exclude_communicate_inverse
        if use_clip:
            w, b = self._owner_grads(owned)
            eng.finalize_dot([self._static[i] for i in owned], [self.m_precon_grad[self.modules[i]] for i in owned],
                             w, b, self._vg)
            backend.comm.allreduce_(self._vg, op=backend.comm.Sum)
        idx = list(range(len(self.modules)))
        w, b = self._grads(idx)
        eng.finalize_write(self._static, [self.m_precon_grad[m] for m in self.modules], w, b,
                           self._vg if use_clip else None, float(self.lr), self.kl_clip, not use_clip, self._nu)
        if self._use_exchange():
            self._ex.release()

    def _compute_pred(self):
        """Owner-only two-sided product (eigen_dp.py:78-93 / inv_dp.py:126-138)."""
        eng = self._get_engine()
        self._ensure_slab()
        # DP: owner only.  MPD: every rank multiplies every layer with whatever inverse it holds
        # (zeros on non-owners unless the inverses were broadcast; inv.py:156-161, eigen.py:137-144)
        owned = self._owned() if self.is_dp else list(range(len(self.modules)))
        use_ex = self._use_exchange()
        if use_ex:
            self._ex.begin()                   # every peer has finished reading the previous step's slab
        if owned:
            descs = [self._static[i] for i in owned]
            w, b = self._owner_grads(owned) if self._fused_reduce_active() else self._grads(owned)
            out = [self.m_precon_grad[self.modules[i]] for i in owned]
            mods = [self.modules[i] for i in owned]
            deltas = self._ex.peer_deltas if (use_ex and self._ex_fused) else None
            if self.uses_eigen:
                eng.precond_eigen(descs, [self.m_QA[m] for m in mods], [self.m_QG[m] for m in mods],
                                  [self.m_dA[m] for m in mods], [self.m_dG[m] for m in mods],
                                  w, b, out, self.damping, peer_deltas=deltas)
            else:
                eng.precond_inverse(descs, [self.m_inv_A[m] for m in mods],
                                    [self.m_inv_G[m] for m in mods], w, b, out, peer_deltas=deltas)
        if use_ex:
            if self._ex_fused:
                self._ex.signal()              # my segment has landed in every peer's slab
            else:
                beg, end = self._slab_segments[backend.comm.rank()]
                self._ex.push(beg or 0, 0 if beg is None else end - beg)

    def _communicate_pred(self):
        """Owner -> everyone exchange of the preconditioned gradients.  The reference
        issues one broadcast per layer (kfac_preconditioner_inv.py:164-175); the slab
        layout makes it one broadcast per owner rank."""
        if self._use_exchange():
            self._ex.wait()                    # every peer's segment has landed in the local slab
            return
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
        if self._use_exchange() and backend.comm.size() > 1:
            self._ex.release()                 # peers may overwrite the local slab from now on

    # ------------------------------------------------------ K-FAC state (resume)
    def kfac_state_dict(self):
        """Factors, inverses / eigenbases and the step counter, keyed by module index.  The reference keeps this
        state in plain dicts outside ``Optimizer.state`` and silently loses it on resume
        (SURVEY.md section 5 / 8f rank 4); ``torch.save`` this next to ``state_dict()``."""
        def pack(d):
            return {i: d[m].detach().to("cpu", copy=True) for i, m in enumerate(self.modules) if m in d}
        if self.module_ranks is None:
            self.schedule_module_ranks()
        return dict(variant=self.variant, steps=self.steps, world_size=backend.comm.size(),
                    rank=backend.comm.rank(), partition=self.partition,
                    owners=[self.module_ranks[m][0] for m in self.modules],
                    m_A=pack(self.m_A), m_G=pack(self.m_G), m_QA=pack(self.m_QA), m_QG=pack(self.m_QG),
                    m_dA=pack(self.m_dA), m_dG=pack(self.m_dG), m_inv_A=pack(self.m_inv_A),
                    m_inv_G=pack(self.m_inv_G))

    def load_kfac_state_dict(self, state):
        if state['variant'] != self.variant or state['world_size'] != backend.comm.size():
            raise ValueError('K-FAC state was saved for variant=%s world_size=%d' % (state['variant'], state['world_size']))
        if self.module_ranks is None:
            self.schedule_module_ranks()
        me = backend.comm.rank()
        # DP variants keep per-rank state (only the owned layers): every rank must load the state IT saved, with the
        # same ownership map; loading rank 0's file on rank 1 would silently drop rank 1's factors
        if self.is_dp:
            owners = [self.module_ranks[m][0] for m in self.modules]
            if 'owners' in state and list(state['owners']) != owners:
                raise ValueError('K-FAC state was saved with a different layer->rank ownership map')
            if 'rank' in state and int(state['rank']) != me and backend.comm.size() > 1:
                raise ValueError('K-FAC state of rank %d loaded on rank %d: the DP variants keep per-rank state, '
                                 'save and load one file per rank' % (int(state['rank']), me))
            if int(state['steps']) > 0:
                need = ['m_A', 'm_G'] + (['m_QA', 'm_QG', 'm_dA', 'm_dG'] if self.uses_eigen else ['m_inv_A', 'm_inv_G'])
                for i, o in enumerate(owners):
                    if o == me:
                        for name in need:
                            if i not in state[name] and str(i) not in state[name]:
                                raise ValueError('K-FAC state lacks %s of owned layer %d' % (name, i))
        dev = self._device()
        for name in ('m_A', 'm_G', 'm_QA', 'm_QG', 'm_dA', 'm_dG', 'm_inv_A', 'm_inv_G'):
            tgt = getattr(self, name)
            tgt.clear()
            for i, t in state[name].items():
                tgt[self.modules[int(i)]] = t.to(dev).contiguous()
        self.steps = int(state['steps'])

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
            if not self.exclude_communicate_inverse and self.communicate_inverse_or_not:
                if backend.comm.size() > 1:
                    self._communicate_inverse()

        if self._fused_reduce_active() and not self.exclude_compute_inverse:
            self._reduce_grads_to_owners()

        if not self.exclude_compute_inverse:
            self._compute_pred()

        if not self.exclude_communicate_inverse and not self.communicate_inverse_or_not:
            if backend.comm.size() > 1:
                self._ensure_slab()
                self._communicate_pred()

        if not self.exclude_compute_inverse:
            if self._fused_reduce_active():
                self._update_grad_fused()
            else:
                self._update_grad_in_place()

        self.steps += 1
        self.m_a, self.m_g = {}, {}


class KFAC_EIGEN_DP(KFAC):
    """``eigen_dp``: kfac/kfac_preconditioner_eigen_dp.py."""
    variant = 'eigen_dp'


class KFAC_INV_DP(KFAC):
    """``inverse_dp``: kfac/kfac_preconditioner_inv_dp.py."""
    variant = 'inverse_dp'


class KFAC_INV(KFAC):
    """``inverse`` (MPD-KFAC, explicit inverse): kfac/kfac_preconditioner_inv.py.  Full factor replicas
    averaged with an all-reduce, owner-only damped inverse, then either the preconditioned gradients
    (default) or the inverses are broadcast.  Same constructor order as the reference (:35-45)."""
    variant = 'inverse'

    def __init__(self, model, lr=0.1, damping=0.001, fac_update_freq=1, kfac_update_freq=1,
                 communicate_inverse_or_not=False, kl_clip=0.001, factor_decay=0.95,
                 exclude_vocabulary_size=None, hook_enabled=True, exclude_parts='', **kw):
        super().__init__(model, lr=lr, damping=damping, fac_update_freq=fac_update_freq,
                         kfac_update_freq=kfac_update_freq, kl_clip=kl_clip, factor_decay=factor_decay,
                         exclude_vocabulary_size=exclude_vocabulary_size, hook_enabled=hook_enabled,
                         exclude_parts=exclude_parts,
                         communicate_inverse_or_not=communicate_inverse_or_not, **kw)


class KFAC_EIGEN(KFAC):
    """``eigen`` (MPD-KFAC with eigendecomposition, SC'20): kfac/kfac_preconditioner_eigen.py.  Full factor
    replicas + all-reduce, distributed eigh (optionally factor-wise), broadcast of (Q, d), every rank
    preconditions.  Same constructor order as the reference (:36-46)."""
    variant = 'eigen'

    def __init__(self, model, lr=0.1, damping=0.001, fac_update_freq=1, kfac_update_freq=1,
                 distribute_layer_factors=None, kl_clip=0.001, factor_decay=0.95,
                 exclude_vocabulary_size=None, hook_enabled=True, exclude_parts='', **kw):
        super().__init__(model, lr=lr, damping=damping, fac_update_freq=fac_update_freq,
                         kfac_update_freq=kfac_update_freq, kl_clip=kl_clip, factor_decay=factor_decay,
                         exclude_vocabulary_size=exclude_vocabulary_size, hook_enabled=hook_enabled,
                         exclude_parts=exclude_parts,
                         distribute_layer_factors=distribute_layer_factors, **kw)
