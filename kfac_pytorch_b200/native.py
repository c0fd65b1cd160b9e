"""ctypes binding of ``libkfacb.so`` (C ABI declared in ``include/kfacb.h``).

The product path has NO fallback: if the library is missing or no CUDA device
is present, constructing :class:`NativeEngine` raises.  Nothing here imports
``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libkfacb.so')

KFACB_LINEAR, KFACB_CONV2D = 0, 1


class LayerDesc(C.Structure):
    """Mirror of ``kfacb_layer_t``."""
    _fields_ = [(n, C.c_int32) for n in (
        'kind', 'has_bias', 'batch', 'seq', 'in_c', 'out_c', 'in_h', 'in_w',
        'kh', 'kw', 'sh', 'sw', 'ph', 'pw', 'out_h', 'out_w')]

    @property
    def dA(self):
        k = self.kh * self.kw if self.kind == KFACB_CONV2D else 1
        return self.in_c * k + (1 if self.has_bias else 0)

    @property
    def dG(self):
        return self.out_c


EXPORTS = [
    'kfacb_version', 'kfacb_build_fingerprint', 'kfacb_launch_count', 'kfacb_last_error', 'kfacb_device_info', 'kfacb_layer_dA', 'kfacb_layer_dG',
    'kfacb_factor_workspace_bytes', 'kfacb_factor_update',
    'kfacb_eigh_workspace_bytes', 'kfacb_eigh_batched', 'kfacb_eigh_last_profile', 'kfacb_eigh_plan_teams',
    'kfacb_spd_inverse_workspace_bytes', 'kfacb_pi_damped_inverse',
    'kfacb_precond_workspace_bytes', 'kfacb_precond_eigen', 'kfacb_precond_inverse',
    'kfacb_finalize_workspace_bytes', 'kfacb_finalize_grads',
    'kfacb_pack_grads', 'kfacb_finalize_dot', 'kfacb_finalize_write',
    'kfacb_precond_eigen_bcast', 'kfacb_precond_inverse_bcast',
    'kfacb_exchange_create', 'kfacb_exchange_handle', 'kfacb_exchange_connect', 'kfacb_exchange_slab',
    'kfacb_exchange_peer_deltas', 'kfacb_exchange_begin', 'kfacb_exchange_signal', 'kfacb_exchange_push',
    'kfacb_exchange_wait', 'kfacb_exchange_release', 'kfacb_exchange_destroy',
]
EXCHANGE_HANDLE_BYTES = 128

_lib = None


class KfacbError(RuntimeError):
    pass


def load_library(path: Optional[str] = None):
    """dlopen the C-ABI library and declare prototypes.  Raises if absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise KfacbError(
            'libkfacb.so not found at %s: build it with `python -m kfac_pytorch_b200.build` '
            '(there is no CPU / PyTorch fallback for preconditioner.step())' % p)
    lib = C.CDLL(p)
    PP = C.POINTER(C.c_void_p)
    I32P = C.POINTER(C.c_int32)
    LP = C.POINTER(LayerDesc)
    lib.kfacb_version.restype = C.c_int
    lib.kfacb_build_fingerprint.restype = C.c_char_p
    lib.kfacb_launch_count.restype = C.c_uint64
    lib.kfacb_last_error.restype = C.c_char_p
    lib.kfacb_device_info.argtypes = [I32P, I32P, I32P]
    lib.kfacb_layer_dA.argtypes = [LP]
    lib.kfacb_layer_dA.restype = C.c_int32
    lib.kfacb_layer_dG.argtypes = [LP]
    lib.kfacb_layer_dG.restype = C.c_int32
    lib.kfacb_factor_workspace_bytes.argtypes = [LP, C.c_int32]
    lib.kfacb_factor_workspace_bytes.restype = C.c_size_t
    lib.kfacb_factor_update.argtypes = [LP, C.c_int32, PP, PP, PP, PP, C.c_float, C.c_int32,
                                        C.c_void_p, C.c_size_t, C.c_void_p]
    lib.kfacb_eigh_workspace_bytes.argtypes = [I32P, C.c_int32]
    lib.kfacb_eigh_workspace_bytes.restype = C.c_size_t
    lib.kfacb_eigh_batched.argtypes = [I32P, C.c_int32, PP, PP, PP, C.c_float, C.c_void_p,
                                       C.c_void_p, C.c_size_t, C.c_void_p]
    lib.kfacb_eigh_last_profile.argtypes = [C.POINTER(C.c_double)]
    lib.kfacb_eigh_plan_teams.argtypes = [I32P, C.c_int32, C.c_int32, I32P]
    lib.kfacb_eigh_plan_teams.restype = C.c_int
    lib.kfacb_spd_inverse_workspace_bytes.argtypes = [I32P, I32P, C.c_int32]
    lib.kfacb_spd_inverse_workspace_bytes.restype = C.c_size_t
    lib.kfacb_pi_damped_inverse.argtypes = [I32P, I32P, C.c_int32, PP, PP, PP, PP, C.c_float,
                                            C.c_void_p, C.c_size_t, C.c_void_p]
    lib.kfacb_precond_workspace_bytes.argtypes = [LP, C.c_int32]
    lib.kfacb_precond_workspace_bytes.restype = C.c_size_t
    lib.kfacb_precond_eigen.argtypes = [LP, C.c_int32, PP, PP, PP, PP, PP, PP, PP, C.c_float,
                                        C.c_void_p, C.c_size_t, C.c_void_p]
    lib.kfacb_precond_inverse.argtypes = [LP, C.c_int32, PP, PP, PP, PP, PP,
                                          C.c_void_p, C.c_size_t, C.c_void_p]
    lib.kfacb_finalize_workspace_bytes.argtypes = [LP, C.c_int32]
    lib.kfacb_finalize_workspace_bytes.restype = C.c_size_t
    lib.kfacb_finalize_grads.argtypes = [LP, C.c_int32, PP, PP, PP, C.c_float, C.c_float,
                                         C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.kfacb_pack_grads.argtypes = [LP, C.c_int32, PP, PP, PP, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.kfacb_finalize_dot.argtypes = [LP, C.c_int32, PP, PP, PP, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.kfacb_finalize_write.argtypes = [LP, C.c_int32, PP, PP, PP, C.c_void_p, C.c_float, C.c_float, C.c_int32,
                                         C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    I64P = C.POINTER(C.c_int64)
    lib.kfacb_precond_eigen_bcast.argtypes = [LP, C.c_int32, PP, PP, PP, PP, PP, PP, PP, C.c_float, C.c_int32, I64P,
                                              C.c_void_p, C.c_size_t, C.c_void_p]
    lib.kfacb_precond_inverse_bcast.argtypes = [LP, C.c_int32, PP, PP, PP, PP, PP, C.c_int32, I64P,
                                                C.c_void_p, C.c_size_t, C.c_void_p]
    lib.kfacb_exchange_create.argtypes = [C.c_int32, C.c_int32, C.c_size_t, C.POINTER(C.c_void_p)]
    lib.kfacb_exchange_handle.argtypes = [C.c_void_p, C.c_void_p]
    lib.kfacb_exchange_connect.argtypes = [C.c_void_p, C.c_void_p]
    lib.kfacb_exchange_slab.argtypes = [C.c_void_p]
    lib.kfacb_exchange_slab.restype = C.c_void_p
    lib.kfacb_exchange_peer_deltas.argtypes = [C.c_void_p, I64P, I32P]
    for fn in ('begin', 'signal', 'wait', 'release'):
        getattr(lib, 'kfacb_exchange_' + fn).argtypes = [C.c_void_p, C.c_void_p]
    lib.kfacb_exchange_push.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
    lib.kfacb_exchange_destroy.argtypes = [C.c_void_p]
    for name in EXPORTS:
        getattr(lib, name)
    if path is None:
        _lib = lib
    return lib


def _ptrs(tensors: Sequence[Optional[torch.Tensor]]):
    arr = (C.c_void_p * max(1, len(tensors)))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


def _i32(vals: Sequence[int]):
    return (C.c_int32 * max(1, len(vals)))(*vals)


def _descs(layers: Sequence[LayerDesc]):
    arr = (LayerDesc * max(1, len(layers)))()
    for i, l in enumerate(layers):
        arr[i] = l
    return arr


def _check_f32(t: Optional[torch.Tensor], what: str):
    if t is None:
        return
    if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
        raise KfacbError('%s must be a contiguous float32 CUDA tensor (got %s, %s, contiguous=%s)'
                         % (what, t.dtype, t.device, t.is_contiguous()))


class NativeEngine:
    """Thin, stateless-ish wrapper: owns only growable device workspaces."""

    def __init__(self, device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise KfacbError('no CUDA device: the K-FAC hot path is CUDA only (sm_100a), '
                             'there is no CPU fallback')
        self.lib = load_library()
        self.device = torch.device(device if device is not None else
                                   ('cuda:%d' % torch.cuda.current_device()))
        self._ws = {}
        with torch.cuda.device(self.device):
            sm, maj, mnr = C.c_int32(), C.c_int32(), C.c_int32()
            self._call(self.lib.kfacb_device_info(C.byref(sm), C.byref(maj), C.byref(mnr)))
            self.sm_count, self.cc = sm.value, (maj.value, mnr.value)
        if self.cc[0] != 10:
            raise KfacbError('libkfacb.so is built for sm_100a only; device is sm_%d%d' % self.cc)

    # -- helpers -----------------------------------------------------------
    def _call(self, rc: int):
        if rc != 0:
            raise KfacbError('kfacb status %d: %s' % (rc, self.lib.kfacb_last_error().decode()))

    def _workspace(self, key: str, nbytes: int) -> torch.Tensor:
        cur = self._ws.get(key)
        if cur is None or cur.numel() < nbytes:
            cur = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=self.device)
            self._ws[key] = cur
        return cur

    def launch_count(self) -> int:
        return int(self.lib.kfacb_launch_count())

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    # -- (1) factors -------------------------------------------------------
    def factor_update(self, layers, a, g, A, G, alpha: float, first_step: bool, workspace_key: str = 'factor'):
        """a[i] / g[i] (and the matching A[i] / G[i]) may be None: that side of layer i is skipped."""
        for t in list(a) + list(g) + list(A) + list(G):
            if t is not None:
                _check_f32(t, 'factor_update operand')
        d = _descs(layers)
        n = len(layers)
        nbytes = self.lib.kfacb_factor_workspace_bytes(d, n)
        cur = self._ws.get(workspace_key)
        if workspace_key != 'factor' and cur is not None and cur.numel() < nbytes:
            # a side-stream workspace may still be in use by the previous launch: drain that stream before it is replaced
            torch.cuda.current_stream(self.device).synchronize()
        ws = self._workspace(workspace_key, nbytes)
        with torch.cuda.device(self.device):
            self._call(self.lib.kfacb_factor_update(d, n, _ptrs(a), _ptrs(g), _ptrs(A), _ptrs(G),
                                                    alpha, int(first_step), ws.data_ptr(),
                                                    ws.numel(), self._stream()))

    # -- (2a) eigen --------------------------------------------------------
    def eigh(self, mats, Qs, ds, eps_floor: float, sweeps_out: Optional[torch.Tensor] = None):
        for t in list(mats) + list(Qs) + list(ds):
            _check_f32(t, 'eigh operand')
        ns = _i32([int(m.shape[0]) for m in mats])
        n = len(mats)
        nbytes = self.lib.kfacb_eigh_workspace_bytes(ns, n)
        ws = self._workspace('eigh', nbytes)
        with torch.cuda.device(self.device):
            self._call(self.lib.kfacb_eigh_batched(
                ns, n, _ptrs(mats), _ptrs(Qs), _ptrs(ds), eps_floor,
                None if sweeps_out is None else sweeps_out.data_ptr(),
                ws.data_ptr(), ws.numel(), self._stream()))

    def eigh_last_profile(self):
        """Totals of the last KFACB_PROFILE=1 eigh call (see include/kfacb.h)."""
        out = (C.c_double * 8)()
        self._call(self.lib.kfacb_eigh_last_profile(out))
        if os.environ.get('KFACB_EIGH_ALGO', '')[:1] in ('j', 'J'):
            keys = ('rounds', 'ms_h', 'ms_solve', 'ms_update', 'bytes_h', 'bytes_update', 'matrices', 'pairs')
        else:   # tridiagonalisation + divide and conquer: stage times of the last call (CUDA events; synchronises on them)
            keys = ('ms_trd', 'ms_dc', 'ms_bt', 'bytes_trd', 'sum_n3', 'matrices', 'grid', 'reserved')
        return dict(zip(keys, list(out)))

    # -- (2b) pi-damped inverse ---------------------------------------------
    def pi_damped_inverse(self, A, G, invA, invG, damping: float):
        for t in list(A) + list(G) + list(invA) + list(invG):
            _check_f32(t, 'inverse operand')
        nA = _i32([0 if t is None else int(t.shape[0]) for t in A])
        nG = _i32([0 if t is None else int(t.shape[0]) for t in G])
        n = len(A)
        nbytes = self.lib.kfacb_spd_inverse_workspace_bytes(nA, nG, n)
        ws = self._workspace('inverse', nbytes)
        with torch.cuda.device(self.device):
            self._call(self.lib.kfacb_pi_damped_inverse(
                nA, nG, n, _ptrs(A), _ptrs(G), _ptrs(invA), _ptrs(invG), damping,
                ws.data_ptr(), ws.numel(), self._stream()))

    # -- (3) preconditioning -------------------------------------------------
    @staticmethod
    def _peer_args(peer_deltas):
        k = len(peer_deltas) if peer_deltas else 0
        return k, (C.c_int64 * max(1, k))(*(peer_deltas or []))

    def precond_eigen(self, layers, QA, QG, dA, dG, wgrad, bgrad, out, damping: float, peer_deltas=None):
        for t in list(QA) + list(QG) + list(dA) + list(dG) + list(wgrad) + list(bgrad) + list(out):
            _check_f32(t, 'precond operand')
        d = _descs(layers)
        n = len(layers)
        nbytes = self.lib.kfacb_precond_workspace_bytes(d, n)
        ws = self._workspace('precond', nbytes)
        with torch.cuda.device(self.device):
            if peer_deltas:      # fused exchange: the last GEMM's epilogue also stores into every peer's slab
                k, arr = self._peer_args(peer_deltas)
                self._call(self.lib.kfacb_precond_eigen_bcast(
                    d, n, _ptrs(QA), _ptrs(QG), _ptrs(dA), _ptrs(dG), _ptrs(wgrad), _ptrs(bgrad),
                    _ptrs(out), damping, k, arr, ws.data_ptr(), ws.numel(), self._stream()))
            else:
                self._call(self.lib.kfacb_precond_eigen(
                    d, n, _ptrs(QA), _ptrs(QG), _ptrs(dA), _ptrs(dG), _ptrs(wgrad), _ptrs(bgrad),
                    _ptrs(out), damping, ws.data_ptr(), ws.numel(), self._stream()))

    def precond_inverse(self, layers, invA, invG, wgrad, bgrad, out, peer_deltas=None):
        for t in list(invA) + list(invG) + list(wgrad) + list(bgrad) + list(out):
            _check_f32(t, 'precond operand')
        d = _descs(layers)
        n = len(layers)
        nbytes = self.lib.kfacb_precond_workspace_bytes(d, n)
        ws = self._workspace('precond', nbytes)
        with torch.cuda.device(self.device):
            if peer_deltas:
                k, arr = self._peer_args(peer_deltas)
                self._call(self.lib.kfacb_precond_inverse_bcast(
                    d, n, _ptrs(invA), _ptrs(invG), _ptrs(wgrad), _ptrs(bgrad), _ptrs(out), k, arr,
                    ws.data_ptr(), ws.numel(), self._stream()))
            else:
                self._call(self.lib.kfacb_precond_inverse(
                    d, n, _ptrs(invA), _ptrs(invG), _ptrs(wgrad), _ptrs(bgrad), _ptrs(out),
                    ws.data_ptr(), ws.numel(), self._stream()))

    # -- (4) KL clip + in-place rewrite ----------------------------------------
    def finalize(self, layers, precon, wgrad, bgrad, lr: float, kl_clip: Optional[float],
                 force_nu_one: bool, nu_out: torch.Tensor):
        for t in list(precon) + list(wgrad) + list(bgrad) + [nu_out]:
            _check_f32(t, 'finalize operand')
        d = _descs(layers)
        n = len(layers)
        nbytes = self.lib.kfacb_finalize_workspace_bytes(d, n)
        ws = self._workspace('finalize', nbytes)
        with torch.cuda.device(self.device):
            self._call(self.lib.kfacb_finalize_grads(
                d, n, _ptrs(precon), _ptrs(wgrad), _ptrs(bgrad), lr,
                float(kl_clip) if kl_clip else 0.0, int(force_nu_one), nu_out.data_ptr(),
                ws.data_ptr(), ws.numel(), self._stream()))


    # -- (4b) fused gradient-reduce path: pack / owner-partial dot / rewrite from the global sum ------------------
    def _fin_ws(self, d, n):
        return self._workspace('finalize', self.lib.kfacb_finalize_workspace_bytes(d, n))

    def pack_grads(self, layers, wgrad, bgrad, dst):
        for t in list(wgrad) + list(bgrad) + list(dst):
            if t is not None:
                _check_f32(t, 'pack operand')
        d, n = _descs(layers), len(layers)
        ws = self._fin_ws(d, n)
        with torch.cuda.device(self.device):
            self._call(self.lib.kfacb_pack_grads(d, n, _ptrs(wgrad), _ptrs(bgrad), _ptrs(dst),
                                                 ws.data_ptr(), ws.numel(), self._stream()))

    def finalize_dot(self, layers, precon, wgrad, bgrad, vg_sum_out: torch.Tensor):
        if vg_sum_out.dtype != torch.float64 or vg_sum_out.numel() < 1:
            raise KfacbError('finalize_dot: vg_sum_out must be a float64 device tensor')
        d, n = _descs(layers), len(layers)
        ws = self._fin_ws(d, n)
        with torch.cuda.device(self.device):
            self._call(self.lib.kfacb_finalize_dot(d, n, _ptrs(precon), _ptrs(wgrad), _ptrs(bgrad),
                                                   vg_sum_out.data_ptr(), ws.data_ptr(), ws.numel(), self._stream()))

    def finalize_write(self, layers, precon, wgrad, bgrad, vg_sum: Optional[torch.Tensor], lr: float,
                       kl_clip: Optional[float], force_nu_one: bool, nu_out: torch.Tensor):
        d, n = _descs(layers), len(layers)
        ws = self._fin_ws(d, n)
        with torch.cuda.device(self.device):
            self._call(self.lib.kfacb_finalize_write(
                d, n, _ptrs(precon), _ptrs(wgrad), _ptrs(bgrad), None if vg_sum is None else vg_sum.data_ptr(), lr,
                float(kl_clip) if kl_clip else 0.0, int(force_nu_one), nu_out.data_ptr(),
                ws.data_ptr(), ws.numel(), self._stream()))


class _RawCudaBuffer:
    """Exposes library-owned device memory to torch (zero copy) through ``__cuda_array_interface__``."""

    def __init__(self, ptr: int, nfloats: int):
        self.__cuda_array_interface__ = dict(shape=(nfloats,), typestr='<f4', data=(ptr, False), version=2)


class Exchange:
    """``kfacb_exchange_*``: the symmetric slab of preconditioned gradients of the ranks of one node (CUDA IPC peer
    memory) and its begin / signal / push / wait / release flag kernels (include/kfacb.h section 3b)."""

    def __init__(self, engine: NativeEngine, rank: int, size: int, nfloats: int):
        self.eng, self.lib = engine, engine.lib
        self.rank, self.size, self.nfloats = rank, size, int(nfloats)
        self._h = C.c_void_p()
        with torch.cuda.device(engine.device):
            engine._call(self.lib.kfacb_exchange_create(rank, size, self.nfloats * 4, C.byref(self._h)))
            ptr = self.lib.kfacb_exchange_slab(self._h)
            self.slab = torch.as_tensor(_RawCudaBuffer(ptr, max(1, self.nfloats)), device=engine.device)
        self.peer_deltas: List[int] = []

    def handle(self) -> bytes:
        buf = (C.c_ubyte * EXCHANGE_HANDLE_BYTES)()
        with torch.cuda.device(self.eng.device):
            self.eng._call(self.lib.kfacb_exchange_handle(self._h, buf))
        return bytes(buf)

    def connect(self, handles: Sequence[bytes]):
        assert len(handles) == self.size and all(len(h) == EXCHANGE_HANDLE_BYTES for h in handles)
        blob = b''.join(handles)
        buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        with torch.cuda.device(self.eng.device):
            self.eng._call(self.lib.kfacb_exchange_connect(self._h, buf))
            d = (C.c_int64 * max(1, self.size))()
            k = C.c_int32()
            self.eng._call(self.lib.kfacb_exchange_peer_deltas(self._h, d, C.byref(k)))
        self.peer_deltas = [int(d[i]) for i in range(k.value)]

    def _op(self, name):
        with torch.cuda.device(self.eng.device):
            self.eng._call(getattr(self.lib, 'kfacb_exchange_' + name)(self._h, self.eng._stream()))

    def begin(self):
        self._op('begin')

    def signal(self):
        self._op('signal')

    def wait(self):
        self._op('wait')

    def release(self):
        self._op('release')

    def push(self, offset_floats: int, nfloats: int):
        with torch.cuda.device(self.eng.device):
            self.eng._call(self.lib.kfacb_exchange_push(self._h, offset_floats * 4, nfloats * 4, self.eng._stream()))

    def close(self):
        if self._h:
            self.slab = None
            self.lib.kfacb_exchange_destroy(self._h)
            self._h = C.c_void_p()
