"""Collective-communication shim with the surface of the reference's
``kfac.backend`` (kfac/backend.py:21-164): a module-global ``comm`` created by
``init("Torch")`` / ``init("Horovod")`` exposing ``size / rank / local_rank /
new_group / allreduce[_][_async_] / broadcast[_][_async_] / synchronize`` and
the ``Average`` / ``Sum`` ops.

Differences from the reference, all on failure paths:
  * ``init`` RAISES when the requested backend is not initialised (the
    reference *returns* the exception object, kfac/backend.py:40-48);
  * Horovod is imported lazily (the reference imports it unconditionally at
    module import, kfac/backend.py:2) -- on a B200 box the data path is
    ``torch.distributed`` over NCCL/NVLink;
  * ``_TorchBackend.synchronize`` does not call into Horovod
    (kfac/backend.py:164 does, which breaks the Torch backend at size > 1);
  * ``init("Single")`` (or constructing a preconditioner with no backend
    initialised and no process group) gives a 1-rank backend so a single GPU
    needs no process group.
"""
from __future__ import annotations

import enum
import os

import torch
import torch.distributed as dist

comm = None


class Ops(enum.Enum):
    Average = "average"
    Sum = "sum"


def init(backend: str):
    global comm
    if comm is None:
        comm = _get_comm_backend(backend)


def _get_comm_backend(backend: str):
    if backend == "Horovod":
        try:
            import horovod.torch as hvd  # noqa: F401
            hvd.size()
        except Exception as e:  # not installed or not initialised
            raise RuntimeError('Horovod must be installed and hvd.init() called before '
                               'backend.init("Horovod")') from e
        return _HorovodBackend(hvd)
    if backend == "Torch":
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError('torch.distributed must be initialised before backend.init("Torch")')
        return _TorchBackend()
    if backend == "Single":
        return _SingleBackend()
    raise RuntimeError('The backend is not implemented. Only Horovod, Torch and Single are supported.')


def ensure_initialized():
    """Used by the preconditioner constructor: pick a backend if the caller has not."""
    global comm
    if comm is None:
        comm = _TorchBackend() if (dist.is_available() and dist.is_initialized()) else _SingleBackend()
    return comm


class _BackendBase:
    def __init__(self):
        self.Average = Ops.Average
        self.Sum = Ops.Sum

    def allreduce(self, tensor, name=None, op=Ops.Average):
        self.allreduce_(tensor, name, op)

    def broadcast(self, tensor, src, group=None, name=None):
        self.broadcast_(tensor, src, group, name)


class _SingleBackend(_BackendBase):
    """One process, one GPU: every collective is the identity."""

    def size(self):
        return 1

    def rank(self):
        return 0

    def local_rank(self):
        return 0

    def new_group(self, ranks):
        return None

    def allreduce_(self, tensor, name=None, op=Ops.Average):
        return None

    def allreduce_async_(self, tensor, name=None, op=Ops.Average):
        return None

    def broadcast_(self, tensor, src, group=None, name=None):
        return None

    def broadcast_async_(self, tensor, src, group=None, name=None):
        return None

    def reduce_async_(self, tensor, dst, op=Ops.Average):
        return None

    def synchronize(self, handle):
        return None


class _TorchBackend(_BackendBase):
    """torch.distributed (NCCL over NVLink on a B200 box, gloo in CPU tests)."""

    def size(self):
        return dist.get_world_size()

    def rank(self):
        return dist.get_rank()

    def local_rank(self):
        try:
            return int(os.environ['LOCAL_RANK'])
        except KeyError:
            raise RuntimeError('LOCAL_RANK must be set in the environment when using torch.distributed')

    def new_group(self, ranks):
        return dist.new_group(ranks)

    def allreduce_(self, tensor, name=None, op=Ops.Average):
        dist.all_reduce(tensor, async_op=False)
        if op == Ops.Average:
            tensor.div_(self.size())

    def allreduce_async_(self, tensor, name=None, op=Ops.Average):
        handle = dist.all_reduce(tensor, async_op=True)
        return handle if op == Ops.Sum else (handle, tensor)

    def broadcast_(self, tensor, src, group=None, name=None):
        dist.broadcast(tensor, src=src, group=group, async_op=False)

    def broadcast_async_(self, tensor, src, group=None, name=None):
        return dist.broadcast(tensor, src=src, group=group, async_op=True)

    def reduce_async_(self, tensor, dst, op=Ops.Average):
        """Reduce-to-owner (not in the reference backend: the fused gradient-reduce path).
        NCCL averages on the wire; gloo (CPU tests) sums and the destination divides."""
        if op == Ops.Average and dist.get_backend() == 'nccl':
            return dist.reduce(tensor, dst=dst, op=dist.ReduceOp.AVG, async_op=True)
        handle = dist.reduce(tensor, dst=dst, op=dist.ReduceOp.SUM, async_op=True)
        if op == Ops.Average and self.rank() == dst:
            return (handle, tensor)
        return handle

    def synchronize(self, handle):
        if handle is None:
            return
        if isinstance(handle, tuple):
            h, tensor = handle
            h.wait()
            tensor.div_(self.size())
        else:
            handle.wait()


class _HorovodBackend(_BackendBase):
    def __init__(self, hvd):
        super().__init__()
        self.hvd = hvd

    def size(self):
        return self.hvd.size()

    def rank(self):
        return self.hvd.rank()

    def local_rank(self):
        return self.hvd.local_rank()

    def new_group(self, ranks):
        return self.hvd.add_process_set(ranks)

    def _op(self, op):
        if op == Ops.Average:
            return self.hvd.Average
        if op == Ops.Sum:
            return self.hvd.Sum
        raise ValueError('Unknown communication operation {}'.format(op))

    def allreduce_(self, tensor, name=None, op=Ops.Average):
        self.hvd.allreduce_(tensor, name=name, op=self._op(op))

    def allreduce_async_(self, tensor, name=None, op=Ops.Average):
        return self.hvd.allreduce_async_(tensor, name=name, op=self._op(op))

    def broadcast_(self, tensor, src, group=None, name=None):
        if group is None:
            self.hvd.broadcast_(tensor, root_rank=src, name=name)
        else:
            self.hvd.broadcast_(tensor, root_rank=src, process_set=group, name=name)

    def broadcast_async_(self, tensor, src, group=None, name=None):
        if group is None:
            return self.hvd.broadcast_async_(tensor, root_rank=src, name=name)
        return self.hvd.broadcast_async_(tensor, root_rank=src, process_set=group, name=name)

    def synchronize(self, handle):
        return self.hvd.synchronize(handle)
