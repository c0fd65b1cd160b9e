#!/usr/bin/env python
"""Generate the golden fixtures by running the UNMODIFIED reference package
(lzhangbv/kfac_pytorch, mounted read-only at /root/reference) on the seeded
cases of ``cases.py``.  Runs only in the authoring container (the reference
does not exist on the GPU box); the resulting ``*.npz`` files are committed.

    python tests/golden/make_golden.py [case ...]

The reference imports ``horovod.torch`` unconditionally (kfac/backend.py:2)
and its Torch backend calls ``hvd.synchronize`` (kfac/backend.py:164), so a
stub module providing exactly that is injected before the import.  Nothing in
the reference is modified.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('KFAC_REF', '/root/reference')


def _import_reference():
    hvd = types.ModuleType('horovod.torch')
    hvd.synchronize = lambda h: None
    hvd.size = lambda: 1
    hv = types.ModuleType('horovod')
    hv.torch = hvd
    sys.modules['horovod'] = hv
    sys.modules['horovod.torch'] = hvd
    # the reference must win over the repo's own drop-in `kfac` package
    sys.path = [REF] + [p for p in sys.path if os.path.abspath(p or '.') != ROOT]
    for k in [k for k in sys.modules if k == 'kfac' or k.startswith('kfac.')]:
        del sys.modules[k]
    import kfac  # noqa
    assert os.path.abspath(kfac.__file__).startswith(REF), kfac.__file__
    return kfac


def _run_case(case, rank, world_size, dtype, port):
    sys.path.insert(0, ROOT)
    from tests.golden.cases import layer_input_hw, digest
    from oracle.kfac_oracle import specs_from_model, synth_layer_inputs, synth_grads
    kfac = _import_reference()
    import kfac.backend as backend
    if not dist.is_initialized():
        os.environ['MASTER_ADDR'] = '127.0.0.1'
        os.environ['MASTER_PORT'] = str(port)
        os.environ['LOCAL_RANK'] = str(rank)
        dist.init_process_group('gloo', rank=rank, world_size=world_size)
    backend.comm = None
    backend.init('Torch')

    torch.manual_seed(0)
    model = case.build().to(dtype)
    specs, mods = specs_from_model(model)
    hw_of = layer_input_hw(model.float(), case.hw)
    model.to(dtype)
    KFAC = kfac.get_kfac_module(case.variant)
    pre = KFAC(model, lr=case.lr, damping=case.damping, fac_update_freq=case.fac_update_freq,
               kfac_update_freq=case.kfac_update_freq, kl_clip=case.kl_clip,
               factor_decay=case.factor_decay, hook_enabled=False, **case.extra)
    assert pre.modules == mods
    mpd = not case.variant.endswith('_dp')
    if pre.module_ranks is None:
        pre.schedule_module_ranks()
    out = {}
    for s in range(case.steps):
        if case.lr_schedule is not None:
            pre.param_groups[0]['lr'] = case.lr_schedule[s]
        gen = torch.Generator().manual_seed(case.seed + 1000 * s + rank)
        ggen = torch.Generator().manual_seed(case.seed + 1000 * s + 500)
        for spec, mod in zip(specs, mods):
            li = synth_layer_inputs(spec, case.batch, hw_of.get(mod), case.seq, gen, dtype)
            if pre.steps % pre.param_groups[0]['fac_update_freq'] == 0:
                ra, rg = pre.module_ranks[mod]
                if mpd or rank == ra:
                    pre.m_a[mod] = li.a
                if mpd or rank == rg:
                    pre.m_g[mod] = li.g
            w, b = synth_grads(spec, ggen, dtype)
            mod.weight.grad = w
            if b is not None:
                mod.bias.grad = b
        pre.step()
        for l, mod in enumerate(mods):
            out[f's{s}_l{l}_wgrad'] = digest(mod.weight.grad)
            if mod.bias is not None:
                out[f's{s}_l{l}_bgrad'] = digest(mod.bias.grad)
            ra, _ = pre.module_ranks[mod]
            if mpd or rank == ra:
                out[f's{s}_l{l}_A'] = digest(pre.m_A[mod])
                out[f's{s}_l{l}_G'] = digest(pre.m_G[mod])
            if case.variant.startswith('eigen') and (mpd or rank == ra):
                out[f"s{s}_l{l}_dA"] = pre.m_dA[mod].numpy().copy()
                out[f"s{s}_l{l}_dG"] = pre.m_dG[mod].numpy().copy()
            if case.variant.startswith('inverse') and (rank == ra or case.extra.get('communicate_inverse_or_not')):
                out[f's{s}_l{l}_invA'] = digest(pre.m_inv_A[mod])
                out[f's{s}_l{l}_invG'] = digest(pre.m_inv_G[mod])
            out[f's{s}_l{l}_precon'] = digest(pre.m_precon_grad[mod])
    return out


def _worker(rank, case_name, world_size, port, retq):
    sys.path.insert(0, ROOT)
    from tests.golden.cases import CASES_BY_NAME
    case = CASES_BY_NAME[case_name]
    torch.set_num_threads(2)
    res = {}
    for tag, dtype in (('ref32', torch.float32), ('ref64', torch.float64)):
        o = _run_case(case, rank, world_size, dtype, port)
        for k, v in o.items():
            res[f'{tag}_r{rank}_{k}'] = v
    retq.put((rank, res))
    dist.barrier()


def hooks_fixture():
    """Real forward/backward through the reference's hooks on a small model;
    the captured a/g and the incoming grads are stored so that the new
    package's hook path can be replayed bit-for-bit on the same tensors."""
    sys.path.insert(0, ROOT)
    from tests.golden.cases import hooks_small
    kfac = _import_reference()
    import kfac.backend as backend
    if not dist.is_initialized():
        os.environ['MASTER_ADDR'] = '127.0.0.1'
        os.environ['MASTER_PORT'] = '29611'
        os.environ['LOCAL_RANK'] = '0'
        dist.init_process_group('gloo', rank=0, world_size=1)
    backend.comm = None
    backend.init('Torch')
    torch.manual_seed(0)
    model = hooks_small()
    pre = kfac.get_kfac_module('eigen_dp')(model, lr=0.1, damping=0.003, kl_clip=0.001)
    out = {}
    for i, p in enumerate(model.parameters()):
        out[f'param{i}'] = p.detach().numpy().copy()
    gen = torch.Generator().manual_seed(7)
    for s in range(2):
        x = torch.randn(4, 3, 8, 8, generator=gen)
        y = torch.randint(0, 10, (4,), generator=gen)
        model.zero_grad()
        loss = torch.nn.functional.cross_entropy(model(x), y)
        loss.backward()
        out[f's{s}_x'] = x.numpy().copy()
        out[f's{s}_y'] = y.numpy().copy()
        for l, mod in enumerate(pre.modules):
            out[f's{s}_l{l}_a'] = pre.m_a[mod].numpy().copy()
            out[f's{s}_l{l}_g'] = pre.m_g[mod].numpy().copy()
            out[f's{s}_l{l}_wgrad_in'] = mod.weight.grad.numpy().copy()
            if mod.bias is not None:
                out[f's{s}_l{l}_bgrad_in'] = mod.bias.grad.numpy().copy()
        pre.step()
        for l, mod in enumerate(pre.modules):
            out[f's{s}_l{l}_wgrad_out'] = mod.weight.grad.numpy().copy()
            if mod.bias is not None:
                out[f's{s}_l{l}_bgrad_out'] = mod.bias.grad.numpy().copy()
    np.savez_compressed(os.path.join(HERE, 'hooks_small.npz'), **out)
    print('hooks_small', sum(v.nbytes for v in out.values()) // 1024, 'KiB')


def main():
    sys.path.insert(0, ROOT)
    from tests.golden.cases import CASES
    want = set(sys.argv[1:])
    port = 29620
    for case in CASES:
        if want and case.name not in want:
            continue
        port += 1
        ctx = mp.get_context('spawn')
        q = ctx.Queue()
        procs = [ctx.Process(target=_worker, args=(r, case.name, case.world_size, port, q))
                 for r in range(case.world_size)]
        for p in procs:
            p.start()
        merged = {}
        for _ in procs:
            _, res = q.get()
            merged.update(res)
        for p in procs:
            p.join()
        path = os.path.join(HERE, case.name + '.npz')
        np.savez_compressed(path, **merged)
        print(case.name, os.path.getsize(path) // 1024, 'KiB')
    if not want or 'hooks_small' in want:
        hooks_fixture()


if __name__ == '__main__':
    main()
