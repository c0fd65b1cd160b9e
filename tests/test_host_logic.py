"""CPU tests (no GPU): host orchestration of the product package driven through
a TEST-ONLY oracle-backed engine, the C-ABI library's exported symbols, the
scheduler and the backend shim."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from tests.golden.cases import CASES, CASES_BY_NAME
from tests.harness import compare_with_fixture, load_fixture, rel_err, run_package_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fresh_backend():
    from kfac_pytorch_b200 import backend
    backend.comm = None
    return backend


@pytest.mark.parametrize('case', [c for c in CASES if c.world_size == 1], ids=lambda c: c.name)
def test_orchestration_single_rank(case):
    """KFAC.step() gating / EMA / identity init / pi-damping persistence / kl-clip
    wiring, with the math supplied by the oracle: must reproduce the reference."""
    from tests.oracle_engine import OracleEngine
    _fresh_backend()
    out, pre = run_package_case(case, torch.device('cpu'), engine=OracleEngine())
    compare_with_fixture(case, out)
    assert pre.steps == case.steps and pre.m_a == {} and pre.m_g == {}


def _ws2_worker(rank, case_name, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from tests.oracle_engine import OracleEngine
    from kfac_pytorch_b200 import backend
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    os.environ['LOCAL_RANK'] = str(rank)
    torch.set_num_threads(1)
    dist.init_process_group('gloo', rank=rank, world_size=2)
    backend.comm = None
    backend.init('Torch')
    case = CASES_BY_NAME[case_name]
    out, pre = run_package_case(case, torch.device('cpu'), rank=rank, engine=OracleEngine())
    q.put((rank, out, [r for r, _ in pre.module_ranks.values()]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('case', [c for c in CASES if c.world_size == 2], ids=lambda c: c.name)
def test_orchestration_world_size_2_gloo(case):
    """Layer-partitioned run over 2 gloo ranks: round-robin ownership, owner-only
    factors from the owner's local batch, slab exchange, identical grads on both ranks."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29700 + (abs(hash(case.name)) % 200)
    procs = [ctx.Process(target=_ws2_worker, args=(r, case.name, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in procs:
        rank, out, owners = q.get(timeout=300)
        res[rank] = out
        if not case.extra.get('distribute_layer_factors'):
            assert owners == [i % 2 for i in range(len(owners))]
    for p in procs:
        p.join(timeout=60)
    for rank in (0, 1):
        compare_with_fixture(case, res[rank])
    # all ranks end with bit-identical gradients
    for k, v in res[0].items():
        if k.endswith('wgrad') or k.endswith('bgrad'):
            np.testing.assert_array_equal(v, res[1][k.replace('r0_', 'r1_', 1)])


def test_c_abi_exports_every_declared_symbol():
    """The shared library must load without a GPU and export everything
    include/kfacb.h declares (no compute calls here)."""
    from kfac_pytorch_b200 import build, native
    build.build(verbose=False)
    hdr = open(os.path.join(ROOT, 'include', 'kfacb.h')).read()
    declared = set(re.findall(r'^\s*(?:int|size_t|int32_t|uint64_t|const char\*|void\*)\s+(kfacb_\w+)\s*\(', hdr, re.M))
    assert declared == set(native.EXPORTS), declared ^ set(native.EXPORTS)
    lib = ctypes.CDLL(native.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    lib.kfacb_version.restype = ctypes.c_int
    assert lib.kfacb_version() >= 100
    d = native.LayerDesc()
    d.kind, d.has_bias, d.in_c, d.out_c, d.kh, d.kw = native.KFACB_CONV2D, 1, 16, 24, 3, 3
    lib.kfacb_layer_dA.argtypes = [ctypes.POINTER(native.LayerDesc)]
    assert lib.kfacb_layer_dA(ctypes.byref(d)) == 145 == d.dA
    assert native.load_library() is not None


def test_step_fails_loudly_without_cuda():
    import kfac_pytorch_b200 as kfac
    _fresh_backend()
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    model = torch.nn.Linear(4, 3)
    pre = kfac.DP_KFAC(model)
    model(torch.randn(2, 4)).sum().backward()
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        pre.step()


def test_api_surface_matches_reference():
    import kfac
    import kfac.backend as backend
    assert list(kfac.kfac_mappers.keys()) == ['inverse', 'eigen', 'inverse_dp', 'eigen_dp']
    assert kfac.get_kfac_module('eigen_dp') is kfac.kfac_mappers['eigen_dp']
    assert kfac.get_kfac_module() is kfac.kfac_mappers['eigen_dp']
    from kfac import DP_KFAC, KFACParamScheduler  # noqa: F401
    backend.comm = None
    model = torch.nn.Sequential(torch.nn.Linear(8, 4), torch.nn.LayerNorm(4), torch.nn.Linear(4, 30522))
    pre = DP_KFAC(model, inv_type='eigen', exclude_vocabulary_size=30522, exclude_parts='CommunicateInverse')
    assert isinstance(pre, torch.optim.Optimizer) and isinstance(pre, kfac.kfac_mappers['eigen_dp'])
    assert len(pre.modules) == 1 and pre.module_names == ['module_name_Linear_0']
    assert pre.exclude_communicate_inverse and not pre.exclude_compute_inverse
    assert isinstance(DP_KFAC(model, inv_type='inverse'), kfac.kfac_mappers['inverse_dp'])
    assert set(pre.param_groups[0]) >= {'lr', 'damping', 'fac_update_freq', 'kfac_update_freq'}
    pre.set_hook_enabled(False)
    assert pre.hook_enabled is False
    for name in ('size', 'rank', 'local_rank', 'new_group', 'allreduce', 'allreduce_', 'allreduce_async_',
                 'broadcast', 'broadcast_', 'broadcast_async_', 'synchronize', 'Average', 'Sum'):
        assert hasattr(backend.comm, name), name
    # the MPD variants keep the reference's positional constructor order (inv.py:35-45, eigen.py:36-46)
    inv = kfac.get_kfac_module('inverse')(model, 0.1, 0.001, 1, 1, True)
    assert inv.variant == 'inverse' and inv.communicate_inverse_or_not is True and inv.module_ranks is None
    eig = kfac.get_kfac_module('eigen')(model, 0.1, 0.001, 1, 1, None, 0.001)
    assert eig.variant == 'eigen' and eig.communicate_inverse_or_not is True and not eig.distribute_layer_factors
    backend.comm = None
    with pytest.raises(RuntimeError):
        backend.init('Torch')            # no process group: must RAISE (reference returns the error)


def test_param_scheduler_semantics():
    import kfac
    _fresh_backend()
    model = torch.nn.Linear(4, 3)
    pre = kfac.DP_KFAC(model, damping=0.03, fac_update_freq=10, kfac_update_freq=100)
    sched = kfac.KFACParamScheduler(pre, damping_alpha=0.5, damping_schedule=[2, 4],
                                    update_freq_alpha=0.5, update_freq_schedule=[3], start_epoch=0)
    seen = []
    for _ in range(5):
        sched.step()
        g = pre.param_groups[0]
        seen.append((sched.epoch, round(g['damping'], 6), g['fac_update_freq'], g['kfac_update_freq']))
    assert seen == [(1, 0.03, 10, 100), (2, 0.015, 10, 100), (3, 0.015, 5, 50),
                    (4, 0.0075, 5, 50), (5, 0.0075, 5, 50)]
    sched.step(epoch=0)
    assert pre.param_groups[0]['damping'] == 0.03
    # LambdaLR drives lr through param_groups like any optimizer (reference: examples/pytorch_imagenet_resnet.py:318)
    lrs = torch.optim.lr_scheduler.LambdaLR(pre, lambda e: 0.1 ** e)
    pre.param_groups[0]['lr']
    lrs.step()
    assert abs(pre.param_groups[0]['lr'] - 0.01) < 1e-12


def test_greedy_partition_balances_better_than_round_robin():
    from kfac_pytorch_b200 import backend
    import kfac_pytorch_b200 as kfac

    class FakeComm(backend._SingleBackend):
        def size(self):
            return 4
    backend.comm = FakeComm()
    try:
        dims = [(64, 64), (4608, 512), (64, 256), (4608, 512), (512, 2048), (4608, 512), (2048, 512), (2049, 1000)]
        model = torch.nn.Sequential(*[torch.nn.Linear(a, g, bias=False) for a, g in dims])

        def max_load(pre):
            load = [0.0] * 4
            for m in pre.modules:
                load[pre.module_ranks[m][0]] += m.in_features ** 3 + m.out_features ** 3
            return max(load)
        rr = kfac.DP_KFAC(model)
        gr = kfac.DP_KFAC(model, partition='greedy')
        assert [r for r, _ in rr.module_ranks.values()] == [i % 4 for i in range(len(dims))]
        assert all(a == g for a, g in gr.module_ranks.values())
        assert max_load(gr) < max_load(rr)
    finally:
        backend.comm = None


def test_kfac_state_round_trip_resumes_bit_identically():
    """save after 2 steps -> load into a fresh preconditioner -> step 3 equals the uninterrupted run."""
    from tests.oracle_engine import OracleEngine
    from tests.harness import build_case, case_inputs
    import kfac_pytorch_b200 as kfac
    _fresh_backend()
    case = CASES_BY_NAME['convnet_bias_inverse']          # damping persists in the factors: state matters
    def run(resume_at=None, state=None):
        model, specs, mods, hw_of = build_case(case)
        pre = kfac.get_kfac_module(case.variant)(model, lr=case.lr, damping=case.damping, kl_clip=case.kl_clip,
                                                 hook_enabled=False, engine=OracleEngine())
        start = 0
        if state is not None:
            pre.load_kfac_state_dict(state); start = resume_at
        saved = None
        for s in range(start, 3):
            per_rank, wg, bg = case_inputs(case, specs, mods, hw_of, s)
            for l, mod in enumerate(mods):
                pre.m_a[mod] = per_rank[0][l].a; pre.m_g[mod] = per_rank[0][l].g
                mod.weight.grad = wg[l].clone()
                if bg[l] is not None:
                    mod.bias.grad = bg[l].clone()
            pre.step()
            if s == 1 and state is None:
                saved = pre.kfac_state_dict()
        return [m.weight.grad.clone() for m in mods], saved
    full, saved = run()
    assert saved['steps'] == 2
    resumed, _ = run(resume_at=2, state=saved)
    for a, b in zip(full, resumed):
        assert torch.equal(a, b)


def test_eigh_team_scheduler_is_sane_for_resnet50():
    """kfacb_eigh_plan_teams (no GPU needed): every n > 128 factor gets a team, the persistent kernel never asks for
    more CTAs than the device has SMs, and the big factors get the big teams (n <= 64 takes the in-SMEM Jacobi path)."""
    from kfac_pytorch_b200 import native
    from kfac_pytorch_b200.workloads import layer_dims, resnet50_layers
    lib = native.load_library()
    ns = []
    for ls in resnet50_layers():
        ns += list(layer_dims(ls))
    for sms in (148, 132, 20):
        arr = (ctypes.c_int32 * len(ns))(*ns)
        out = (ctypes.c_int32 * len(ns))()
        grid = lib.kfacb_eigh_plan_teams(arr, len(ns), sms, out)
        assert 0 < grid <= sms
        for n, t in zip(ns, out):
            assert (t == 0) == (n <= 64) and t <= sms
        big = max(t for n, t in zip(ns, out) if n == 4608)
        assert all(t <= big for t in out)
    assert lib.kfacb_eigh_workspace_bytes(arr, len(ns)) > 4 * sum(4 * n * n for n in ns if n > 64)
    lone = (ctypes.c_int32 * 1)(4608)
    out1 = (ctypes.c_int32 * 1)()
    assert lib.kfacb_eigh_plan_teams(lone, 1, 148, out1) == out1[0] >= 64      # a lone big factor gets most of the GPU


def _fused_worker(rank, port, q, variant):
    import os
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE='2')
    dist.init_process_group('gloo', rank=rank, world_size=2)
    try:
        from kfac_pytorch_b200 import backend
        import kfac_pytorch_b200 as kfac
        from tests.oracle_engine import OracleEngine
        backend.comm = None
        backend.init('Torch')
        res = {}
        for fused in (False, True):
            torch.manual_seed(5)
            model = torch.nn.Sequential(torch.nn.Linear(7, 6), torch.nn.Tanh(), torch.nn.Linear(6, 5, bias=False),
                                        torch.nn.Tanh(), torch.nn.Linear(5, 3))
            pre = kfac.get_kfac_module(variant)(model, lr=0.1, damping=0.03, kl_clip=0.001, engine=OracleEngine(),
                                                fuse_grad_reduce=fused)
            gen = torch.Generator().manual_seed(100 + rank)            # every rank has its own mini-batch
            for s in range(3):
                x = torch.randn(9, 7, generator=gen)
                model.zero_grad()
                model(x).square().mean().backward()
                if not fused:                                          # the caller's gradient averaging (Horovod / DDP)
                    for p in model.parameters():
                        dist.all_reduce(p.grad)
                        p.grad.div_(2)
                pre.step()
            res[fused] = [p.grad.detach().numpy().copy() for p in model.parameters()]     # numpy: no fd passing
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('variant', ['eigen_dp', 'inverse_dp'])
def test_fused_gradient_reduce_matches_allreduce_then_step(variant):
    """SURVEY 8(f) rank 2 on 2 gloo ranks: reduce-to-owner -> precondition -> exchange + KL clip from the owners' partial
    sums gives the gradients of `all-reduce, then step()` (examples/pytorch_imagenet_resnet.py:371-383 followed by
    kfac_preconditioner_base.py:185-230) on every rank."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 200)
    procs = [ctx.Process(target=_fused_worker, args=(r, port, q, variant)) for r in range(2)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    for r in range(2):
        for a, b in zip(out[r][False], out[r][True]):
            assert np.allclose(a, b, rtol=2e-4, atol=1e-7), np.abs(a - b).max()     # fp32: gloo sums in another order
    for a, b in zip(out[0][True], out[1][True]):
        assert np.array_equal(a, b)         # every rank ends with identical gradients


def test_optional_flags_are_refused_for_the_mpd_variants():
    import kfac_pytorch_b200 as kfac
    from tests.oracle_engine import OracleEngine
    _fresh_backend()
    model = torch.nn.Linear(4, 3)
    for flag in ('overlap_factor_capture', 'fuse_grad_reduce'):
        with pytest.raises(ValueError, match='distributed-preconditioning'):
            kfac.get_kfac_module('eigen')(model, engine=OracleEngine(), **{flag: True})
    kfac.get_kfac_module('eigen_dp')(model, engine=OracleEngine(), overlap_factor_capture=True, fuse_grad_reduce=True)
