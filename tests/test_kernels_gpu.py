"""Kernel-level GPU tests through the C ABI (NativeEngine): each entry point
against the oracle / fp64 LAPACK on seeded inputs, including the edge shapes
(ragged tiles, n = 1, odd dA from the bias column, strided/padded convs,
split-K tall-skinny factors, degenerate spectra)."""
import numpy as np
import pytest
import torch

from oracle import kfac_oracle as orc
from tests.oracle_engine import spec_of

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.fixture(scope='module')
def eng():
    from kfac_pytorch_b200.native import NativeEngine
    return NativeEngine(torch.device(DEV))


def _desc(module, a, g):
    from kfac_pytorch_b200.preconditioner import _full_desc
    return _full_desc(module, a, g)


def _rel(x, ref):
    x = x.double().cpu(); ref = ref.double().cpu()
    return float((x - ref).norm() / ref.norm().clamp_min(1e-300))


CONV_CASES = [
    # (Cin, Cout, k, stride, pad, bias, B, H)
    (3, 8, 7, 2, 3, False, 4, 32),      # ResNet stem
    (16, 16, 3, 1, 1, False, 4, 14),
    (16, 24, 3, 2, 1, False, 3, 15),    # odd size, stride 2
    (24, 40, 1, 1, 0, False, 5, 9),     # pointwise fast path
    (24, 40, 1, 2, 0, False, 5, 9),     # strided 1x1 (im2col path)
    (3, 16, 3, 1, 1, True, 6, 8),       # bias -> odd dA = 28
    (64, 32, 3, 1, 1, True, 2, 12),     # dA = 577
    (8, 200, 3, 1, 0, False, 2, 6),     # dG > 128 (2 tiles), no padding
    (4, 4, 3, 1, 1, False, 16, 64),     # tall-skinny: rows 65536 -> split-K
    (130, 20, 3, 1, 1, True, 2, 5),     # dA = 1171 (10 tiles, ragged)
]


@pytest.mark.parametrize('cfg', CONV_CASES, ids=lambda c: 'c%d_%d_k%ds%dp%d_b%d_B%d_H%d' % (c[0], c[1], c[2], c[3], c[4], int(c[5]), c[6], c[7]))
@pytest.mark.parametrize('first', [True, False])
def test_factor_conv(eng, cfg, first):
    cin, cout, k, s, p, bias, B, H = cfg
    gen = torch.Generator().manual_seed(hash(cfg) % 100000)
    m = torch.nn.Conv2d(cin, cout, k, stride=s, padding=p, bias=bias)
    a = torch.randn(B, cin, H, H + 1, generator=gen)
    oh, ow = (H + 2 * p - k) // s + 1, (H + 1 + 2 * p - k) // s + 1
    g = torch.randn(B, cout, oh, ow, generator=gen) / B
    d = _desc(m, a, g)
    spec = spec_of(d)
    A0 = torch.randn(d.dA, d.dA, generator=gen); A0 = A0 @ A0.t()
    G0 = torch.randn(d.dG, d.dG, generator=gen); G0 = G0 @ G0.t()
    A, G = A0.to(DEV).contiguous(), G0.to(DEV).contiguous()
    eng.factor_update([d], [a.to(DEV)], [g.to(DEV)], [A], [G], 0.95, first)
    curA = torch.eye(d.dA, dtype=torch.float64) if first else A0.double()
    curG = torch.eye(d.dG, dtype=torch.float64) if first else G0.double()
    refA = orc.running_average(orc.factor_A(a.double(), spec), curA, 0.95)
    refG = orc.running_average(orc.factor_G(g.double(), spec), curG, 0.95)
    # 3xTF32 tensor-core accumulation: ~5e-7 of sum|a||b| per entry (fp32 SGEMM: ~1e-7); bar = 6e-6 << 1e-4
    assert _rel(A, refA) < 6e-6 and _rel(G, refG) < 6e-6
    assert torch.equal(A, A.t()) and torch.equal(G, G.t())      # exactly symmetric


@pytest.mark.parametrize('cfg', [(20, 16, True, 8, None), (784, 512, True, 32, None), (33, 130, False, 5, None),
                                 (24, 48, True, 4, 6), (768, 96, True, 3, 50), (7, 3, False, 2, 3)],
                         ids=str)
def test_factor_linear(eng, cfg):
    din, dout, bias, B, T = cfg
    gen = torch.Generator().manual_seed(din * 1000 + dout)
    m = torch.nn.Linear(din, dout, bias=bias)
    shape_a = (B, din) if T is None else (B, T, din)
    shape_g = (B, dout) if T is None else (B, T, dout)
    a = torch.randn(*shape_a, generator=gen); g = torch.randn(*shape_g, generator=gen) / B
    d = _desc(m, a, g)
    spec = spec_of(d)
    A = torch.empty(d.dA, d.dA, device=DEV); G = torch.empty(d.dG, d.dG, device=DEV)
    eng.factor_update([d], [a.to(DEV)], [g.to(DEV)], [A], [G], 0.95, True)
    refA = orc.running_average(orc.factor_A(a.double(), spec), torch.eye(d.dA, dtype=torch.float64), 0.95)
    refG = orc.running_average(orc.factor_G(g.double(), spec), torch.eye(d.dG, dtype=torch.float64), 0.95)
    assert _rel(A, refA) < 6e-6 and _rel(G, refG) < 6e-6
    # skipping: NULL a leaves A untouched
    A2 = A.clone()
    eng.factor_update([d], [None], [g.to(DEV)], [A2], [G], 0.95, False)
    assert torch.equal(A, A2)


def _spectra(n, kind, gen):
    if kind == 'wishart':
        x = torch.randn(3 * n + 5, n, generator=gen, dtype=torch.float64)
        return x.t() @ x / x.shape[0]
    if kind == 'lowrank':        # the step-0 K-FAC factor: 0.05 I + 0.95 * rank-r gram
        r = max(1, n // 8)
        x = torch.randn(r, n, generator=gen, dtype=torch.float64)
        return 0.05 * torch.eye(n, dtype=torch.float64) + 0.95 * x.t() @ x / r
    if kind == 'diag':
        return torch.diag(torch.linspace(0.0, 2.0, n, dtype=torch.float64))
    if kind == 'graded':         # eigenvalues over 8 decades, random basis
        q, _ = torch.linalg.qr(torch.randn(n, n, generator=gen, dtype=torch.float64))
        return (q * torch.logspace(-8, 0, n, dtype=torch.float64)) @ q.t()
    raise ValueError(kind)


@pytest.mark.parametrize('n', [1, 2, 3, 10, 33, 64, 65, 127, 128, 129, 147, 200, 256, 513, 1000])
@pytest.mark.parametrize('kind', ['wishart', 'lowrank', 'diag', 'graded'])
def test_eigh(eng, n, kind):
    gen = torch.Generator().manual_seed(n * 7 + len(kind))
    M64 = _spectra(n, kind, gen)
    M64 = 0.5 * (M64 + M64.t())
    M = M64.float().to(DEV).contiguous()
    Q = torch.empty(n, n, device=DEV); d = torch.empty(n, device=DEV)
    sw = torch.zeros(1, dtype=torch.int32, device=DEV)
    eng.eigh([M], [Q], [d], 1e-10, sw)
    ref_d = torch.linalg.eigvalsh(M.double().cpu())
    scale = ref_d.abs().max().item()
    dd, QQ = d.double().cpu(), Q.double().cpu()
    assert torch.all(dd[1:] >= dd[:-1]), 'ascending'
    ref_floor = ref_d * (ref_d > 1e-10)
    # fp32 eigensolver bars (LAPACK ssyevd lands at ~1e-6 / 1e-6 / 1e-6 on the same inputs)
    assert (dd - ref_floor).abs().max().item() <= 5e-6 * scale, ((dd - ref_floor).abs().max().item(), scale)
    assert (QQ.t() @ QQ - torch.eye(n, dtype=torch.float64)).abs().max().item() < 2e-5
    rec = (QQ * dd) @ QQ.t()
    assert (rec - M.double().cpu()).norm().item() <= 4e-5 * M.double().norm().item() + 1e-9 * scale
    assert 1 <= int(sw.item()) <= 30


def test_eigh_batched_mixed_sizes(eng):
    gen = torch.Generator().manual_seed(5)
    ns = [4, 300, 64, 129, 1, 513, 96, 257]
    mats = [(_spectra(n, 'wishart', gen)).float().to(DEV).contiguous() for n in ns]
    Qs = [torch.empty(n, n, device=DEV) for n in ns]
    ds = [torch.empty(n, device=DEV) for n in ns]
    sw = torch.zeros(len(ns), dtype=torch.int32, device=DEV)
    eng.eigh(mats, Qs, ds, 1e-10, sw)
    for M, Q, d in zip(mats, Qs, ds):
        ref = torch.linalg.eigvalsh(M.double().cpu())
        assert (d.double().cpu() - ref).abs().max().item() <= 5e-6 * ref.abs().max().item()
        QQ = Q.double().cpu()
        assert ((QQ * d.double().cpu()) @ QQ.t() - M.double().cpu()).norm() <= 4e-5 * M.double().norm().cpu()
    assert torch.all(sw > 0)


@pytest.mark.parametrize('knobs', [
    {'KFACB_SOLVE_SPLIT': '0'}, {'KFACB_SOLVE_SPLIT': '1'},
    {'KFACB_H_KCHUNK': '256', 'KFACB_UPD_TILES': '5'}, {'KFACB_H_KCHUNK': '512', 'KFACB_UPD_TILES': '16'},
], ids=lambda k: ','.join('%s=%s' % (a[6:], b) for a, b in k.items()))
def test_eigh_launch_shapes_do_not_change_the_result(eng, monkeypatch, knobs):
    """The host picks the solve layout (one CTA / a 2-CTA cluster per pair), the k-chunk of the projected-block
    GEMM and the tiles per update CTA from the group's size; every choice must give the same decomposition."""
    gen = torch.Generator().manual_seed(77)
    ns = [641, 300]
    mats = [(_spectra(n, 'lowrank', gen)).float().to(DEV).contiguous() for n in ns]

    def run():
        Qs = [torch.empty(n, n, device=DEV) for n in ns]
        ds = [torch.empty(n, device=DEV) for n in ns]
        sw = torch.zeros(len(ns), dtype=torch.int32, device=DEV)
        eng.eigh(mats, Qs, ds, 1e-10, sw)
        return Qs, ds, sw

    _, d_ref, sw_ref = run()
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    Qs, ds, sw = run()
    for M, Q, d, d0 in zip(mats, Qs, ds, d_ref):
        scale = d0.abs().max().item()
        assert (d - d0).abs().max().item() <= 2e-6 * scale
        QQ, dd = Q.double().cpu(), d.double().cpu()
        assert ((QQ * dd) @ QQ.t() - M.double().cpu()).norm() <= 4e-5 * M.double().norm().cpu()
        assert (QQ.t() @ QQ - torch.eye(M.shape[0], dtype=torch.float64)).abs().max().item() < 2e-5
    assert torch.all((sw - sw_ref).abs() <= 1)


@pytest.mark.parametrize('dims', [(5, 3), (64, 10), (65, 64), (200, 130), (513, 40), (1153, 128)], ids=str)
def test_pi_damped_inverse(eng, dims):
    nA, nG = dims
    gen = torch.Generator().manual_seed(nA + nG)
    A64 = _spectra(nA, 'lowrank', gen); G64 = _spectra(nG, 'wishart', gen) * 1e-3
    A = A64.float().to(DEV).contiguous(); G = G64.float().to(DEV).contiguous()
    iA = torch.empty_like(A); iG = torch.empty_like(G)
    A_in, G_in = A.double().cpu(), G.double().cpu()
    eng.pi_damped_inverse([A], [G], [iA], [iG], 0.03)
    Ad, Gd, ra, rg = orc.pi_damped_inverses(A_in, G_in, 0.03)
    assert _rel(A, Ad) < 1e-6 and _rel(G, Gd) < 1e-6          # in-place damping persisted
    assert _rel(iA, ra) < 1e-4 and _rel(iG, rg) < 1e-4
    assert torch.equal(iA, iA.t()) and torch.equal(iG, iG.t())
    # None entries are skipped
    eng.pi_damped_inverse([None, A], [None, G], [None, iA], [None, iG], 0.0)


@pytest.mark.parametrize('dims', [(20, 16, True), (785, 512, True), (147, 64, False), (577, 130, True), (64, 256, False)], ids=str)
def test_precondition_and_finalize(eng, dims):
    from kfac_pytorch_b200.native import LayerDesc, KFACB_LINEAR
    dA, dG, bias = dims
    gen = torch.Generator().manual_seed(dA * 3 + dG)
    d = LayerDesc(); d.kind = KFACB_LINEAR; d.has_bias = int(bias); d.batch = 1; d.seq = 1
    d.in_c = dA - int(bias); d.out_c = dG; d.kh = d.kw = d.sh = d.sw = 1
    spec = spec_of(d)
    A = _spectra(dA, 'lowrank', gen); G = _spectra(dG, 'wishart', gen) * 1e-2
    da, qa = torch.linalg.eigh(A); dg, qg = torch.linalg.eigh(G)
    w = torch.randn(dG, dA - int(bias), generator=gen); b = torch.randn(dG, generator=gen) if bias else None
    gm = orc.grad_matrix(spec, w.double(), None if b is None else b.double())
    ref = orc.precondition_eigen(gm, qa, da, qg, dg, 0.003)
    dev = lambda t: None if t is None else t.float().to(DEV).contiguous()
    out = torch.empty(dG, dA, device=DEV)
    wd, bd = dev(w), dev(b)
    eng.precond_eigen([d], [dev(qa)], [dev(qg)], [dev(da)], [dev(dg)], [wd], [bd], [out], 0.003)
    assert _rel(out, ref) < 2e-5
    iA = torch.linalg.inv(A + 0.1 * torch.eye(dA, dtype=torch.float64)); iG = torch.linalg.inv(G + 0.1 * torch.eye(dG, dtype=torch.float64))
    out2 = torch.empty(dG, dA, device=DEV)
    eng.precond_inverse([d], [dev(iA)], [dev(iG)], [wd], [bd], [out2])
    assert _rel(out2, orc.precondition_inverse(gm, iA, iG)) < 2e-5
    # finalize: KL clip + in-place rewrite
    nu = torch.zeros(1, device=DEV)
    vw, vb = orc.split_preconditioned(spec, out.double().cpu(), w.shape, None if b is None else b.shape)
    vs, gs = [vw], [w.double()]
    if bias:
        vs.append(vb); gs.append(b.double())
    nu_ref = orc.kl_clip_scale(vs, gs, 0.1, 0.001)
    eng.finalize([d], [out], [wd], [bd], 0.1, 0.001, False, nu)
    assert abs(nu.item() - nu_ref) <= 1e-5 * nu_ref
    assert _rel(wd, vw * nu_ref) < 1e-5
    if bias:
        assert _rel(bd, vb * nu_ref) < 1e-5
    wd2 = dev(w)
    bd2 = dev(b)
    eng.finalize([d], [out], [wd2], [bd2], 0.1, None, False, nu)      # kl_clip disabled -> nu = 1
    assert nu.item() == 1.0 and _rel(wd2, vw) < 1e-6


@pytest.mark.parametrize('bias', [True, False])
def test_pack_grads_and_split_finalize(eng, bias):
    """kfacb_pack_grads / kfacb_finalize_dot / kfacb_finalize_write (the pieces of the fused gradient-reduce path): two
    layers; the split form with the dot summed over 'owners' equals kfacb_finalize_grads on the same data."""
    from kfac_pytorch_b200.native import LayerDesc, KFACB_LINEAR
    gen = torch.Generator().manual_seed(7)
    descs, vs, ws, bs = [], [], [], []
    for (dA, dG) in ((33, 20), (130, 9)):
        d = LayerDesc(); d.kind = KFACB_LINEAR; d.has_bias = int(bias); d.batch = 1; d.seq = 1
        d.in_c = dA - int(bias); d.out_c = dG; d.kh = d.kw = d.sh = d.sw = 1
        descs.append(d)
        vs.append(torch.randn(dG, dA, generator=gen).to(DEV))
        ws.append(torch.randn(dG, dA - int(bias), generator=gen).to(DEV))
        bs.append(torch.randn(dG, generator=gen).to(DEV) if bias else None)
    # pack: [W | b] per layer, contiguous
    dst = [torch.full((d.dA * d.dG,), float('nan'), device=DEV) for d in descs]
    eng.pack_grads(descs, ws, bs, dst)
    for d, w, b, o in zip(descs, ws, bs, dst):
        nw = w.numel()
        assert torch.equal(o[:nw], w.reshape(-1))
        if bias:
            assert torch.equal(o[nw:], b)
    # dot of each 'owner' (one layer each) from the PACKED gradients, summed like the 8-byte all-reduce would
    tot = torch.zeros(1, dtype=torch.float64, device=DEV)
    part = torch.zeros(1, dtype=torch.float64, device=DEV)
    for d, v, o, w in zip(descs, vs, dst, ws):
        nw = w.numel()
        eng.finalize_dot([d], [v], [o[:nw]], [o[nw:] if bias else None], part)
        tot += part
    ref = sum(float((v.double() * (torch.cat([w, b[:, None]], 1) if bias else w).double()).sum()) for v, w, b in zip(vs, ws, bs))
    assert abs(tot.item() - ref) <= 1e-6 * abs(ref) + 1e-9            # (the products are formed in fp32 like the reference, the sum in double)
    # write from the global sum == the monolithic call
    nu1, nu2 = torch.zeros(1, device=DEV), torch.zeros(1, device=DEV)
    w1 = [w.clone() for w in ws]; b1 = [None if b is None else b.clone() for b in bs]
    w2 = [w.clone() for w in ws]; b2 = [None if b is None else b.clone() for b in bs]
    eng.finalize(descs, vs, w1, b1, 0.1, 0.001, False, nu1)
    eng.finalize_write(descs, vs, w2, b2, tot, 0.1, 0.001, False, nu2)
    assert abs(nu1.item() - nu2.item()) <= 1e-6 * nu1.item() and nu1.item() < 1.0
    for a, c in zip(w1 + [x for x in b1 if x is not None], w2 + [x for x in b2 if x is not None]):
        assert torch.allclose(a, c, rtol=1e-6, atol=0)
    eng.finalize_write(descs, vs, w2, b2, None, 0.1, None, True, nu2)        # clip off: nu = 1, no sum needed
    assert nu2.item() == 1.0 and torch.equal(w2[0], vs[0][:, :w2[0].shape[1]].contiguous())


def test_bad_arguments_return_status_not_abort(eng):
    from kfac_pytorch_b200.native import KfacbError, LayerDesc
    d = LayerDesc(); d.kind = 7; d.batch = 1; d.in_c = 2; d.out_c = 2
    t = torch.zeros(2, 2, device=DEV)
    with pytest.raises(KfacbError, match='unsupported layer kind'):
        eng.factor_update([d], [t], [t], [t], [t], 0.95, True)
    with pytest.raises(KfacbError, match='contiguous float32 CUDA'):
        eng.eigh([torch.zeros(2, 2)], [t], [t[0]], 1e-10)


@pytest.mark.parametrize('scale', [1.0, 2.5e-3], ids=str)
def test_eigh_accuracy_does_not_depend_on_the_matrix_norm(eng, scale):
    """K-FAC's conv A factors are ~0.0025 I + small (kfac/utils.py:85-96 divides by the spatial size twice).  The divide
    and conquer's deflation tolerance must carry the norm of the tridiagonal matrix (LAPACK scales to unit norm first);
    an absolute tolerance reconstructed ResNet-50's conv1 factor to 6.9e-5 only (profiles/r02_eigh_accuracy.txt).
    CPU twin: tests/test_eigh_model.py."""
    gen = torch.Generator(device=DEV).manual_seed(3)
    for n in (147, 576):
        x = torch.randn(400 * n, n, device=DEV, generator=gen)
        A = (scale * (0.05 * torch.eye(n, device=DEV) + 0.95 * (x.t() @ x) / x.shape[0])).contiguous()
        Q = torch.empty_like(A); d = torch.empty(n, device=DEV)
        eng.eigh([A], [Q], [d], 1e-10)
        Ad, Qd, dd = A.double(), Q.double(), d.double()
        rec = (((Qd * dd) @ Qd.t() - Ad).norm() / Ad.norm()).item()
        ref = torch.linalg.eigvalsh(Ad)
        assert rec < 2e-5, (scale, n, rec)
        assert (dd - ref).abs().max().item() <= 1e-5 * ref.abs().max().item(), (scale, n)


@pytest.mark.parametrize('ns', [(2304,), (4608, 512), (1664, 1664)], ids=str)
def test_eigh_back_to_back_without_host_sync(eng, ns):
    """Pipelined use (consecutive step() calls, no host synchronisation in between -- what bench.py and real
    training do): every call must deliver a full decomposition.  Round 1 shipped a host-written progress flag
    that the still queued kernels of the previous call overwrote, truncating lone large factors."""
    gen = torch.Generator().manual_seed(sum(ns))
    calls = 3
    mats = [[(_spectra(n, 'lowrank', gen)).float().to(DEV).contiguous() for n in ns] for _ in range(calls)]
    Qs = [[torch.empty(n, n, device=DEV) for n in ns] for _ in range(calls)]
    ds = [[torch.empty(n, device=DEV) for n in ns] for _ in range(calls)]
    torch.cuda.synchronize()
    for c in range(calls):                       # no sync, no D2H between the calls
        eng.eigh(mats[c], Qs[c], ds[c], 1e-10)
    torch.cuda.synchronize()
    for c in range(calls):
        for M, Q, d in zip(mats[c], Qs[c], ds[c]):
            n = M.shape[0]
            Md, QQ, dd = M.double(), Q.double(), d.double()
            rec = ((QQ * dd) @ QQ.t() - Md).norm() / Md.norm()
            orth = (QQ.t() @ QQ - torch.eye(n, dtype=torch.float64, device=DEV)).abs().max()
            assert rec.item() <= 4e-5, (c, n, rec.item())
            assert orth.item() < 2e-5, (c, n, orth.item())
            ref = torch.linalg.eigvalsh(Md)
            assert (dd - ref * (ref > 1e-10)).abs().max().item() <= 5e-6 * ref.abs().max().item()
