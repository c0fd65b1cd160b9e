"""BASELINE-size checks (ResNet-50 / VGG-16 layer shapes) through size-independent properties:
the oracle cannot run a whole 4608-dim eigh in seconds on every box, so at full size we check
trace preservation, ordering, orthogonality, reconstruction, EMA linearity and inverse * matrix = I."""
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


def _kfac_like_factor(n, rows, gen):
    x = torch.randn(rows, n, device=DEV, generator=gen)
    return (0.05 * torch.eye(n, device=DEV) + 0.95 * (x.t() @ x) / rows).contiguous()


@pytest.mark.parametrize('n,rows', [(2304, 6272), (4608, 1568)])
def test_eigh_resnet50_factor_sizes(eng, n, rows):
    gen = torch.Generator(device=DEV).manual_seed(n)
    A = _kfac_like_factor(n, rows, gen)
    Q = torch.empty_like(A); d = torch.empty(n, device=DEV)
    sw = torch.zeros(1, dtype=torch.int32, device=DEV)
    eng.eigh([A], [Q], [d], 1e-10, sw)
    A64, Q64, d64 = A.double(), Q.double(), d.double()
    assert torch.all(d64[1:] >= d64[:-1])
    assert abs(d64.sum().item() - A64.trace().item()) <= 2e-6 * A64.trace().item()      # trace preserved
    assert (Q64.t() @ Q64 - torch.eye(n, device=DEV, dtype=torch.float64)).abs().max().item() < 1e-5
    rec = (Q64 * d64) @ Q64.t()
    assert (rec - A64).norm().item() <= 1e-4 * A64.norm().item()
    if n == 4608:       # rank(x^T x) = 1568 < n: the 0.05 cluster of the identity part must come out flat --
        # to the backward error of an fp32 eigensolver, a few eps * ||A|| (||A|| ~ 7 here; LAPACK ssyevd: ~3e-6)
        assert (d64[: n - rows] - 0.05).abs().max().item() < 1e-6 * d64.max().item()
    assert 1 <= int(sw.item()) <= 16


def test_factor_resnet50_conv3x3_full_size_vs_oracle(eng):
    """conv 64->64 3x3 pad 1 at 56x56, batch 32: rows = 100352, dA = 576 (tall-skinny, split-K)."""
    from kfac_pytorch_b200.preconditioner import _full_desc
    gen = torch.Generator().manual_seed(3)
    m = torch.nn.Conv2d(64, 64, 3, padding=1, bias=False)
    a = torch.randn(32, 64, 56, 56, generator=gen)
    g = torch.randn(32, 64, 56, 56, generator=gen) / 32
    d = _full_desc(m, a, g)
    A = torch.empty(576, 576, device=DEV); G = torch.empty(64, 64, device=DEV)
    eng.factor_update([d], [a.to(DEV)], [g.to(DEV)], [A], [G], 0.95, True)
    spec = spec_of(d)
    refA = orc.running_average(orc.factor_A(a, spec).double(), torch.eye(576, dtype=torch.float64), 0.95)
    refG = orc.running_average(orc.factor_G(g, spec).double(), torch.eye(64, dtype=torch.float64), 0.95)
    assert ((A.double().cpu() - refA).norm() / refA.norm()).item() < 1e-5
    assert ((G.double().cpu() - refG).norm() / refG.norm()).item() < 1e-5
    # EMA linearity: a second update with the same batch moves the factor by 0.05*(old - new) exactly
    A2 = A.clone(); G2 = G.clone()
    eng.factor_update([d], [a.to(DEV)], [g.to(DEV)], [A2], [G2], 0.95, False)
    newA = (refA - 0.05 * torch.eye(576, dtype=torch.float64)) / 0.95
    assert ((A2.double().cpu() - (0.05 * refA + 0.95 * newA)).norm() / refA.norm()).item() < 1e-5


def test_inverse_vgg16_factor_size(eng):
    n = 2305
    gen = torch.Generator(device=DEV).manual_seed(7)
    A = _kfac_like_factor(n, 8192, gen); G = _kfac_like_factor(256, 8192, gen) * 1e-3
    A0 = A.clone()
    iA = torch.empty_like(A); iG = torch.empty_like(G)
    eng.pi_damped_inverse([A], [G], [iA], [iG], 0.03)
    assert not torch.equal(A, A0)                       # damping persisted in place
    eye = torch.eye(n, device=DEV, dtype=torch.float64)
    assert ((iA.double() @ A.double()) - eye).abs().max().item() < 5e-4
    assert torch.equal(iA, iA.t())


R50_LAYER_CLASSES = [
    # (module ctor, input H=W, what)
    (lambda: torch.nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False), 224, 'conv1 7x7 s2: dA=147, rows=401408'),
    (lambda: torch.nn.Conv2d(64, 64, 3, padding=1, bias=False), 56, 'conv 3x3 @56: dA=576, rows=100352'),
    (lambda: torch.nn.Conv2d(256, 512, 1, stride=2, bias=False), 56, 'downsample 1x1 s2: dA=256, dG=512'),
    (lambda: torch.nn.Conv2d(512, 512, 3, padding=1, bias=False), 7, 'conv 3x3 @7: dA=4608, dG=512, rows=1568'),
    (lambda: torch.nn.Linear(2048, 1000), None, 'fc: dA=2049, dG=1000, rows=32'),
]


VGG_BERT_LAYER_CLASSES = [
    # (variant, hyper-parameters of the reference's train_*.sh, module ctor, input H=W, seq, batch, steps, what)
    ('inverse_dp', dict(lr=0.1, damping=0.03, kl_clip=0.001), lambda: torch.nn.Conv2d(512, 512, 3, padding=1), 2, None, 128, 3,
     'vgg16 conv 512->512 @2x2 +bias: dA=4609, dG=512, rows=512, damping accumulates over 3 steps'),
    ('inverse_dp', dict(lr=0.1, damping=0.03, kl_clip=0.001), lambda: torch.nn.Conv2d(256, 256, 3, padding=1), 8, None, 128, 3,
     'vgg16 conv 256->256 @8x8 +bias: dA=2305, dG=256, rows=8192'),
    ('eigen_dp', dict(lr=3e-5, damping=0.03, kl_clip=None), lambda: torch.nn.Linear(3072, 768), None, 384, 8, 2,
     'bert ffn-out 3072->768, (B,T,D) input, kl_clip=None: dA=3073, dG=768'),
    ('eigen_dp', dict(lr=3e-5, damping=0.03, kl_clip=None), lambda: torch.nn.Linear(768, 3072), None, 384, 8, 2,
     'bert ffn-in 768->3072, (B,T,D) input: dA=769, dG=3072'),
]


def _layer_step_vs_oracle(variant, hyper, ctor, hw, seq, batch, steps, what):
    """Whole step() of ONE layer at the BASELINE batch size against the oracle in fp64 (truth) and fp32
    (the reference's own noise).  Returns (our error, reference fp32 noise)."""
    import kfac_pytorch_b200 as kfac
    from kfac_pytorch_b200 import backend
    from oracle.kfac_oracle import (OracleDPKFAC, specs_from_model, synth_grads, synth_layer_inputs)
    backend.comm = None
    torch.manual_seed(0)
    mod = ctor()
    model = torch.nn.Sequential(mod)
    specs, mods = specs_from_model(model)
    spec = specs[0]
    torch.set_num_threads(16)
    hwt = None if hw is None else (hw, hw)
    res = {}
    for tag, dt in (('f64', torch.float64), ('f32', torch.float32)):
        orc_ = OracleDPKFAC([spec], variant=variant, **hyper)
        for s in range(steps):
            gen = torch.Generator().manual_seed(100 + s)
            li = synth_layer_inputs(spec, batch, hwt, seq, gen, dt)
            w, b = synth_grads(spec, gen, dt)
            ow, ob = orc_.step([[li]], [w], [b])
        res[tag] = torch.cat([ow[0].double().reshape(-1)] + ([] if ob[0] is None else [ob[0].double().reshape(-1)]))
    model.to(DEV)
    pre = kfac.get_kfac_module(variant)(model, hook_enabled=False, **hyper)
    for s in range(steps):                           # back to back: no host synchronisation between the steps
        gen = torch.Generator().manual_seed(100 + s)
        li = synth_layer_inputs(spec, batch, hwt, seq, gen)
        w, b = synth_grads(spec, gen)
        pre.m_a[mod] = li.a.to(DEV); pre.m_g[mod] = li.g.to(DEV)
        mod.weight.grad = w.to(DEV)
        if b is not None:
            mod.bias.grad = b.to(DEV)
        pre.step()
    ours = torch.cat([mod.weight.grad.double().cpu().reshape(-1)] +
                     ([] if mod.bias is None else [mod.bias.grad.double().cpu().reshape(-1)]))
    noise = ((res['f32'] - res['f64']).norm() / res['f64'].norm()).item()
    err = ((ours - res['f64']).norm() / res['f64'].norm()).item()
    line = 'FULLSIZE %-14s %-80s ours-vs-fp64 %.2e   reference-fp32-vs-fp64 %.2e' % (variant, what, err, noise)
    print('\n' + line)
    try:                                              # evidence file for profiles/ (best effort)
        import os
        os.makedirs('gpurun_out', exist_ok=True)
        with open('gpurun_out/r02_parity.txt', 'a') as fh:
            fh.write(line + '\n')
    except OSError:
        pass
    return err, noise


@pytest.mark.parametrize('cls', VGG_BERT_LAYER_CLASSES, ids=lambda c: c[7].split(':')[0].replace(' ', '_').replace(',', ''))
def test_vgg16_bert_layer_step_vs_oracle_at_full_size(cls):
    """BASELINE configs 4 and 5 at full factor size: VGG-16 `inverse_dp` (dA = 4609 / 2305 with bias; the pi-damping is
    added to the running factor IN PLACE every step, kfac_preconditioner_inv_dp.py:98-123) and BERT-base `eigen_dp`
    ((B,T,D) inputs averaged over T, kfac/utils.py:98-99; n = 3073 / 3072; kl_clip=None)."""
    variant, hyper, ctor, hw, seq, batch, steps, what = cls
    err, noise = _layer_step_vs_oracle(variant, hyper, ctor, hw, seq, batch, steps, what)
    assert err <= max(1e-4, 3 * noise), (what, err, noise)


@pytest.mark.parametrize('cls', R50_LAYER_CLASSES, ids=lambda c: c[2].split(':')[0].replace(' ', '_'))
def test_resnet50_layer_step_vs_oracle_at_full_size(cls):
    """Whole step() (factors -> eigh -> two-sided product -> KL clip) of ONE ResNet-50 layer at the BASELINE
    batch size, 2 steps back to back, against the oracle in fp64 (truth) and fp32 (the reference's own noise)."""
    ctor, hw, what = cls
    err, noise = _layer_step_vs_oracle('eigen_dp', dict(lr=0.0125, damping=0.002, kl_clip=0.001), ctor, hw, None, 32, 2, what)
    assert err <= max(1e-4, 3 * noise), (what, err, noise)
