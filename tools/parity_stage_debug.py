"""Stage-by-stage error of ONE full-size ResNet-50 layer step (conv1) vs the fp64 oracle: factors, eigenvalues,
preconditioned gradient (basis invariant), final gradient -- with the default eigensolver and with KFACB_EIGH_ALGO=jacobi
(set in the environment by the caller)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kfac_pytorch_b200 as kfac  # noqa: E402
from kfac_pytorch_b200 import backend  # noqa: E402
from oracle.kfac_oracle import OracleDPKFAC, specs_from_model, synth_grads, synth_layer_inputs  # noqa: E402

dev = torch.device('cuda:0')
which = sys.argv[1] if len(sys.argv) > 1 else 'conv1'
ctor, hw = {'conv1': (lambda: torch.nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False), 224),
            'c3x3': (lambda: torch.nn.Conv2d(64, 64, 3, padding=1, bias=False), 56)}[which]
hyper = dict(lr=0.0125, damping=0.002, kl_clip=0.001)
backend.comm = None
torch.manual_seed(0)
mod = ctor()
model = torch.nn.Sequential(mod)
specs, mods = specs_from_model(model)
spec = specs[0]
torch.set_num_threads(16)
rel = lambda x, y: float((x.double().cpu() - y.double()).norm() / y.double().norm())
orc = OracleDPKFAC([spec], variant='eigen_dp', **hyper)
model.to(dev)
pre = kfac.get_kfac_module('eigen_dp')(model, hook_enabled=False, **hyper)
for s in range(2):
    gen = torch.Generator().manual_seed(100 + s)
    li64 = synth_layer_inputs(spec, 32, (hw, hw), None, gen, torch.float64)
    w64, b64 = synth_grads(spec, gen, torch.float64)
    ow, ob = orc.step([[li64]], [w64], [b64])
    gen = torch.Generator().manual_seed(100 + s)
    li = synth_layer_inputs(spec, 32, (hw, hw), None, gen)
    w, b = synth_grads(spec, gen)
    pre.m_a[mod] = li.a.to(dev); pre.m_g[mod] = li.g.to(dev)
    mod.weight.grad = w.to(dev)
    pre.step()
    torch.cuda.synchronize()
    A, G = pre.m_A[mod], pre.m_G[mod]
    QA, QG, dA, dG = pre.m_QA[mod], pre.m_QG[mod], pre.m_dA[mod], pre.m_dG[mod]
    recA = rel((QA.double() * dA.double()) @ QA.double().t(), A.cpu())
    recG = rel((QG.double() * dG.double()) @ QG.double().t(), G.cpu())
    # preconditioned gradient recomputed in fp64 from OUR (Q, d) and from an fp64 eigh of OUR factors
    gm = w.double().reshape(w.shape[0], -1).to(dev)
    def precond(QA_, dA_, QG_, dG_):
        v1 = QG_.t() @ gm @ QA_
        return QG_ @ (v1 / (dG_[:, None] * dA_[None, :] + hyper['damping'])) @ QA_.t()
    v_ours = precond(QA.double(), dA.double(), QG.double(), dG.double())
    d64A, Q64A = torch.linalg.eigh(A.double()); d64G, Q64G = torch.linalg.eigh(G.double())
    v_64 = precond(Q64A, d64A, Q64G, d64G)
    print('%s step %d [%s]: A %.2e  G %.2e (vs fp64 oracle factors) | recon A %.1e G %.1e | dA %.1e dG %.1e (vs fp64 eig of OUR factors) | '
          'v(our Q,d) vs v(fp64 eig of our factors) %.2e | slab v vs oracle v %.2e | final grad vs oracle %.2e' % (
              which, s, os.environ.get('KFACB_EIGH_ALGO', 'tdc'), rel(A, orc.m_A[0]), rel(G, orc.m_G[0]), recA, recG,
              float((dA.double() - d64A).abs().max() / d64A[-1]), float((dG.double() - d64G).abs().max() / d64G[-1]),
              float((v_ours - v_64).norm() / v_64.norm()), rel(pre.m_precon_grad[mod], orc.m_precon[0]),
              rel(mod.weight.grad, ow[0])), flush=True)
