"""Which property of the eigendecomposition limits the accuracy of the preconditioned gradient?  conv1-like layer
(dA = 147 from 401 408 N(0,1) rows, dG = 64): v = QG (QG^T g QA / (dG dA^T + lambda)) QA^T against the fp64 truth with
(a) kfacb_eigh_batched, (b) the same Q after one Newton-Schulz step in fp64, (c) the same Q with Rayleigh-quotient
eigenvalues in fp64, (d) torch.linalg.eigh in fp32 on the GPU."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kfac_pytorch_b200.native import NativeEngine  # noqa: E402

dev = torch.device('cuda:0')
eng = NativeEngine(dev)
gen = torch.Generator(device=dev).manual_seed(0)
lam = 0.002
for (nA, rowsA, nG, rowsG) in ((147, 401408, 64, 401408), (576, 100352, 64, 100352), (256, 25088, 512, 6272)):
    a = torch.randn(rowsA, nA, device=dev, generator=gen, dtype=torch.float64)
    g = torch.randn(rowsG, nG, device=dev, generator=gen, dtype=torch.float64) / 32
    A = 0.05 * torch.eye(nA, device=dev, dtype=torch.float64) + 0.95 * (a.t() @ a) / rowsA
    G = 0.05 * torch.eye(nG, device=dev, dtype=torch.float64) + 0.95 * (g.t() @ g) * (rowsG / 32)
    grad = torch.randn(nG, nA, device=dev, generator=gen, dtype=torch.float64)

    def precond(QA, dA, QG, dG):
        QA, dA, QG, dG = QA.double(), dA.double(), QG.double(), dG.double()
        v1 = QG.t() @ grad @ QA
        return QG @ (v1 / (dG[:, None] * dA[None, :] + lam)) @ QA.t()

    dA64, QA64 = torch.linalg.eigh(A)
    dG64, QG64 = torch.linalg.eigh(G)
    truth = precond(QA64, dA64, QG64, dG64)
    err = lambda v: float((v - truth).norm() / truth.norm())
    A32, G32 = A.float().contiguous(), G.float().contiguous()
    QA, QG = torch.empty_like(A32), torch.empty_like(G32)
    dA, dG = torch.empty(nA, device=dev), torch.empty(nG, device=dev)
    eng.eigh([A32, G32], [QA, QG], [dA, dG], 1e-10, None)
    torch.cuda.synchronize()
    ns = lambda Q: Q.double() @ (1.5 * torch.eye(Q.shape[0], device=dev, dtype=torch.float64) - 0.5 * Q.double().t() @ Q.double())
    rq = lambda Q, M: ((Q.double().t() @ M.double() @ Q.double()).diagonal())
    orth = lambda Q: float(torch.linalg.matrix_norm(Q.double().t() @ Q.double() - torch.eye(Q.shape[0], device=dev, dtype=torch.float64), 2))
    dAt, QAt = torch.linalg.eigh(A32)
    dGt, QGt = torch.linalg.eigh(G32)
    print('nA=%d nG=%d  cond-ish |A||G|/lam = %.1e' % (nA, nG, float(dA64[-1] * dG64[-1] / lam)))
    print('   kfacb eigh                     err %.2e   orth2 A %.1e G %.1e  d-err A %.1e G %.1e' % (
        err(precond(QA, dA, QG, dG)), orth(QA), orth(QG),
        float((dA.double() - dA64).abs().max() / dA64[-1]), float((dG.double() - dG64).abs().max() / dG64[-1])))
    print('   + Newton-Schulz on Q (fp64)    err %.2e' % err(precond(ns(QA), dA, ns(QG), dG)))
    print('   + Rayleigh d (fp64), same Q    err %.2e' % err(precond(QA, rq(QA, A32), QG, rq(QG, G32))))
    print('   + both                         err %.2e' % err(precond(ns(QA), rq(ns(QA), A32), ns(QG), rq(ns(QG), G32))))
    print('   only A from kfacb, G fp64      err %.2e   only G from kfacb, A fp64  err %.2e' % (
        err(precond(QA, dA, QG64, dG64)), err(precond(QA64, dA64, QG, dG))))
    print('   torch.linalg.eigh fp32 (GPU)   err %.2e   orth2 A %.1e G %.1e' % (err(precond(QAt, dAt, QGt, dGt)), orth(QAt), orth(QGt)))
