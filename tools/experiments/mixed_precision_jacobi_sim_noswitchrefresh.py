import sys, time
import numpy as np
from mixed_precision_jacobi_sim import *

def run2(A, low_prec, switch_thr, tol=1e-4, max_sweeps=20, refresh_at_switch=False, log=True):
    """low precision while the off-norm seen in the previous sweep is > switch_thr; accurate afterwards; no refresh at
    the switch (unless asked); final clean-up = NS + W refresh + one accurate sweep (as eigh.cu does today)."""
    n = A.shape[0]; b = 64; p = n // b
    A64 = A.astype(np.float64)
    Wt = A.astype(np.float32).copy(); Vt = np.eye(n, dtype=np.float32)
    low = True; sweeps = []; cleanup = False
    def one_sweep(prec):
        nonlocal Wt, Vt
        off2 = 0.0
        for r in range(p - 1):
            for (bi, bj) in rr_pairs(p, r):
                R = np.r_[bi * b:(bi + 1) * b, bj * b:(bj + 1) * b]
                H = mm(Vt[R], Wt[R].T, prec); H = 0.5 * (H + H.T)
                off2 += 2 * float(np.sum(H[:b, b:].astype(np.float64) ** 2))
                U = jacobi_sweep_inner(H, bipartite=(r > 0))
                Wt[R] = mm(U.T, Wt[R], prec); Vt[R] = mm(U.T, Vt[R], prec)
        return np.sqrt(off2) / np.linalg.norm(A64)
    def refresh():
        nonlocal Wt, Vt
        G = (Vt.astype(np.float64) @ Vt.astype(np.float64).T)
        Vt = ((1.5 * np.eye(n) - 0.5 * G) @ Vt.astype(np.float64)).astype(np.float32)
        Wt = (Vt.astype(np.float64) @ A64).astype(np.float32)
    while len(sweeps) < max_sweeps:
        prec = low_prec if low else 'fp32'
        seen = one_sweep(prec)
        sweeps.append(prec)
        if log: print('  sweep %2d [%s]: seen %.2e true %.2e' % (len(sweeps), prec, seen, true_off(A64, Vt)))
        if low and seen <= switch_thr:
            low = False
            if refresh_at_switch: refresh()
            continue
        if not low and seen <= tol: break
    orth = np.abs(Vt.astype(np.float64) @ Vt.astype(np.float64).T - np.eye(n)).max()
    refresh(); seen = one_sweep('fp32'); sweeps.append('fp32*')
    V = Vt.astype(np.float64).T
    lam = np.sum(V * (A64 @ V), axis=0)
    rec = np.linalg.norm((V * lam) @ V.T - A64) / np.linalg.norm(A64)
    return sweeps, orth, true_off(A64, Vt), rec

if __name__ == '__main__':
    n = int(sys.argv[1]); kind = sys.argv[2]
    rng = np.random.default_rng(0)
    if kind == 'wishart':
        X = rng.standard_normal((2 * n, n)); A = 0.05 * np.eye(n) + 0.95 * X.T @ X / (2 * n)
    elif kind == 'lowrank':
        X = rng.standard_normal((n // 4, n)); A = 0.05 * np.eye(n) + 0.95 * X.T @ X / (n // 4)
    else:
        Q, _ = np.linalg.qr(rng.standard_normal((n, n))); A = (Q * np.logspace(-6, 0, n)) @ Q.T
    A = (0.5 * (A + A.T)).astype(np.float32)
    for prec, thr in [('fp32', 1e9), ('tf32', 3e-2), ('tf32', 1e-2), ('tf32', 3e-3), ('bf16', 3e-2), ('bf16', 1e-2)]:
        sw, orth, off, rec = run2(A, prec, thr, log=False)
        nl = sum(1 for s in sw if s == prec and prec != 'fp32')
        print('n=%d %-8s low=%s switch<=%.0e: %2d sweeps (%d low + %d accurate incl. clean-up)  orth before clean-up %.1e  final true off %.1e  recon %.1e'
              % (n, kind, prec, thr, len(sw), nl, len(sw) - nl, orth, off, rec))
