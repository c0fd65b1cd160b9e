"""CPU study: implicit two-sided block Jacobi (the algorithm of eigh.cu) with LOW-PRECISION early sweeps.
Emulates 1xTF32 GEMMs (inputs rounded to 10 mantissa bits, fp32 accumulate) for H and the row updates in the first
k_low sweeps, then Newton-Schulz + W refresh, then accurate (fp32) sweeps."""
import sys, time
import numpy as np

def tf32(x):
    u = x.astype(np.float32).view(np.uint32)
    u = (u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)
    return u.view(np.float32)

def bf16(x):
    u = x.astype(np.float32).view(np.uint32)
    u = (u + np.uint32(0x8000)) & np.uint32(0xFFFF0000)
    return u.view(np.float32)

def mm(a, b, prec):
    if prec == 'fp32':
        return (a.astype(np.float32) @ b.astype(np.float32)).astype(np.float32)
    if prec == 'tf32':
        return (tf32(a) @ tf32(b)).astype(np.float32)
    if prec == 'bf16':
        return (bf16(a) @ bf16(b)).astype(np.float32)
    raise ValueError(prec)

def jacobi_sweep_inner(S, bipartite=True):
    """one cyclic two-sided Jacobi sweep on the 128x128 block, returns U (like bj_solve_kernel)"""
    m = S.shape[0]; h = m // 2
    S = S.astype(np.float64).copy(); U = np.eye(m)
    rounds = range(h) if bipartite else range(m - 1)
    for r in rounds:
        if bipartite:
            pairs = [(i, h + (i + r) % h) for i in range(h)]
        else:
            pairs = []
            for t in range(h):
                if t == 0: p, q = m - 1, r
                else: p, q = (r + t) % (m - 1), (r - t) % (m - 1)
                pairs.append((min(p, q), max(p, q)))
        P = np.array([p for p, _ in pairs]); Q = np.array([q for _, q in pairs])
        app, aqq, apq = S[P, P], S[Q, Q], S[P, Q]
        c = np.ones(h); s = np.zeros(h)
        act = np.abs(apq) > 1e-7 * np.sqrt(np.abs(app * aqq)) + 1e-300
        tau = (aqq[act] - app[act]) / (2 * apq[act])
        t = np.sign(tau + (tau == 0)) / (np.abs(tau) + np.sqrt(tau * tau + 1))
        c[act] = 1 / np.sqrt(t * t + 1); s[act] = t * c[act]
        # columns
        Sp, Sq = S[:, P].copy(), S[:, Q].copy()
        S[:, P] = c * Sp - s * Sq; S[:, Q] = s * Sp + c * Sq
        Up, Uq = U[:, P].copy(), U[:, Q].copy()
        U[:, P] = c * Up - s * Uq; U[:, Q] = s * Up + c * Uq
        Sp, Sq = S[P, :].copy(), S[Q, :].copy()
        S[P, :] = c[:, None] * Sp - s[:, None] * Sq; S[Q, :] = s[:, None] * Sp + c[:, None] * Sq
    return U.astype(np.float32)

def rr_pairs(p, r):
    """round-robin tournament on p (even) players, round r"""
    idx = [0] + [(r + k) % (p - 1) + 1 for k in range(p - 1)]
    return [(idx[k], idx[p - 1 - k]) for k in range(p // 2)]

def true_off(A64, Vt):
    V = Vt.astype(np.float64).T
    B = V.T @ A64 @ V
    return np.sqrt(np.sum(B * B) - np.sum(np.diag(B) ** 2)) / np.linalg.norm(A64)

def run(A, k_low, low_prec, max_sweeps=16, tol=1e-4, ns_iters=2, log=True):
    n = A.shape[0]; b = 64; p = n // b
    A64 = A.astype(np.float64)
    Wt = A.astype(np.float32).copy(); Vt = np.eye(n, dtype=np.float32)
    hist = []
    sweep = 0
    refreshed = (k_low == 0)
    while sweep < max_sweeps:
        prec = low_prec if sweep < k_low else 'fp32'
        if sweep == k_low and k_low > 0:
            # Newton-Schulz on V, W refresh
            for _ in range(ns_iters):
                G = (Vt.astype(np.float64) @ Vt.astype(np.float64).T)
                Vt = ((1.5 * np.eye(n) - 0.5 * G) @ Vt.astype(np.float64)).astype(np.float32)
            Wt = (Vt.astype(np.float64) @ A64).astype(np.float32)
            refreshed = True
            if log: print('   refresh: orth err %.2e, true off %.2e' % (np.abs(Vt.astype(np.float64) @ Vt.astype(np.float64).T - np.eye(n)).max(), true_off(A64, Vt)))
        off2 = 0.0
        for r in range(p - 1):
            for (bi, bj) in rr_pairs(p, r):
                R = np.r_[bi * b:(bi + 1) * b, bj * b:(bj + 1) * b]
                H = mm(Vt[R], Wt[R].T, prec)
                H = 0.5 * (H + H.T)
                off2 += 2 * float(np.sum(H[:b, b:].astype(np.float64) ** 2))
                U = jacobi_sweep_inner(H, bipartite=(r > 0))
                Wt[R] = mm(U.T, Wt[R], prec); Vt[R] = mm(U.T, Vt[R], prec)
        sweep += 1
        seen = np.sqrt(off2) / np.linalg.norm(A64)
        to = true_off(A64, Vt)
        hist.append((prec, seen, to))
        if log: print('  sweep %2d [%s]: off seen %.2e  true off %.2e' % (sweep, prec, seen, to))
        if prec == 'fp32' and refreshed and seen <= tol:
            break
    return sweep, hist, Vt, Wt

if __name__ == '__main__':
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    kind = sys.argv[2] if len(sys.argv) > 2 else 'wishart'
    rng = np.random.default_rng(0)
    if kind == 'wishart':
        X = rng.standard_normal((2 * n, n)); A = 0.05 * np.eye(n) + 0.95 * X.T @ X / (2 * n)
    elif kind == 'lowrank':
        X = rng.standard_normal((n // 4, n)); A = 0.05 * np.eye(n) + 0.95 * X.T @ X / (n // 4)
    else:
        Q, _ = np.linalg.qr(rng.standard_normal((n, n))); A = (Q * np.logspace(-6, 0, n)) @ Q.T
    A = (0.5 * (A + A.T)).astype(np.float32)
    for k_low, prec in [(0, 'fp32'), (3, 'tf32'), (5, 'tf32'), (7, 'tf32'), (5, 'bf16')]:
        t0 = time.time()
        print('n=%d %s  k_low=%d (%s)' % (n, kind, k_low, prec))
        sw, hist, Vt, Wt = run(A, k_low, prec)
        acc = sum(1 for h in hist if h[0] == 'fp32')
        print('  -> %d sweeps total, %d accurate, %d low;  %.0f s' % (sw, acc, sw - acc, time.time() - t0))
