"""Executable fp32 model of the tridiagonalisation + divide-and-conquer eigensolver of libkfacb
(csrc/eigh_trd.cu, csrc/eigh_dc.cu): same formulas, same order of phases, numpy instead of CUDA.

It exists to pin down the algebra of the kernels before / while they are written (step structure of the
persistent tridiagonalisation kernel with its two team barriers per column, the panel bookkeeping, the secular
solver on float bit patterns, Gu-Eisenstat vectors, deflation, block-reflector back-transformation) and to
measure what accuracy the scheme delivers in fp32 on the spectra the GPU tests use.

    python tools/eigh_trd_dc_model.py [n ...]
"""
from __future__ import annotations

import sys

import numpy as np

f32 = np.float32
EPS = f32(2.0 ** -24)


# ---------------------------------------------------------------------------------------------
# (1) blocked Householder tridiagonalisation, lower variant, A = Q_H T Q_H^T
# ---------------------------------------------------------------------------------------------
def sytrd_model(A, NB=32):
    n = A.shape[0]
    T = np.array(A, dtype=f32)                    # workspace copy; row j holds v_j in T[j, j+1:] afterwards
    d = np.zeros(n, f32)
    e = np.zeros(max(n - 1, 1), f32)
    tau = np.zeros(max(n - 1, 1), f32)
    V = np.zeros((n, NB), f32)
    W = np.zeros((n, NB), f32)
    y = np.zeros(n, f32)
    x = np.zeros(n, f32)
    s = t = np.zeros(0, f32)
    vy = f32(0)
    j0 = 0

    def finish_w(jcol, k, rows, tj):
        # w = p - 0.5 tau (p.v) v,  p = tau (y - V s - W t),  p.v = tau (v.y - 2 s.t)
        p = tj * (y[rows] - V[rows, :k] @ s - W[rows, :k] @ t)
        al = f32(0.5) * tj * tj * (vy - f32(2) * f32(s @ t))
        W[rows, k] = p - al * V[rows, k]

    for j in range(n - 1):
        jp = j - j0
        # ---------------- phase A: finish w_{j-1} (rows >= j), form column j ----------------
        rows = slice(j, n)
        if jp > 0:
            finish_w(j - 1, jp - 1, rows, tau[j - 1])
        xcol = T[j, j:] - V[rows, :jp] @ W[j, :jp] - W[rows, :jp] @ V[j, :jp]
        d[j] = xcol[0]
        x[j + 1:] = xcol[1:]
        sigma2 = f32(np.sum(x[j + 2:] * x[j + 2:], dtype=f32))
        # ---------------- barrier 1; phase B: reflector, symv, panel inner products ----------------
        alpha = x[j + 1]
        if sigma2 == 0:
            tj, beta, scale = f32(0), alpha, f32(0)
        else:
            beta = -np.copysign(np.sqrt(alpha * alpha + sigma2, dtype=f32), alpha)
            tj = (beta - alpha) / beta
            scale = f32(1) / (alpha - beta)
        v = x[j + 1:] * scale
        v[0] = 1
        e[j], tau[j] = beta, tj
        V[j + 1:, jp] = v
        T[j, j + 1:] = v
        y[j + 1:] = T[j + 1:, j + 1:] @ v
        s = W[j + 1:, :jp].T @ v
        t = V[j + 1:, :jp].T @ v
        vy = f32(v @ y[j + 1:])
        # ---------------- barrier 2; panel end: finish w_j, rank-2k update of the trailing square ----------------
        if jp == NB - 1 or j == n - 2:
            rows = slice(j + 1, n)
            finish_w(j, jp, rows, tj)
            k = jp + 1
            T[j + 1:, j + 1:] -= V[rows, :k] @ W[rows, :k].T + W[rows, :k] @ V[rows, :k].T
            j0 = j + 1
    d[n - 1] = T[n - 1, n - 1]
    return T, d, e[:n - 1], tau[:n - 1]


# ---------------------------------------------------------------------------------------------
# (2) divide and conquer on the tridiagonal matrix; Zt rows = eigenvectors
# ---------------------------------------------------------------------------------------------
def leaf_bounds(n, leaf):
    L = max(1, -(-n // leaf))
    b = [(n * i) // L for i in range(L + 1)]
    return b


def secular_roots(dl, w2, rho):
    """All K roots of 1 + rho sum w2_i / (dl_i - lam) by bisection on the float bit pattern of the offset mu
    from the nearer pole.  Returns (org index, mu signed, delta[j][i] = dl_i - lam_j)."""
    K = len(dl)
    one = f32(1)
    org = np.zeros(K, np.int64)
    mu = np.zeros(K, f32)
    for j in range(K):
        if j < K - 1:
            gap = dl[j + 1] - dl[j]
            half = f32(0.5) * gap
            dj = dl - dl[j]
            gmid = one + rho * np.sum(w2 / (dj - half), dtype=f32)
            if gmid > 0:         # root in the left half: origin = lower pole, mu in (0, half]
                o, sgn, hi = j, 1, half
            else:
                o, sgn, hi = j + 1, -1, half
        else:
            o, sgn, hi = j, 1, rho * np.sum(w2, dtype=f32)
        dd = dl - dl[o]
        # g(nu) with lam = d_o + sgn * nu, nu > 0: sgn=+1: g increasing in nu from -inf; sgn=-1: decreasing from +inf
        lo_b = np.int64(0x00800000 >> 8)            # tiny positive float bits (denormal range ok)
        lo_b = np.int64(1)
        hi_b = np.int64(np.array([hi], f32).view(np.uint32)[0])
        for _ in range(34):
            if hi_b - lo_b <= 1:
                break
            mid_b = (lo_b + hi_b) // 2
            nu = np.array([mid_b], np.uint32).view(f32)[0]
            g = one + rho * np.sum(w2 / (dd - f32(sgn) * nu), dtype=f32)
            # root is where g crosses 0.  sgn=+1: g<0 -> nu too small.  sgn=-1: g>0 -> nu too small
            too_small = (g < 0) if sgn > 0 else (g > 0)
            if too_small:
                lo_b = mid_b
            else:
                hi_b = mid_b
        nu = np.array([hi_b], np.uint32).view(f32)[0]
        org[j] = o
        mu[j] = f32(sgn) * nu
    delta = np.zeros((K, K), f32)
    lam = np.zeros(K, f32)
    for j in range(K):
        delta[j] = (dl - dl[org[j]]) - mu[j]
        lam[j] = dl[org[j]] + mu[j]
    return lam, delta


def dc_merge(d1, Z1, d2, Z2, e_split, tnorm=1.0):
    """Children (d sorted ascending, Zt rows = eigenvectors) -> merged (d, Zt)."""
    n1, n2 = len(d1), len(d2)
    nm = n1 + n2
    sgn = f32(1) if e_split >= 0 else f32(-1)
    rho = f32(2) * abs(e_split)
    inv_sqrt2 = f32(1) / np.sqrt(f32(2))
    Zin = np.zeros((nm, nm), f32)
    Zin[:n1, :n1] = Z1
    Zin[n1:, n1:] = Z2
    z = np.concatenate([Z1[:, n1 - 1], sgn * Z2[:, 0]]).astype(f32) * inv_sqrt2
    dd = np.concatenate([d1, d2]).astype(f32)
    order = np.argsort(dd, kind='stable')
    # LAPACK (slaed2) uses 8 eps max(|d|max, |z|max) on a matrix SCALED to unit max-norm (sstedc); unscaled, the
    # dimensionless z has to carry the scale of the tridiagonal matrix, otherwise factors of small norm (K-FAC's
    # conv A factors are ~0.0025 I + ..) deflate far too eagerly (measured: 6e-5 instead of 3e-6)
    tol = f32(8) * EPS * max(np.max(np.abs(dd)), f32(tnorm) * np.max(np.abs(z)))
    # ---- deflation scan (sequential) ----
    nd_idx, defl = [], []          # rows of Zin
    rot = []
    if rho * np.max(np.abs(z)) <= tol:
        defl = list(order)
    else:
        pj = -1
        for jj in order:
            if rho * abs(z[jj]) <= tol:
                defl.append(jj)
                continue
            if pj < 0:
                pj = jj
                continue
            s_, c_ = z[pj], z[jj]
            tau_ = np.sqrt(c_ * c_ + s_ * s_, dtype=f32)
            tt = dd[jj] - dd[pj]
            c_, s_ = c_ / tau_, -s_ / tau_
            if abs(tt * c_ * s_) <= tol:
                z[jj] = tau_
                z[pj] = 0
                rot.append((pj, jj, c_, s_))
                # rows (pj, jj) <- [c s; -s c] applied as in LAPACK srot(Q(:,pj), Q(:,jj), c, s)
                a, b = Zin[pj].copy(), Zin[jj].copy()
                Zin[pj] = c_ * a + s_ * b
                Zin[jj] = c_ * b - s_ * a
                tnew = dd[pj] * c_ * c_ + dd[jj] * s_ * s_
                dd[jj] = dd[pj] * s_ * s_ + dd[jj] * c_ * c_
                dd[pj] = tnew
                defl.append(pj)
                pj = jj
            else:
                nd_idx.append(pj)
                pj = jj
        if pj >= 0:
            nd_idx.append(pj)
    K = len(nd_idx)
    vals = np.zeros(nm, f32)
    rows = np.zeros((nm, nm), f32)
    if K > 0:
        dl = dd[nd_idx]
        w = z[nd_idx]
        lam, delta = secular_roots(dl, w * w, rho)
        # Gu-Eisenstat: zhat_i = sign(w_i) sqrt(prod_j |delta[j][i]| / prod_{j != i} |dl_i - dl_j|)
        zh = np.zeros(K, f32)
        for i in range(K):
            pr = f32(1)
            for j in range(K):
                if j != i:
                    pr = pr * (abs(delta[j][i]) / abs(dl[i] - dl[j]))
            pr = pr * abs(delta[i][i])
            zh[i] = np.copysign(np.sqrt(pr, dtype=f32), w[i])
        Ut = zh[None, :] / delta                  # Ut[j][i]
        Ut = Ut / np.sqrt(np.sum(Ut * Ut, axis=1, dtype=f32))[:, None]
        newrows = (Ut @ Zin[nd_idx]).astype(f32)
        vals[:K] = lam
        rows[:K] = newrows
    vals[K:] = dd[defl]
    rows[K:] = Zin[defl]
    fin = np.argsort(vals, kind='stable')
    return vals[fin], rows[fin], K


def dc_model(d, e, leaf=32):
    n = len(d)
    b = leaf_bounds(n, leaf)
    dm = np.array(d, dtype=f32)
    tnorm = max(np.max(np.abs(d)), np.max(np.abs(e)) if len(e) else 0.0)       # max-norm of the tridiagonal matrix
    for k in b[1:-1]:                 # rank-one tear at every leaf boundary
        dm[k - 1] -= abs(e[k - 1])
        dm[k] -= abs(e[k - 1])
    nodes = []
    for a, c in zip(b[:-1], b[1:]):
        Tl = np.diag(dm[a:c].astype(np.float64)) + np.diag(e[a:c - 1].astype(np.float64), 1) + np.diag(e[a:c - 1].astype(np.float64), -1)
        w, Q = np.linalg.eigh(Tl)     # (the CUDA path uses the in-SMEM Jacobi kernel for the leaves)
        nodes.append((a, c, w.astype(f32), Q.T.astype(f32)))
    ks = []
    while len(nodes) > 1:
        nxt = []
        for i in range(0, len(nodes) - 1, 2):
            a, m, d1, Z1 = nodes[i]
            _, c, d2, Z2 = nodes[i + 1]
            dn, Zn, K = dc_merge(d1, Z1, d2, Z2, e[m - 1], tnorm)
            ks.append((c - a, K))
            nxt.append((a, c, dn, Zn))
        if len(nodes) % 2:
            nxt.append(nodes[-1])
        nodes = nxt
    return nodes[0][2], nodes[0][3], ks


# ---------------------------------------------------------------------------------------------
# (3) back-transformation Qt = Zt Q_H^T with block reflectors, T_b^{-1} = diag(1/tau) + striu(V^T V)
# ---------------------------------------------------------------------------------------------
def backtransform_model(Tm, tau, Zt, NBT=128):
    n = Zt.shape[0]
    Qt = np.array(Zt, dtype=f32)
    nref = n - 1
    starts = list(range(0, nref, NBT))
    for jb in reversed(starts):
        je = min(nref, jb + NBT)
        k = je - jb
        Vt = np.zeros((k, n), f32)                      # Vt[r][c] = component c of reflector jb + r
        for r in range(k):
            j = jb + r
            Vt[r, j + 1:] = Tm[j, j + 1:]
        G = Vt @ Vt.T
        tb = tau[jb:je].astype(f32)
        live = tb != 0
        Tinv = np.triu(G, 1) + np.diag(np.where(live, f32(1) / np.where(live, tb, f32(1)), f32(1)))
        Tb = np.linalg.inv(Tinv.astype(np.float64)).astype(f32)
        Tb[:, ~live] = 0
        Tb[~live, :] = 0
        Y = Qt @ Vt.T                                    # n x k
        Qt = Qt - (Y @ Tb.T) @ Vt
    return Qt


def eigh_model(A, leaf=32):
    Tm, d, e, tau = sytrd_model(A)
    lam, Zt, ks = dc_model(d, e, leaf)
    Qt = backtransform_model(Tm, tau, Zt)
    return lam, Qt.T.copy(), ks


def spectra(n, kind, rng):
    if kind == 'wishart':
        x = rng.standard_normal((3 * n + 5, n))
        return x.T @ x / x.shape[0]
    if kind == 'lowrank':
        r = max(1, n // 8)
        x = rng.standard_normal((r, n))
        return 0.05 * np.eye(n) + 0.95 * x.T @ x / r
    if kind == 'diag':
        return np.diag(np.linspace(0.0, 2.0, n))
    if kind == 'graded':
        q, _ = np.linalg.qr(rng.standard_normal((n, n)))
        return (q * np.logspace(-8, 0, n)) @ q.T
    raise ValueError(kind)


def main():
    ns = [int(a) for a in sys.argv[1:]] or [3, 10, 65, 129, 200, 513]
    rng = np.random.default_rng(0)
    for n in ns:
        for kind in ('wishart', 'lowrank', 'diag', 'graded'):
            A = spectra(n, kind, rng)
            A = 0.5 * (A + A.T)
            A32 = A.astype(f32)
            lam, Q, ks = eigh_model(A32)
            ref = np.linalg.eigvalsh(A32.astype(np.float64))
            scale = np.abs(ref).max()
            Qd, ld = Q.astype(np.float64), lam.astype(np.float64)
            ev = np.abs(ld - ref).max() / scale
            orth = np.abs(Qd.T @ Qd - np.eye(n)).max()
            rec = np.linalg.norm((Qd * ld) @ Qd.T - A32) / np.linalg.norm(A32)
            ok = ev <= 5e-6 and orth < 2e-5 and rec <= 4e-5
            print('n=%4d %-8s eig %.2e  orth %.2e  rec %.2e  nondeflated@top %s  %s'
                  % (n, kind, ev, orth, rec, ks[-1] if ks else None, 'ok' if ok else 'FAIL'))


if __name__ == '__main__':
    main()
