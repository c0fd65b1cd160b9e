"""CPU pin of the eigensolver's algebra (tools/eigh_trd_dc_model.py: the numpy model the CUDA kernels follow, phase by
phase): accuracy must not depend on the SCALE of the matrix.  K-FAC's conv A factors are ~0.0025 I + small; with LAPACK's
deflation tolerance applied to an UNSCALED tridiagonal matrix (LAPACK scales to unit norm first, sstedc) they were
deflated far too eagerly and reconstructed to 6e-5 only -- found on the GPU with tools/parity_stage_debug.py
(profiles/r02_eigh_accuracy.txt), reproduced and fixed here first."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tools'))
import eigh_trd_dc_model as model  # noqa: E402


def _clustered(n, scale, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((40 * n, n))
    a = x.T @ x / x.shape[0]                       # eigenvalues in a tight cluster around 1 (conv1-like)
    return (scale * (0.05 * np.eye(n) + 0.95 * a)).astype(np.float32)


@pytest.mark.parametrize('scale', [1.0, 2.5e-3, 3e2])
def test_model_accuracy_is_scale_invariant(scale):
    A = _clustered(96, scale, 5)
    lam, Q, _ = model.eigh_model(A, leaf=32)
    A64, Q64, lam64 = A.astype(np.float64), Q.astype(np.float64), lam.astype(np.float64)
    recon = np.linalg.norm((Q64 * lam64) @ Q64.T - A64) / np.linalg.norm(A64)
    orth = np.abs(Q64.T @ Q64 - np.eye(A.shape[0])).max()
    ref = np.linalg.eigvalsh(A64)
    assert recon < 5e-6, recon
    assert orth < 5e-6, orth
    assert np.abs(np.sort(lam64) - ref).max() / ref.max() < 2e-6
