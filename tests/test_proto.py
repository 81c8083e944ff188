"""CPU checks of the numpy design prototypes under scripts/proto/ (the two-stage tridiagonalisation planned next,
DESIGN.md 6b).  They are not on the product path; the tests only keep the documented algorithms reproducible."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts", "proto"))


def _spd(n, seed):
    r = np.random.default_rng(seed)
    S = r.standard_normal((n, n))
    return S @ S.T / n


def test_two_stage_reduction_keeps_the_spectrum_and_vectors():
    import two_stage_sytrd as P
    n, b = 70, 6
    S = _spd(n, 0)
    ref = np.linalg.eigvalsh(S)
    Bm, _ = P.stage1_dense_to_band(S, b)
    i, j = np.indices((n, n))
    assert np.abs(Bm[np.abs(i - j) > b]).max() < 1e-12
    np.testing.assert_allclose(np.linalg.eigvalsh(Bm), ref, atol=1e-13 * ref.max())
    d, e, tasks = P.stage2_band_to_tridiag(Bm, b)
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    np.testing.assert_allclose(np.linalg.eigvalsh(T), ref, atol=1e-13 * ref.max())
    d2, e2 = P.stage2_band_storage(P.to_band_storage(Bm, b), b)
    T2 = np.diag(d2) + np.diag(e2, 1) + np.diag(e2, -1)
    np.testing.assert_allclose(np.linalg.eigvalsh(T2), ref, atol=1e-13 * ref.max())
    lam, X = P.two_stage_with_vectors(S, b, 5)
    np.testing.assert_allclose(S @ X, X * lam, atol=1e-12 * lam[0])


def test_stage1_panels_by_choleskyqr2_and_householder_reconstruction():
    import stage1_cholqr_panels as Q
    M = Q.engine_matrix(96, 200, seed=2, k=6)
    M = 0.5 * (M + M.T)
    ref = np.linalg.eigvalsh(M)
    Bm, breakdowns, orth = Q.stage1(M, 8)
    assert breakdowns == 0 and orth < 1e-13
    i, j = np.indices(M.shape)
    assert np.abs(Bm[np.abs(i - j) > 8]).max() == 0.0
    np.testing.assert_allclose(np.linalg.eigvalsh(Bm), ref, atol=1e-13 * ref.max())


def test_bulge_chasing_sweeps_may_overlap_with_lag_two():
    import two_stage_sytrd as P
    import bulge_chase_schedule as C
    n, b = 48, 4
    Bm, _ = P.stage1_dense_to_band(_spd(n, 1), b)
    seq = C.chase(Bm, b)
    par, width = C.chase(Bm, b, lag=2, seed=3)
    np.testing.assert_array_equal(par, seq)          # bit for bit: the rule orders every conflicting pair of tasks
    assert width >= 3                                # and several sweeps are in flight
    bad, _ = C.chase(Bm, b, lag=0, seed=3)
    assert np.abs(bad - seq).max() > 1e-6            # without the lag the sweeps trample on each other


def test_varimax_polar_expansion_is_third_order():
    """The arithmetic of varimax.cu's in-loop polar factor (gram_inv_sqrt2): at the kernel's threshold (largest
    cosine 1e-3) the factor is good to ~1e-8 and the sum of singular values far below the 1e-8 stop tolerance."""
    import varimax_polar_expansion as V
    rng = np.random.default_rng(5)
    errs = []
    for eps in (1.2e-3, 1.2e-4):
        X = V.nearly_orthogonal(40, eps, rng)
        G = X.T @ X
        sn = np.sqrt(np.diag(G))
        cmax = np.abs(G / np.outer(sn, sn) - np.eye(40)).max()
        P, d = V.polar_by_expansion(X)
        u, sv, vt = np.linalg.svd(X)
        errs.append((cmax, np.abs(P - u @ vt).max(), abs(d - sv.sum()) / sv.sum()))
    (c0, e0, d0), (c1, e1, d1) = errs
    assert c0 < 1.5e-3 and e0 < 5e-8 and d0 < 1e-9
    assert e1 < e0 * (c1 / c0) ** 3 * 20             # third order in the cosines


def test_two_barrier_alpha_identity():
    """tridiag.cu (two-barrier column step): alpha = -tau/2 * w_pre^T v equals -tau^2/2 * (v^T A v - 2 (V^T v).(W^T v))
    with w_pre = tau (A v - V (W^T v) - W (V^T v)) -- the form that needs no third grid-wide reduction."""
    rng = np.random.default_rng(6)
    n, i = 60, 7
    A = rng.standard_normal((n, n)); A = A + A.T
    V, W = rng.standard_normal((n, i)), rng.standard_normal((n, i))
    v = rng.standard_normal(n); tau = 1.3
    p1, p2 = V.T @ v, W.T @ v
    w_pre = tau * (A @ v - V @ p2 - W @ p1)
    np.testing.assert_allclose(-0.5 * tau * (w_pre @ v), -0.5 * tau * tau * (v @ A @ v - 2.0 * (p1 @ p2)), rtol=1e-12)
