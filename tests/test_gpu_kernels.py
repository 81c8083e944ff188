"""GPU parity tests of the individual C-ABI kernels against numpy (fp64)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def D():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from xmca_b200 import device
    return device


def _rng(seed=0):
    return np.random.default_rng(seed)


@pytest.mark.parametrize("ta,tb", [(False, False), (True, False), (False, True), (True, True)])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_gemm_layouts(D, ta, tb, dt):
    r = _rng(1)
    M, N, K = 155, 162, 492
    A = r.standard_normal((K, M) if ta else (M, K)).astype(dt)
    B = r.standard_normal((N, K) if tb else (K, N)).astype(dt)
    want = (A.T if ta else A).astype(np.float64) @ (B.T if tb else B).astype(np.float64)
    got = D.to_host(D.matmul(D.to_device(A), D.to_device(B), trans_a=ta, trans_b=tb, alpha=0.5))
    np.testing.assert_allclose(got, 0.5 * want, rtol=1e-12, atol=1e-12 * np.abs(want).max())


def test_gemm_split_k_and_accumulate(D):
    r = _rng(2)
    A = r.standard_normal((40, 9000))
    B = r.standard_normal((9000, 24))
    out = D.to_device(np.ones((40, 24)))
    got = D.to_host(D.matmul(D.to_device(A), D.to_device(B), out=out, accumulate=True))
    np.testing.assert_allclose(got, A @ B + 1.0, rtol=1e-12, atol=1e-10)
    got32 = D.to_host(D.matmul(D.to_device(A.astype(np.float32)), D.to_device(B.astype(np.float32)),
                               acc="f32", out_dtype=D.f32()))
    np.testing.assert_allclose(got32, A @ B, rtol=2e-3, atol=2e-2)


@pytest.mark.parametrize("ta,tb", [(False, False), (True, False), (False, True), (True, True)])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_gemm_skinny_n_tile(D, ta, tb, dt):
    """Projections onto a few modes (N <= 64, long K, M >= 512) run on the 256 x 64 tile of the fp64 DMMA path: every
    operand layout, ragged M / N / K, mixed dtypes, split-K (chosen by `matmul`) and accumulation."""
    r = _rng(5)
    M, N, K = 1111, 50, 2600
    A = r.standard_normal((K, M) if ta else (M, K)).astype(dt)
    B = r.standard_normal((N, K) if tb else (K, N))                        # fp64 vectors against fp32 / fp64 fields
    want = (A.T if ta else A).astype(np.float64) @ (B.T if tb else B)
    got = D.to_host(D.matmul(D.to_device(A), D.to_device(B), trans_a=ta, trans_b=tb, alpha=-0.25))
    np.testing.assert_allclose(got, -0.25 * want, rtol=1e-12, atol=1e-12 * np.abs(want).max())
    out = D.to_device(np.full((M, N), 2.0))
    got = D.to_host(D.matmul(D.to_device(A), D.to_device(B), trans_a=ta, trans_b=tb, out=out, accumulate=True))
    np.testing.assert_allclose(got, want + 2.0, rtol=1e-12, atol=1e-12 * np.abs(want).max())
    got32 = D.to_host(D.matmul(D.to_device(A), D.to_device(B), trans_a=ta, trans_b=tb, out_dtype=D.f32()))
    assert got32.dtype == np.float32
    np.testing.assert_allclose(got32, want, rtol=1e-5, atol=1e-5 * np.abs(want).max())


@pytest.mark.parametrize("shape", [(128, 256, 64), (300, 200, 100), (512, 640, 1000), (96, 40, 37)])
def test_tc_gemm_matches_fp64(D, shape):
    """3xTF32 tcgen05 product vs fp64 numpy: error at the fp32 level."""
    M, N, K = shape
    r = _rng(3)
    A = r.standard_normal((K, M)).astype(np.float32)      # time-major fields (T x S)
    B = r.standard_normal((K, N)).astype(np.float32)
    C, frob2 = D.cov_gemm_tc(D.to_device(A), D.to_device(B), 0.25)
    want = 0.25 * A.astype(np.float64).T @ B.astype(np.float64)
    got = D.to_host(C).astype(np.float64)
    scale = np.sqrt(K) * 0.25
    assert np.abs(got - want).max() < 2e-6 * scale * 4
    np.testing.assert_allclose(D.to_host(frob2)[0], (got ** 2).sum(), rtol=1e-6)


def test_tc_gemm_planted_structure(D):
    """Catches k-slab / swizzle / descriptor mistakes: structured operands."""
    M, N, K = 256, 512, 160
    A = np.zeros((K, M), np.float32)
    B = np.zeros((K, N), np.float32)
    for k in range(K):
        A[k, (3 * k) % M] = 1.0 + k
        B[k, (7 * k + 1) % N] = 2.0 - 0.01 * k
    C, _ = D.cov_gemm_tc(D.to_device(A), D.to_device(B), 1.0)
    want = A.astype(np.float64).T @ B.astype(np.float64)
    np.testing.assert_allclose(D.to_host(C), want, rtol=1e-6, atol=1e-4)


@pytest.mark.parametrize("n,m", [(150, 200), (64, 64), (33, 500), (200, 200)])
def test_jacobi_svd(D, n, m):
    r = _rng(4)
    X = r.standard_normal((n, m)) * np.logspace(0, -3, n)[:, None]
    Xr, sig, Jt, sweeps = D.jacobi_svd(D.to_device(X))
    s = D.to_host(sig)
    order = np.argsort(-s)[:min(n, m)]
    want = np.linalg.svd(X, compute_uv=False)
    np.testing.assert_allclose(s[order], want[:order.size], rtol=1e-10, atol=1e-13 * want[0])
    J = D.to_host(Jt)[:, :n]
    np.testing.assert_allclose(J[:n] @ J[:n].T, np.eye(n), atol=1e-12)
    # X = J Sigma U^T  with rows of Xr = sigma_j u_j^T and columns of J = rows of Jt
    rec = J[:n].T @ D.to_host(Xr)[:n]
    np.testing.assert_allclose(rec, X, atol=1e-12 * want[0])
    assert sweeps <= 20


def test_jacobi_symmetric_psd_gives_eigenpairs(D):
    r = _rng(5)
    Y = r.standard_normal((90, 300))
    Y -= Y.mean(axis=0)
    G = Y @ Y.T
    Gr, lam, Jt, _ = D.jacobi_svd(D.to_device(G))
    lam, J = D.to_host(lam), D.to_host(Jt)
    order = np.argsort(-lam)[:90]
    w = np.linalg.eigvalsh(G)[::-1]
    np.testing.assert_allclose(lam[order], np.maximum(w, 0), atol=1e-10 * w[0])
    U = J[order][:, :90]
    np.testing.assert_allclose(U @ G @ U.T, np.diag(lam[order]), atol=1e-9 * w[0])


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n,p", [(317, 8), (5000, 20), (2048, 50), (100, 3)])
def test_varimax_matches_oracle(D, dt, n, p):
    from oracle import mca_oracle as orc
    r = _rng(6)
    L = (r.standard_normal((n, p)) * (1 + 5 * (r.random((n, p)) < 0.1))).astype(dt)
    L *= np.linspace(2.0, 0.5, p)
    try:
        Bo, Ro, it_o = orc.varimax(L, tol=1e-8)
    except orc.NotConverged:
        pytest.skip("oracle did not converge on this input")
    B, R, it = D.varimax(D.to_device(L), tol=1e-8)
    R, B = D.to_host(R), D.to_host(B)
    np.testing.assert_allclose(R.T @ R, np.eye(p), atol=1e-10)
    assert abs(it - it_o) <= max(3, it_o // 20)
    np.testing.assert_allclose(B, Bo, atol=2e-5 * np.abs(Bo).max())


def test_elementwise_helpers(D):
    r = _rng(7)
    X = r.standard_normal((70, 45)).astype(np.float32) + 3.0
    Xd = D.to_device(X.copy())
    mean = D.to_host(D.center_columns(Xd))
    np.testing.assert_allclose(mean, X.astype(np.float64).mean(axis=0), rtol=1e-6)
    np.testing.assert_allclose(D.to_host(Xd), X - X.mean(axis=0), atol=1e-5)
    np.testing.assert_allclose(D.to_host(D.transpose(D.to_device(X), out_dtype=D.f64())), X.T.astype(np.float64))
    np.testing.assert_allclose(D.to_host(D.col_sumsq(D.to_device(X), 10, 60)),
                               (X[10:60].astype(np.float64) ** 2).sum(axis=0), rtol=1e-12)
    np.testing.assert_allclose(D.to_host(D.row_sumsq(D.to_device(X))),
                               (X.astype(np.float64) ** 2).sum(axis=1), rtol=1e-12)
    rs = r.random(70)
    np.testing.assert_allclose(D.to_host(D.col_absmax(D.to_device(X), row_scale=D.to_device(rs))),
                               np.abs(X * rs[:, None]).max(axis=0), rtol=1e-7)
    idx = np.array([5, 0, 69, 5], dtype=np.int64)
    got = D.to_host(D.gather_rows(D.to_device(X), D.to_device(idx), row_scale=D.to_device(np.array([1., 2., 3., 4.]))))
    np.testing.assert_allclose(got, X[idx] * np.array([1, 2, 3, 4])[:, None], rtol=1e-6)


def test_philox_normal_is_standard_and_geometry_independent(D):
    t = D.torch()
    X = D.fill_normal(D.empty((1000, 501), t.float64), seed=42, stream_id=3)
    Y = D.fill_normal(D.empty((501 * 1000 // 3, 3), t.float64), seed=42, stream_id=3)
    x, y = D.to_host(X), D.to_host(Y)
    assert np.array_equal(x.ravel(), y.ravel())           # counter-based: same stream, any shape
    assert abs(x.mean()) < 5e-3 and abs(x.std() - 1) < 5e-3
    assert abs(((x - x.mean()) ** 4).mean() - 3) < 0.05
    Z = D.to_host(D.fill_normal(D.empty((1000, 501), t.float64), seed=42, stream_id=4))
    assert abs(np.corrcoef(x.ravel(), Z.ravel())[0, 1]) < 5e-3


@pytest.mark.parametrize("n,nrhs", [(64, 5), (200, 37), (515, 515), (1000, 64)])
def test_cholesky_and_triangular_solve(D, n, nrhs):
    """Blocked fp64 Cholesky + L^T W = R against numpy."""
    r = _rng(11)
    X = r.standard_normal((n, 2 * n))
    G = X @ X.T
    Ld, inv = D.cholesky(D.to_device(G.copy()))
    Lh = D.to_host(Ld)
    want = np.linalg.cholesky(G)
    np.testing.assert_allclose(Lh, want, rtol=0, atol=1e-11 * np.abs(want).max())
    assert np.all(np.triu(Lh, 1) == 0.0)
    R = r.standard_normal((n, nrhs))
    W = D.to_host(D.trsm_lt(Ld, inv, D.to_device(R.copy())))
    np.testing.assert_allclose(want.T @ W, R, atol=1e-9 * np.abs(R).max() * n)


def test_cholesky_rejects_indefinite(D):
    G = np.eye(130)
    G[77, 77] = -1.0
    with pytest.raises(np.linalg.LinAlgError):
        D.cholesky(D.to_device(G))


@pytest.mark.parametrize("case", ["diagonal", "toeplitz", "decoupled", "graded", "clustered", "wilkinson", "n1", "n2", "n9",
                                  "negative", "tiny_scale", "huge_scale"])
def test_bisection_on_special_tridiagonals(D, case, monkeypatch):
    """`xmca_stebz` (Sturm sequence in product form with renormalisation, pivmin rule applied on redo) against LAPACK on
    tridiagonals that exercise exact zeros in the sequence, decoupled blocks, 300 orders of magnitude of grading,
    clusters, every remainder length of the 8-step chunks, and extreme scalings; both recurrence forms."""
    r = _rng(17)
    n = 257
    if case == "diagonal":
        d, e = r.standard_normal(n), np.zeros(n - 1)
    elif case == "toeplitz":                         # zero diagonal: p_j vanishes exactly at x = 0 for odd j
        d, e = np.zeros(n), np.ones(n - 1)
    elif case == "decoupled":
        d, e = r.standard_normal(n), r.standard_normal(n - 1)
        e[::7] = 0.0
    elif case == "graded":
        d = np.logspace(0, -150, n)
        e = 0.1 * np.sqrt(d[:-1] * d[1:])
    elif case == "clustered":
        d = np.concatenate([np.full(100, 1.0), np.full(100, 1.0 + 1e-13), r.uniform(0, 2, 57)])
        e = np.full(n - 1, 1e-14)
    elif case == "wilkinson":
        d, e = np.abs(np.arange(n) - n // 2).astype(float), np.ones(n - 1)
    elif case in ("n1", "n2", "n9"):
        n = int(case[1:])
        d, e = r.standard_normal(n), r.standard_normal(max(n - 1, 1))[:n - 1]
    elif case == "negative":
        d, e = -np.abs(r.standard_normal(n)) - 3.0, 0.3 * r.standard_normal(n - 1)
    else:
        sc = 1e-140 if case == "tiny_scale" else 1e140
        d, e = sc * r.standard_normal(n), sc * r.standard_normal(n - 1)
    Tm = np.diag(d) + (np.diag(e, 1) + np.diag(e, -1) if n > 1 else 0.0)
    ref = np.linalg.eigvalsh(Tm)[::-1]
    tol = 4e-15 * n * max(np.abs(ref).max(), 1e-300)
    epad = np.concatenate([e, [0.0]])                # device arrays: d (n), e (n - 1, one spare entry)
    for mode in ("prod", "quot"):
        monkeypatch.setenv("XMCA_STEBZ", mode)
        w = D.to_host(D.stebz(D.to_device(d.copy()), D.to_device(epad.copy())))
        assert w.shape == (n,) and np.all(np.diff(w) <= 0.0)
        np.testing.assert_allclose(w, ref, atol=tol, rtol=0, err_msg=mode)


@pytest.mark.parametrize("n", [5, 64, 200, 777, 1500, 4700])
def test_tridiagonal_eigensolver(D, n):
    """sytrd + stebz give all eigenvalues; stein + ormtr the leading eigenvectors (also inside
    exactly degenerate clusters)."""
    r = _rng(n)
    Q, _ = np.linalg.qr(r.standard_normal((n, n)))
    lam = np.sort(r.uniform(0.5, 2.0, n) * np.logspace(0, -3, n))[::-1].copy()
    if n >= 64:
        lam[3] = lam[2]                      # an exactly double eigenvalue
        lam[10:14] = lam[10]                 # and a fourfold one
    S = (Q * lam) @ Q.T
    S = 0.5 * (S + S.T)
    Sd = D.to_device(S)
    d, e, tau = D.sytrd(Sd)
    w = D.to_host(D.stebz(d, e))
    ref = np.linalg.eigvalsh(S)[::-1]
    np.testing.assert_allclose(w, ref, atol=5e-14 * n * ref[0])
    # tridiagonal really is similar to S: compare with numpy on (d, e)
    dh, eh = D.to_host(d), D.to_host(e)[:n - 1]
    Tm = np.diag(dh) + np.diag(eh, 1) + np.diag(eh, -1)
    np.testing.assert_allclose(np.linalg.eigvalsh(Tm)[::-1], ref, atol=5e-14 * n * ref[0])
    m = min(n, 40)
    gap = 1e-6 * w[0]
    starts = [0] + [i for i in range(1, m) if w[i - 1] - w[i] > gap] + [m]
    Z = D.stein(d, e, w[:m], np.asarray(starts), w[0])
    Th = Tm @ D.to_host(Z).T
    np.testing.assert_allclose(Th, D.to_host(Z).T * w[:m], atol=1e-11 * ref[0])      # T z = lambda z
    X = D.to_host(D.ormtr(Sd, tau, Z)).T                                             # n x m
    np.testing.assert_allclose(X.T @ X, np.eye(m), atol=1e-9)
    np.testing.assert_allclose(S @ X, X * w[:m], atol=1e-10 * ref[0])


@pytest.mark.parametrize("n", [130, 1500, 4700])
def test_batched_tridiagonalisation(D, n):
    """xmca_sytrd_batched (two problems, half of the SMs each) against the single-problem call and numpy."""
    r = _rng(n + 1)
    mats = []
    for k in range(2):
        Q, _ = np.linalg.qr(r.standard_normal((n, n)))
        lam = np.sort(r.uniform(0.5, 2.0, n) * np.logspace(0, -3 - k, n))[::-1]
        S = (Q * lam) @ Q.T
        mats.append(0.5 * (S + S.T))
    Sp = D.to_device(np.stack(mats))
    d, e, tau = D.sytrd_pair(Sp)
    for k in range(2):
        ref = np.linalg.eigvalsh(mats[k])[::-1]
        w = D.to_host(D.stebz(d[k], e[k, :n - 1]))
        np.testing.assert_allclose(w, ref, atol=5e-14 * n * ref[0])
        d1, e1, _ = D.sytrd(D.to_device(mats[k]))
        w1 = D.to_host(D.stebz(d1, e1))
        np.testing.assert_allclose(w, w1, atol=2e-14 * n * ref[0])
        # reflectors of the batched call reproduce eigenvectors of the ORIGINAL matrix
        m = 6
        starts = np.arange(m + 1)
        Z = D.stein(d[k], e[k, :n - 1], w[:m], starts, w[0])
        X = D.to_host(D.ormtr(Sp[k], tau[k], Z)).T
        np.testing.assert_allclose(mats[k] @ X, X * w[:m], atol=1e-10 * ref[0])


def test_gemm_structure_flags(D):
    """Symmetric-result and triangular-operand shortcuts give the same numbers as the plain product."""
    r = _rng(5)
    n, K = 300, 700
    X = r.standard_normal((n, K))
    G = D.to_host(D.matmul(D.to_device(X), D.to_device(X), trans_b=True, symmetric=True, alpha=0.5))
    np.testing.assert_allclose(G, 0.5 * X @ X.T, rtol=1e-12, atol=1e-11)
    Lm = np.tril(r.standard_normal((n, n)))
    Gs = X @ X.T
    W = D.to_host(D.matmul(D.to_device(Gs), D.to_device(Lm), b_lower=True))
    np.testing.assert_allclose(W, Gs @ Lm, rtol=1e-12, atol=1e-9)
    S = D.to_host(D.matmul(D.to_device(Lm), D.to_device(W), trans_a=True, symmetric=True, a_lower_t=True))
    np.testing.assert_allclose(S, Lm.T @ Gs @ Lm, rtol=1e-11, atol=1e-8)
    # accumulate into a symmetric matrix (the rank-2k update of the tridiagonalisation)
    V, Wp = r.standard_normal((n, 128)), r.standard_normal((n, 128))
    VW, WV = np.hstack([V, Wp]), np.hstack([Wp, V])
    out = D.to_device(Gs.copy())
    D.matmul(D.to_device(VW), D.to_device(WV), trans_b=True, alpha=-1.0, out=out, accumulate=True, symmetric=True)
    np.testing.assert_allclose(D.to_host(out), Gs - VW @ WV.T, rtol=1e-11, atol=1e-8)


@pytest.mark.parametrize("T,dt", [(64, np.float64), (77, np.float64), (492, np.float32), (250, np.float32)])
def test_hilbert_operators_match_analytic_signal(D, T, dt):
    """Device analytic signal (circulant Hilbert operator) and one-sided spectrum operator against the
    FFT definition of scipy.signal.hilbert (oracle.analytic_signal), even and odd lengths."""
    from oracle import mca_oracle as orc
    r = _rng(T)
    X = r.standard_normal((T, 37)).astype(dt)
    X -= X.mean(axis=0)
    Z = orc.analytic_signal(X)
    Xd = D.to_device(X)
    Y, _ = D.apply_time_operator(D.hilbert_matrix(T, Xd.dtype), Xd)
    tol = 2e-5 if dt == np.float32 else 1e-11
    np.testing.assert_allclose(D.to_host(Y), Z.imag, atol=tol * np.abs(Z).max())
    # spectrum operator: Z_A^H Z_B is preserved, with half the rows
    ZZ, _ = D.apply_time_operator(D.dft_matrix(T, Xd.dtype), Xd)
    zz = D.to_host(ZZ).astype(np.float64)
    Tp = T // 2
    assert zz.shape == (2 * Tp, 37)
    Zh = zz[:Tp] + 1j * zz[Tp:]
    Zc = Z.astype(np.complex128)
    np.testing.assert_allclose(Zh.conj().T @ Zh, Zc.conj().T @ Zc, atol=(5e-5 if dt == np.float32 else 1e-10) * T)
    E = D.to_host(D.embed_complex(ZZ))
    np.testing.assert_array_equal(E[:Tp, 37:], -E[Tp:, :37])
    np.testing.assert_array_equal(E[:Tp, :37], E[Tp:, 37:])


@pytest.mark.parametrize("shape", [(300, 1000), (512, 4096), (1000, 777), (130, 20000)])
def test_tensor_core_gram_fp64_accumulation(D, shape):
    """G = X X^T of fp32 data on tcgen05 (3xTF32) with split TMEM accumulators and fp64 chunk sums:
    ~1e-7 of the diagonal scale, also for all-positive data (worst case for the accumulator truncation)."""
    n, K = shape
    r = _rng(n + K)
    for kind in ("gauss", "positive"):
        X = r.standard_normal((n, K)).astype(np.float32)
        if kind == "positive":
            X = np.abs(X) + np.float32(0.5)
        G = D.to_host(D.gram_tc(D.to_device(X), alpha=0.5))
        want = 0.5 * X.astype(np.float64) @ X.astype(np.float64).T
        scale = np.sqrt(np.outer(np.diag(want), np.diag(want)))
        assert np.abs(G - want).max() / scale.max() < 4e-7
        assert np.abs((G - want) / scale).max() < 4e-7
        assert np.abs(G - G.T).max() <= 1e-8 * scale.max()       # diagonal tiles compute both triangles
