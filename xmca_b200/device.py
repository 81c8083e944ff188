"""Thin device-level wrappers: torch.cuda tensors in, C-ABI calls out.

torch is used for allocation, H2D/D2H copies and stream handles only; every
arithmetic operation on device data goes through ``libxmca_b200.so``.
All matrices are 2-D, row-major, unit inner stride (``stride(0)`` is the
leading dimension).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L

_SM = None
last_varimax_stats = None      # device tensor: [iterations, converged, d, svd sweeps, 6 x phase clocks]


def torch():
    return L._torch()


def sm_count() -> int:
    global _SM
    if _SM is None:
        t = torch()
        _SM = t.cuda.get_device_properties(t.cuda.current_device()).multi_processor_count
    return _SM


def dev():
    t = torch()
    return t.device("cuda", t.cuda.current_device())


def empty(shape, dtype):
    return torch().empty(shape, dtype=dtype, device=dev())


def zeros(shape, dtype):
    return torch().zeros(shape, dtype=dtype, device=dev())


def f64():
    return torch().float64


def f32():
    return torch().float32


def to_device(a: np.ndarray):
    t = torch()
    return t.from_numpy(np.ascontiguousarray(a)).to(dev(), non_blocking=False)


def to_host(x) -> np.ndarray:
    return x.detach().cpu().numpy()


def _ld(x):
    assert x.dim() == 2 and (x.shape[1] == 1 or x.stride(1) == 1), "row-major 2-D tensor expected"
    return x.stride(0) if x.shape[0] > 1 else max(x.shape[1], x.stride(0))


# ---------------------------------------------------------------------- GEMM
def matmul(A, B, trans_a=False, trans_b=False, alpha=1.0, out=None, out_dtype=None,
           acc="f64", accumulate=False, symmetric=False, a_lower_t=False, b_lower=False):
    """out = alpha * op(A) @ op(B) (+ out).  fp64 accumulation by default.
    symmetric: the result is symmetric (only the lower tiles are computed, then mirrored);
    a_lower_t: op(A) = L^T with L lower triangular (needs trans_a); b_lower: op(B) = L lower
    triangular (no trans_b): the all-zero part of the k range is skipped."""
    lib = L.load()
    t = torch()
    if trans_a:
        K, M = A.shape
    else:
        M, K = A.shape
    if trans_b:
        N, Kb = B.shape
    else:
        Kb, N = B.shape
    if K != Kb:
        raise ValueError("matmul: inner dimensions differ (%d vs %d)" % (K, Kb))
    if out is None:
        if accumulate:
            raise ValueError("matmul: accumulate needs `out`")
        out = empty((M, N), out_dtype or t.float64)
    acc_code = L.F64 if acc == "f64" else L.F32
    tiles = ((M + 127) // 128) * ((N + 127) // 128)
    split = 1
    flags = (L.GEMM_SYMMETRIC if symmetric else 0) | (L.GEMM_A_LOWER_T if a_lower_t else 0) | \
            (L.GEMM_B_LOWER if b_lower else 0)
    if flags and (M != N or (a_lower_t and not trans_a) or (b_lower and trans_b)):
        raise ValueError("matmul: structure flags need a square result and L^T @ . / . @ L operands")
    if not flags and tiles < sm_count() and K >= 2048:
        split = int(min(max(1, (2 * sm_count()) // tiles), K // 512, 64))
    ws, ws_bytes = None, 0
    if split > 1:
        ws_bytes = lib.xmca_gemm_workspace_bytes(M, N, split, acc_code)
        ws = empty((ws_bytes,), t.uint8)
    rc = lib.xmca_gemm_ex(0 if trans_a else 1, 1 if trans_b else 0, M, N, K, float(alpha),
                          L.ptr(A), L.dtype_code(A), _ld(A), L.ptr(B), L.dtype_code(B), _ld(B),
                          L.ptr(out), L.dtype_code(out), _ld(out), 1 if accumulate else 0,
                          acc_code, split, L.ptr(ws), ws_bytes, flags, L.stream_ptr())
    L.check(rc, "xmca_gemm")
    return out


# ----------------------------------------------------- tensor-core covariance
def _pad4(n):
    return (n + 3) // 4 * 4


def split_tf32(X, transpose=False):
    """(hi, lo) TF32 planes of X (or X^T), K-major, pitch padded to 4 floats."""
    lib = L.load()
    rows, cols = X.shape
    orow, ocol = (cols, rows) if transpose else (rows, cols)
    ldo = _pad4(ocol)
    hi = zeros((orow, ldo), f32()) if ldo != ocol else empty((orow, ldo), f32())
    lo = zeros((orow, ldo), f32()) if ldo != ocol else empty((orow, ldo), f32())
    rc = lib.xmca_split_tf32(L.ptr(X), L.dtype_code(X), rows, cols, _ld(X), 1 if transpose else 0,
                             L.ptr(hi), L.ptr(lo), ldo, L.stream_ptr())
    L.check(rc, "xmca_split_tf32")
    return hi, lo, ocol


def tc_gemm_nt(Ahi, Alo, Bhi, Blo, K, alpha=1.0, out=None, frob2=None):
    """out[M,N] = alpha * A[M,K] @ B[N,K]^T on tcgen05 (3xTF32); planes from split_tf32."""
    lib = L.load()
    M, N = Ahi.shape[0], Bhi.shape[0]
    if out is None:
        out = empty((M, N), f32())
    rc = lib.xmca_tc_gemm_nt(M, N, K, float(alpha), L.ptr(Ahi), L.ptr(Alo), Ahi.stride(0),
                             L.ptr(Bhi), L.ptr(Blo), Bhi.stride(0), L.ptr(out), _ld(out),
                             L.ptr(frob2), L.stream_ptr())
    L.check(rc, "xmca_tc_gemm_nt")
    return out


def gram_tc(X, alpha=1.0):
    """G = alpha * X X^T (rows x rows, fp64, symmetric) of an fp32 matrix on the tensor cores
    (3xTF32, fp64 chunk accumulation: xmca_tc_gemm_nt_f64)."""
    lib = L.load()
    hi, lo, K = split_tf32(X)                                  # rows x K planes, K-major
    n = X.shape[0]
    G = empty((n, n), f64())
    rc = lib.xmca_tc_gemm_nt_f64(n, n, K, float(alpha), L.ptr(hi), L.ptr(lo), hi.stride(0), L.ptr(hi), L.ptr(lo),
                                 hi.stride(0), L.ptr(G), n, 1, L.stream_ptr())
    L.check(rc, "xmca_tc_gemm_nt_f64")
    return G


def cov_gemm_tc(A, B, alpha):
    """C = alpha * A^T B (S1 x S2, fp32) from time-major fp32 fields, tensor-core path.
    Returns (C, frob2 tensor)."""
    ahi, alo, K = split_tf32(A, transpose=True)
    if B is A:
        bhi, blo = ahi, alo
    else:
        bhi, blo, _ = split_tf32(B, transpose=True)
    frob2 = zeros((1,), f64())
    Cm = tc_gemm_nt(ahi, alo, bhi, blo, K, alpha=alpha, frob2=frob2)
    return Cm, frob2


# -------------------------------------------------------------------- Jacobi
def jacobi_svd(X, want_v=True, max_sweeps=60, tol=0.0):
    """One-sided Jacobi on the ROWS of X (n x m, fp64, row-major == column-major m x n).

    Returns (Xr, sigma, Jt, sweeps): Xr[j] = sigma_j * u_j (rows, rotated in a
    padded copy), sigma (n_pad,), Jt (n_pad x n_pad) with Jt[j] = right singular
    vector j (None if want_v is False).  Rows are NOT sorted."""
    lib = L.load()
    t = torch()
    n, m = X.shape
    n_pad = int(lib.xmca_jacobi_padded_cols(n))
    Xp = zeros((n_pad, m), t.float64)
    rc = lib.xmca_scale_copy(L.ptr(X), L.dtype_code(X), _ld(X), L.ptr(Xp), L.F64, m, n, m, None, None,
                             L.stream_ptr())
    L.check(rc, "xmca_scale_copy")
    Jt = empty((n_pad, n_pad), t.float64) if want_v else None
    sigma = empty((n_pad,), t.float64)
    ws_bytes = lib.xmca_jacobi_workspace_bytes(m, n)
    ws = empty((ws_bytes,), t.uint8)
    sweeps, off = C.c_int(0), C.c_double(0.0)
    rc = lib.xmca_jacobi_svd(m, n, L.ptr(Xp), m, L.ptr(Jt), n_pad, L.ptr(sigma), max_sweeps, float(tol),
                             C.byref(sweeps), C.byref(off), L.ptr(ws), ws_bytes, L.stream_ptr())
    L.check(rc, "xmca_jacobi_svd")
    return Xp, sigma, Jt, sweeps.value


# ------------------------------------------------------------------ Cholesky
def cholesky(G, min_pivot=0.0):
    """In-place lower Cholesky factor of the fp64 SPD matrix G (n x n).
    Returns (L (= G, overwritten), invdiag).  Raises LinAlgError if a pivot is not
    above ``min_pivot`` (G not numerically SPD)."""
    lib = L.load()
    t = torch()
    n = G.shape[0]
    assert G.dtype == t.float64 and G.shape[1] == n
    invdiag = empty((lib.xmca_cholesky_invdiag_bytes(n) // 8,), t.float64)
    ws_bytes = lib.xmca_cholesky_workspace_bytes(n)
    ws = empty((ws_bytes,), t.uint8)
    info = C.c_int(0)
    rc = lib.xmca_cholesky(n, L.ptr(G), _ld(G), L.ptr(invdiag), float(min_pivot), C.byref(info), L.ptr(ws),
                           ws_bytes, L.stream_ptr())
    L.check(rc, "xmca_cholesky")
    return G, invdiag


def trsm_lt(Lmat, invdiag, R):
    """Solve Lmat^T W = R in place (R: n x nrhs fp64, row-major); returns R."""
    lib = L.load()
    t = torch()
    n, nrhs = R.shape
    assert R.dtype == t.float64 and Lmat.shape[0] == n
    ws_bytes = lib.xmca_trsm_workspace_bytes(n, nrhs)
    ws = empty((ws_bytes,), t.uint8)
    rc = lib.xmca_trsm_lt(n, nrhs, L.ptr(Lmat), _ld(Lmat), L.ptr(invdiag), L.ptr(R), _ld(R), L.ptr(ws), ws_bytes,
                          L.stream_ptr())
    L.check(rc, "xmca_trsm_lt")
    return R


# ------------------------------------------------- tridiagonal eigen-solver
def sytrd_max_n():
    return int(L.load().xmca_sytrd_max_n())


def sytrd(S):
    """Householder tridiagonalisation of the symmetric fp64 matrix S (n x n, both triangles,
    DESTROYED: its rows then hold the reflectors).  Returns (d, e, tau) device vectors."""
    lib = L.load()
    t = torch()
    n = S.shape[0]
    assert S.dtype == t.float64 and S.shape[1] == n
    d, e, tau = empty((n,), t.float64), zeros((max(n - 1, 1),), t.float64), empty((n,), t.float64)
    ws_bytes = lib.xmca_sytrd_workspace_bytes(n)
    ws = empty((ws_bytes,), t.uint8)
    rc = lib.xmca_sytrd(n, L.ptr(S), _ld(S), L.ptr(d), L.ptr(e), L.ptr(tau), L.ptr(ws), ws_bytes, L.stream_ptr())
    L.check(rc, "xmca_sytrd")
    return d, e, tau


def sytrd_pair(S2):
    """Two tridiagonalisations of the same size in one batched call (S2: 2 x n x n fp64, contiguous,
    DESTROYED).  Returns (d, e, tau) as 2 x n device tensors (e: the first n - 1 entries of each row)."""
    lib = L.load()
    t = torch()
    assert S2.dtype == t.float64 and S2.dim() == 3 and S2.shape[0] == 2 and S2.shape[1] == S2.shape[2] and S2.is_contiguous()
    n = S2.shape[1]
    d, e, tau = empty((2, n), t.float64), zeros((2, n), t.float64), empty((2, n), t.float64)
    ws_bytes = 2 * lib.xmca_sytrd_workspace_bytes(n)
    ws = empty((ws_bytes,), t.uint8)
    rc = lib.xmca_sytrd_batched(n, 2, L.ptr(S2), n, n * n, L.ptr(d), L.ptr(e), L.ptr(tau), n, L.ptr(ws), ws_bytes,
                                L.stream_ptr())
    L.check(rc, "xmca_sytrd_batched")
    return d, e, tau


def sytrd2(S, want_vectors=True, sync=True):
    """Two-stage tridiagonalisation (dense -> band 64 on the DMMA pipe -> tridiagonal by bulge chasing) of the
    symmetric fp64 matrix S (n x n, both triangles, DESTROYED: it then holds both reflector sets).
    Returns (d, e, tfac).  Raises LinAlgError if a panel factorisation broke down.
    sync=False: only enqueues on the current stream and returns (d, e, tfac, pending); call `sytrd2_check(pending)`
    after the stream has been synchronised (it raises LinAlgError on a breakdown)."""
    lib = L.load()
    t = torch()
    n = S.shape[0]
    assert S.dtype == t.float64 and S.shape[1] == n
    d, e = empty((n,), t.float64), zeros((max(n - 1, 1),), t.float64)
    tfac = empty((lib.xmca_sytrd2_tfac_bytes(n) // 8,), t.float64)
    ws_bytes = lib.xmca_sytrd2_workspace_bytes(n)
    ws = empty((ws_bytes,), t.uint8)
    flags = (1 if want_vectors else 0) | (0 if sync else 8)
    rc = lib.xmca_sytrd2(n, L.ptr(S), _ld(S), L.ptr(d), L.ptr(e), L.ptr(tfac), flags,
                         L.ptr(ws), ws_bytes, L.stream_ptr())
    L.check(rc, "xmca_sytrd2")
    if sync:
        return d, e, tfac
    off = int(lib.xmca_sytrd2_info_offset(n))
    return d, e, tfac, ws[off:off + 4]


def sytrd2_check(pending):
    """Breakdown flag of an asynchronous `sytrd2` (reads 4 bytes from the device: synchronises)."""
    if int(pending.view(torch().int32).item()) != 0:
        raise np.linalg.LinAlgError("xmca_sytrd2: panel factorisation broke down")


def ormtr2(S_reflectors, tfac, Z):
    """Rows of Z <- Q row (in place), Q = Q1 Q2 from `sytrd2`."""
    lib = L.load()
    n = S_reflectors.shape[0]
    ws_bytes = lib.xmca_ormtr2_workspace_bytes(n, Z.shape[0])
    ws = empty((ws_bytes,), torch().uint8)
    rc = lib.xmca_ormtr2(n, L.ptr(S_reflectors), _ld(S_reflectors), L.ptr(tfac), Z.shape[0], L.ptr(Z), _ld(Z),
                         L.ptr(ws), ws_bytes, L.stream_ptr())
    L.check(rc, "xmca_ormtr2")
    return Z


def stebz(d, e):
    """All eigenvalues of the symmetric tridiagonal (d, e), descending (device fp64)."""
    lib = L.load()
    n = d.shape[0]
    w = empty((n,), f64())
    scratch = empty((2 * n + 8,), f64())
    rc = lib.xmca_stebz(n, L.ptr(d), L.ptr(e), L.ptr(w), L.ptr(scratch), L.stream_ptr())
    L.check(rc, "xmca_stebz")
    return w


def stein(d, e, lam_host, cluster_start, tnorm, iterations=2):
    """Eigenvectors (rows of the returned k x n tensor) of the tridiagonal for the eigenvalues
    `lam_host` (descending numpy array); `cluster_start`: numpy int array of cluster boundaries."""
    lib = L.load()
    t = torch()
    n, k = d.shape[0], int(lam_host.size)
    ncl = int(len(cluster_start) - 1)
    lam = to_device(np.ascontiguousarray(lam_host, dtype=np.float64))
    cs = to_device(np.ascontiguousarray(cluster_start, dtype=np.int32))
    Z = empty((k, n), t.float64)
    ws_bytes = lib.xmca_stein_workspace_bytes(n, ncl)
    ws = empty((ws_bytes,), t.uint8)
    rc = lib.xmca_stein(n, L.ptr(d), L.ptr(e), k, L.ptr(lam), L.ptr(cs), ncl, float(tnorm), int(iterations),
                        L.ptr(Z), n, L.ptr(ws), ws_bytes, L.stream_ptr())
    L.check(rc, "xmca_stein")
    return Z


def ormtr(S_reflectors, tau, Z):
    """Rows of Z <- Q row (in place), Q from `sytrd`."""
    lib = L.load()
    n = S_reflectors.shape[0]
    rc = lib.xmca_ormtr(n, L.ptr(S_reflectors), _ld(S_reflectors), L.ptr(tau), Z.shape[0], L.ptr(Z), _ld(Z),
                        L.stream_ptr())
    L.check(rc, "xmca_ormtr")
    return Z


# ----------------------------------------------------------- analytic signal
def hilbert_matrix(T, dtype):
    """Circulant H (T x T) with imag(scipy.signal.hilbert(x, axis=0)) = H @ x (array.py:464)."""
    lib = L.load()
    H = empty((T, T), dtype)
    taps = empty((T,), f64())
    rc = lib.xmca_hilbert_matrix(T, L.ptr(H), L.dtype_code(H), T, L.ptr(taps), L.stream_ptr())
    L.check(rc, "xmca_hilbert_matrix")
    return H


def hilbert_block(N, rows, cols, row0, col0, dtype):
    """rows x cols block at (row0, col0) of the length-N circulant Hilbert operator."""
    lib = L.load()
    H = empty((rows, cols), dtype)
    taps = empty((N,), f64())
    rc = lib.xmca_hilbert_block(N, rows, cols, row0, col0, L.ptr(H), L.dtype_code(H), cols, L.ptr(taps), L.stream_ptr())
    L.check(rc, "xmca_hilbert_block")
    return H


def hilbert_matrix_exp_extension(T, theta, dtype):
    """T x T operator M with  imag(analytic signal of the EXTENDED series)[T:2T] = M x  for the exponential
    fore/back-cast of array.py:378-411 / :455-472 (extend='exp', period = theta): the series is extended to
    [pre, x, post] (3T), transformed, and the middle third kept.  pre/post are linear in x -- three functionals
    (slope with the reference's `xstd = mean(x)`, end of the regression line, offset) times three basis
    vectors -- so  M = H_mid,mid + (H_mid,pre B_pre) F_pre + (H_mid,post B_post) F_post  (two rank-3 updates)."""
    N = T
    x = np.arange(N, dtype=np.float64)
    xmean = (N - 1) / 2.0
    s = (x - xmean) / (N * xmean ** 2)                    # slope = s . f   (array.py:384: xstd = np.mean(x))
    lin = np.full(N, 1.0 / N) + s * xmean                 # linear_end = ymean + slope * xmean
    off = -lin.copy()
    off[-1] += 1.0                                        # offset = f[-1] - linear_end
    F_post = np.stack([s, lin, off])                      # 3 x T
    B_post = np.stack([x, np.ones(N), np.exp(-(x + 1.0) / theta)], axis=1)      # T x 3
    F_pre, B_pre = F_post[:, ::-1].copy(), B_post[::-1].copy()                  # pre = extend(f[::-1])[::-1]
    N3 = 3 * T
    M = hilbert_block(N3, T, T, T, T, f64())
    for col0, Bm, Fm in ((0, B_pre, F_pre), (2 * T, B_post, F_post)):
        Hb = hilbert_block(N3, T, T, T, col0, f64())
        U = matmul(Hb, to_device(Bm))                                            # T x 3
        matmul(U, to_device(Fm), out=M, accumulate=True)
        del Hb
    return M if dtype == f64() else scale_copy(M, out_dtype=dtype)


def dft_matrix(T, dtype):
    """Stacked [Re; Im] one-sided, weighted, orthonormally scaled DFT operator (2 floor(T/2) x T)."""
    lib = L.load()
    F = empty((int(lib.xmca_dft_rows(T)), T), dtype)
    rc = lib.xmca_dft_matrix(T, L.ptr(F), L.dtype_code(F), T, L.stream_ptr())
    L.check(rc, "xmca_dft_matrix")
    return F


def embed_complex(Z):
    """[[Zr, -Zi], [Zi, Zr]] from the stacked Z = [Zr; Zi] (2 Tp x S)."""
    lib = L.load()
    rows2, S = Z.shape
    Tp = rows2 // 2
    E = empty((rows2, 2 * S), Z.dtype)
    rc = lib.xmca_embed_complex(L.ptr(Z), L.dtype_code(Z), _ld(Z), Tp, S, L.ptr(E), L.dtype_code(E), 2 * S,
                                L.stream_ptr())
    L.check(rc, "xmca_embed_complex")
    return E


def apply_time_operator(Op, X, x_planes=None):
    """Op (M x T) @ X (T x S) in the dtype of X: 3xTF32 tcgen05 GEMM for fp32 fields (Op and X^T are
    split into TF32 planes; pass `x_planes` to reuse the split of X), fp64 CUDA cores otherwise.
    Returns (result, x_planes)."""
    t = torch()
    if X.dtype == t.float32:
        if x_planes is None:
            x_planes = split_tf32(X, transpose=True)          # S x T, K-major
        ohi, olo, K = split_tf32(Op)
        out = tc_gemm_nt(ohi, olo, x_planes[0], x_planes[1], K)
        return out, x_planes
    return matmul(Op, X, out_dtype=X.dtype), None


# ------------------------------------------------------------- element-wise
def scale_copy(X, out_dtype=None, col_scale=None, row_scale=None, out=None):
    lib = L.load()
    rows, cols = X.shape
    if out is None:
        out = empty((rows, cols), out_dtype or X.dtype)
    rc = lib.xmca_scale_copy(L.ptr(X), L.dtype_code(X), _ld(X), L.ptr(out), L.dtype_code(out), _ld(out),
                             rows, cols, L.ptr(col_scale), L.ptr(row_scale), L.stream_ptr())
    L.check(rc, "xmca_scale_copy")
    return out


def transpose(X, out_dtype=None):
    lib = L.load()
    rows, cols = X.shape
    out = empty((cols, rows), out_dtype or X.dtype)
    rc = lib.xmca_transpose(L.ptr(X), L.dtype_code(X), rows, cols, _ld(X), L.ptr(out), L.dtype_code(out),
                            _ld(out), L.stream_ptr())
    L.check(rc, "xmca_transpose")
    return out


def col_sumsq(X, row0=0, row1=None):
    lib = L.load()
    rows, cols = X.shape
    out = empty((cols,), f64())
    rc = lib.xmca_col_sumsq(L.ptr(X), L.dtype_code(X), _ld(X), row0, rows if row1 is None else row1, cols,
                            L.ptr(out), L.stream_ptr())
    L.check(rc, "xmca_col_sumsq")
    return out


def center_columns(X):
    lib = L.load()
    rows, cols = X.shape
    mean = empty((cols,), f64())
    rc = lib.xmca_center_columns(L.ptr(X), L.dtype_code(X), rows, cols, _ld(X), L.ptr(mean), L.stream_ptr())
    L.check(rc, "xmca_center_columns")
    return mean


def field_stats(X):
    """Column mean / std / NaN flag and row validity of a raw field (host numpy results)."""
    lib = L.load()
    t = torch()
    rows, cols = X.shape
    mean, std = empty((cols,), f64()), empty((cols,), f64())
    col_nan, row_ok = empty((cols,), t.int32), empty((rows,), t.int32)
    rc = lib.xmca_field_stats(L.ptr(X), L.dtype_code(X), rows, cols, _ld(X), L.ptr(mean), L.ptr(std),
                              L.ptr(col_nan), L.ptr(row_ok), L.stream_ptr())
    L.check(rc, "xmca_field_stats")
    return to_host(mean), to_host(std), to_host(col_nan) != 0, to_host(row_ok) != 0


def compact_center(X, keep_idx, mean):
    """Device field without its NaN columns, centred: Y[:, j] = X[:, idx[j]] - mean[idx[j]]."""
    lib = L.load()
    rows = X.shape[0]
    n_keep = int(keep_idx.size)
    Y = empty((rows, n_keep), X.dtype)
    if n_keep == 0:
        return Y
    idx = to_device(np.ascontiguousarray(keep_idx, dtype=np.int64))
    mu = to_device(np.ascontiguousarray(mean, dtype=np.float64))
    rc = lib.xmca_compact_center(L.ptr(X), L.dtype_code(X), rows, _ld(X), L.ptr(idx), n_keep, L.ptr(mu),
                                 L.ptr(Y), L.dtype_code(Y), _ld(Y), L.stream_ptr())
    L.check(rc, "xmca_compact_center")
    return Y


def fill_normal(X, seed, stream_id):
    lib = L.load()
    rows, cols = X.shape
    rc = lib.xmca_fill_normal(L.ptr(X), L.dtype_code(X), rows, cols, _ld(X), int(seed) & (2 ** 64 - 1),
                              int(stream_id) & (2 ** 64 - 1), L.stream_ptr())
    L.check(rc, "xmca_fill_normal")
    return X


def gather_rows(X, idx, row_scale=None, out_dtype=None, cols=None):
    """Y[i] = X[idx[i], :cols] * row_scale[i]; idx is a device int64 tensor."""
    lib = L.load()
    n_out = idx.shape[0]
    cols = X.shape[1] if cols is None else cols
    out = empty((n_out, cols), out_dtype or X.dtype)
    rc = lib.xmca_gather_rows(L.ptr(X), L.dtype_code(X), _ld(X), L.ptr(idx), n_out, cols, L.ptr(row_scale),
                              L.ptr(out), L.dtype_code(out), _ld(out), L.stream_ptr())
    L.check(rc, "xmca_gather_rows")
    return out


def row_sumsq(X):
    lib = L.load()
    rows, cols = X.shape
    out = empty((rows,), f64())
    rc = lib.xmca_row_sumsq(L.ptr(X), L.dtype_code(X), _ld(X), rows, cols, L.ptr(out), L.stream_ptr())
    L.check(rc, "xmca_row_sumsq")
    return out


def col_absmax(X, row_scale=None):
    lib = L.load()
    rows, cols = X.shape
    out = empty((cols,), f64())
    rc = lib.xmca_col_absmax(L.ptr(X), L.dtype_code(X), _ld(X), rows, cols, L.ptr(row_scale), L.ptr(out),
                             L.stream_ptr())
    L.check(rc, "xmca_col_absmax")
    return out


def promax_target(B, row_scale, colmax, power):
    """X = B * row_scale[:, None]; P = Xn |Xn|^(power-1), Xn = X / colmax (rotation.py:115-124)."""
    lib = L.load()
    rows, cols = B.shape
    X = empty((rows, cols), f64())
    P = empty((rows, cols), f64())
    rc = lib.xmca_promax_target(L.ptr(B), _ld(B), rows, cols, L.ptr(row_scale), L.ptr(colmax), float(power),
                                L.ptr(X), L.ptr(P), cols, L.stream_ptr())
    L.check(rc, "xmca_promax_target")
    return X, P


# ------------------------------------------------------------------ Varimax
def varimax(Ld, gamma=1.0, max_iter=1000, tol=1e-8):
    """Device Varimax (rotation.py:15-78).  Ld: n x p loadings (fp32/fp64).
    Returns (B fp64 n x p, R fp64 p x p, iterations).  Raises NotConvergedError."""
    lib = L.load()
    t = torch()
    n, p = Ld.shape
    B = empty((n, p), t.float64)
    R = empty((p, p), t.float64)
    out = zeros((16,), t.float64)
    ws_bytes = lib.xmca_varimax_workspace_bytes(n, p)
    ws = empty((ws_bytes,), t.uint8)
    iters = C.c_int(0)
    rc = lib.xmca_varimax(L.ptr(Ld), L.dtype_code(Ld), n, p, _ld(Ld), float(gamma), int(max_iter), float(tol),
                          L.ptr(B), p, L.ptr(R), C.byref(iters), L.ptr(out), L.ptr(ws), ws_bytes,
                          L.stream_ptr())
    global last_varimax_stats
    last_varimax_stats = out
    L.check(rc, "xmca_varimax")
    return B, R, iters.value


def varimax_complex(Lr, Li, gamma=1.0, max_iter=1000, tol=1e-8):
    """Device Varimax for complex loadings given as planar (re, im) n x p tensors (p <= 32).
    Returns (Br, Bi fp64 n x p, R complex128 host p x p, iterations).  Raises NotConvergedError."""
    lib = L.load()
    t = torch()
    n, p = Lr.shape
    assert Li.shape == Lr.shape and Li.dtype == Lr.dtype and _ld(Li) == _ld(Lr)
    Br, Bi = empty((n, p), t.float64), empty((n, p), t.float64)
    Rr, Ri = empty((p, p), t.float64), empty((p, p), t.float64)
    out = zeros((16,), t.float64)
    ws_bytes = lib.xmca_varimax_complex_workspace_bytes(n, p)
    ws = empty((ws_bytes,), t.uint8)
    iters = C.c_int(0)
    rc = lib.xmca_varimax_complex(L.ptr(Lr), L.ptr(Li), L.dtype_code(Lr), n, p, _ld(Lr), float(gamma), int(max_iter),
                                  float(tol), L.ptr(Br), L.ptr(Bi), p, L.ptr(Rr), L.ptr(Ri), C.byref(iters),
                                  L.ptr(out), L.ptr(ws), ws_bytes, L.stream_ptr())
    global last_varimax_stats
    last_varimax_stats = out
    L.check(rc, "xmca_varimax_complex")
    return Br, Bi, to_host(Rr) + 1j * to_host(Ri), iters.value


def col_absmax_complex(Xr, Xi, row_scale=None):
    lib = L.load()
    rows, cols = Xr.shape
    out = empty((cols,), f64())
    rc = lib.xmca_col_absmax_complex(L.ptr(Xr), L.ptr(Xi), _ld(Xr), rows, cols, L.ptr(row_scale), L.ptr(out),
                                     L.stream_ptr())
    L.check(rc, "xmca_col_absmax_complex")
    return out


def promax_target_complex(Br, Bi, row_scale, colmax, power):
    """Planar complex X = B * row_scale[:, None] and P = Xn |Xn|^(power-1), Xn = X / colmax."""
    lib = L.load()
    rows, cols = Br.shape
    out = [empty((rows, cols), f64()) for _ in range(4)]
    rc = lib.xmca_promax_target_complex(L.ptr(Br), L.ptr(Bi), _ld(Br), rows, cols, L.ptr(row_scale), L.ptr(colmax),
                                        float(power), L.ptr(out[0]), L.ptr(out[1]), L.ptr(out[2]), L.ptr(out[3]),
                                        cols, L.stream_ptr())
    L.check(rc, "xmca_promax_target_complex")
    return out
