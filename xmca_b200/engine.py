"""Device pipelines behind ``MCA.solve`` / ``rotate`` / getters / ``rule_n``.

Everything here runs on the GPU through the C ABI (``device.py``); the host
only sorts the O(rank) singular values and does O(p^3) algebra for p <= 64
(Promax fit), exactly the pieces SURVEY.md section 2b leaves on the host.

Math (real fields A: T x S1, B: T x S2, centred; dof = T-1; the reference gets
the same quantities from three LAPACK SVDs, array.py:552-584):

  direct route (min(S1,S2) <= T):  C = A^T B / dof  (S1 x S2) is formed and its
      SVD C = V_L Sigma V_R^T computed by the blocked one-sided Jacobi kernel.
  Gram route  (T < min(S1,S2)):   G_A = A A^T = U_A L_A U_A^T, G_B likewise
      (T x T, fp64).  With F = U sqrt(L):  K = F_A^T F_B / dof = P Sigma Q^T has the
      singular values of C, and  V_L = A^T (F_B Q) / (Sigma dof),
      V_R = B^T (F_A P) / (Sigma dof)  -- divisions only by Sigma itself.
  PCA: C = A^T A / dof (direct) or  sigma = L_A / dof, V = A^T U_A / sqrt(L_A).

Complex fields use the real embedding E(Z) = [[X, -Y], [Y, X]], a ring
homomorphism (E(Z1 Z2) = E(Z1) E(Z2), E(Z^H) = E(Z)^T); every singular value of
E(C) is doubled and each singular pair yields one complex vector.
"""
from __future__ import annotations

import numpy as np

from . import device as D
from ._lib import NotConvergedError as L_NotConverged


# The tridiagonal route is used when the symmetric eigenproblem has at least this many rows
# (below it the Jacobi routes finish in milliseconds and deliver every vector at once) ...
TRIDIAG_MIN_N = 768
# ... and it computes singular vectors on demand for requests of up to this many leading modes;
# larger requests (e.g. the full `_V` of the reference) fall back to the Jacobi routes.
TRIDIAG_MAX_VECTORS = 512
# Tridiagonalisation: "two_stage" (xmca_sytrd2: dense -> band on the DMMA pipe -> bulge chasing; the default) or
# "one_stage" (xmca_sytrd: BLAS-2, HBM bound; also the fallback when a panel factorisation of the two-stage
# reduction reports a breakdown).  XMCA_SYTRD=one_stage in the environment selects the old path (A/B runs).
import os as _os
SYTRD_MODE = _os.environ.get("XMCA_SYTRD", "two_stage")


class SolveResult:
    """Device-resident result of one solve: sigma (host fp64, descending, length
    rank) and V[k] (S'_k x rank, field dtype, device)."""

    def __init__(self, sigma, V, route, sweeps, frob2=None):
        self.sigma = sigma
        self.V = V
        self.route = route
        self.sweeps = sweeps
        self.frob2 = frob2

    def vectors(self, m):
        """First m singular vectors per field (device, S'_k x m)."""
        return {k: v[:, :m] for k, v in self.V.items()}


def _jacobi_tol(field):
    """Largest admissible cosine between two rotated columns when the Jacobi sweeps stop:
    1e-6 for fp32 fields (singular values then carry ~1e-12 / relative gap, far below the
    fp32 noise of the data), the library default (1e-11) for fp64 fields."""
    return 1e-6 if field.dtype == D.f32() else 0.0


def _order(sigma_dev, keep):
    s = D.to_host(sigma_dev)
    order = np.argsort(-s, kind="stable")[:keep]
    return s[order], order


def _idx(order):
    return D.to_device(np.ascontiguousarray(order, dtype=np.int64))


def _inv_or_zero(x, floor):
    out = np.zeros_like(x)
    ok = x > floor
    out[ok] = 1.0 / x[ok]
    return out


def _rows_to_cols(rows, out_dtype):
    """(r x S) row vectors -> (S x r) column vectors in the requested dtype."""
    return D.transpose(rows, out_dtype=out_dtype)


def solve_real(A, B, want_vectors=True, force_route=None, use_tensor_cores=None, null_basis=None, dof=None):
    """A, B: centred device fields (T x S1, T x S2), fp32 or fp64; B may be None (PCA).
    null_basis: T x k device matrix of exact null vectors of X X^T (default: the constant
    vector that centring creates; 0: none).  dof: degrees of freedom (default T - 1)."""
    t = D.torch()
    T, S1 = A.shape
    dof = float(T - 1) if dof is None else float(dof)
    out_dtype = A.dtype
    pca = B is None
    S2 = S1 if pca else B.shape[1]
    rank = min(T, S1, S2)
    route = force_route or ("direct" if min(S1, S2) <= T else "gram")
    if force_route is None and TRIDIAG_MIN_N <= rank <= D.sytrd_max_n():
        route = "tridiag"
    if route == "tridiag":
        try:
            return TridiagResult(A, B, null_basis, dof)
        except np.linalg.LinAlgError:
            route = "gram_eig" if min(S1, S2) > T else "direct"
    sweeps = []
    if route in ("gram", "cholqr"):
        try:
            return _solve_cholqr(A, B, want_vectors, null_basis, dof)
        except np.linalg.LinAlgError:
            if route == "cholqr":
                raise
            route = "gram_eig"      # Gram matrix not numerically SPD: eigen route copes with rank deficiency

    if route == "direct":
        # ---- C (or C^T) with the short side as rows, fp64 accumulation -------------
        left_short = S1 <= S2
        frob2 = None
        tc = use_tensor_cores if use_tensor_cores is not None else False
        Bx = A if pca else B
        if tc and A.dtype == t.float32:
            X32, frob2 = D.cov_gemm_tc(A if left_short else Bx, Bx if left_short else A, 1.0 / dof)
            X = X32
        else:
            X = D.matmul(A if left_short else Bx, Bx if left_short else A, trans_a=True, alpha=1.0 / dof)
        Xr, sig, Jt, sw = D.jacobi_svd(X, want_v=want_vectors, tol=_jacobi_tol(A))
        sweeps.append(sw)
        sigma, order = _order(sig, rank)
        if not want_vectors:
            return SolveResult(sigma, {}, route, sweeps, frob2)
        idx = _idx(order)
        n_short = X.shape[0]
        inv = D.to_device(_inv_or_zero(sigma, 0.0))
        short_rows = D.gather_rows(Jt, idx, cols=n_short)                 # rank x S_short
        long_rows = D.gather_rows(Xr, idx, row_scale=inv)                 # rank x S_long (unit rows)
        Vs, Vl = _rows_to_cols(short_rows, out_dtype), _rows_to_cols(long_rows, out_dtype)
        if pca:
            return SolveResult(sigma, {"left": Vs}, route, sweeps, frob2)
        V = {"left": Vs, "right": Vl} if left_short else {"left": Vl, "right": Vs}
        return SolveResult(sigma, V, route, sweeps, frob2)

    # -------------------------------- Gram route ---------------------------------
    # G = X X^T is symmetric PSD, so one-sided Jacobi on it needs NO accumulated
    # rotations: at convergence the rotated rows are lambda_j u_j^T, i.e. the
    # eigenvectors are the normalised rows themselves (saves a third of the sweep cost).
    # Directions with lambda_j below the rounding floor (the null vector that centring
    # creates) carry no information and are zeroed instead of normalised.
    eps = 2.220446049250313e-16

    def gram_factor(X):
        G = D.matmul(X, X, trans_b=True)                                  # T x T fp64
        Gr, lam, _, sw = D.jacobi_svd(G, want_v=False, tol=_jacobi_tol(A))
        sweeps.append(sw)
        return D.to_host(lam), Gr

    lamA, GrA = gram_factor(A)
    npad = GrA.shape[0]
    if pca:
        order = np.argsort(-lamA, kind="stable")[:rank]
        lam = lamA[order]
        sigma = lam / dof
        if not want_vectors:
            return SolveResult(sigma, {}, route, sweeps)
        floor = lam[0] * T * eps * 64 if lam.size else 0.0
        inv = _inv_or_zero(lam, floor)
        Ut = D.gather_rows(GrA, _idx(order), row_scale=D.to_device(inv * np.sqrt(inv)), cols=T)  # u_j / s_j
        V = D.matmul(A, Ut, trans_a=True, trans_b=True, out_dtype=out_dtype)   # S x rank
        return SolveResult(sigma, {"left": V}, route, sweeps)

    lamB, GrB = gram_factor(B)
    all_idx = _idx(np.arange(npad))

    def inv_sqrt(lam):
        floor = lam.max() * T * eps * 64 if lam.size else 0.0
        return D.to_device(np.sqrt(_inv_or_zero(np.maximum(lam, 0.0), floor)))

    FAt = D.gather_rows(GrA, all_idx, row_scale=inv_sqrt(lamA), cols=T)   # rows sqrt(lambda_j) u_j^T (= F_A^T)
    FBt = D.gather_rows(GrB, all_idx, row_scale=inv_sqrt(lamB), cols=T)
    del GrA, GrB
    K = D.matmul(FAt, FBt, trans_b=True, alpha=1.0 / dof)                 # npad x npad, = F_A^T F_B / dof
    Kr, sig, Jt, sw = D.jacobi_svd(K, want_v=want_vectors, tol=_jacobi_tol(A))
    sweeps.append(sw)
    sigma, order = _order(sig, rank)
    if not want_vectors:
        return SolveResult(sigma, {}, route, sweeps)
    # K = J Sigma U^T (rows of K orthogonalised): P = J columns, Q = unit rows of Kr
    floor = sigma[0] * T * eps if sigma.size else 0.0
    inv = _inv_or_zero(sigma, floor)
    idx = _idx(order)
    Pt = D.gather_rows(Jt, idx, row_scale=D.to_device(inv / dof), cols=npad)          # rank x npad : p_j / (sigma dof)
    Qt = D.gather_rows(Kr, idx, row_scale=D.to_device(inv * inv / dof), cols=npad)    # unit q_j / (sigma dof)
    WLt = D.matmul(Qt, FBt)                                               # rank x T : (F_B q_j)^T / (sigma dof)
    WRt = D.matmul(Pt, FAt)
    VL = D.matmul(A, WLt, trans_a=True, trans_b=True, out_dtype=out_dtype)            # S1 x rank
    VR = D.matmul(B, WRt, trans_a=True, trans_b=True, out_dtype=out_dtype)
    return SolveResult(sigma, {"left": VL, "right": VR}, route, sweeps)


# ------------------------------------------------------- tridiagonal route
class TridiagResult:
    """Full spectrum by Householder tridiagonalisation + bisection, vectors on demand.

    sigma(C)^2 (MCA) / sigma(C) (PCA) are the eigenvalues of ONE symmetric matrix S:
      Gram side  (T < min(S1, S2)), with G_X = X X^T (T x T, fp64):
         MCA:  S = L_B^T G_A L_B / dof^2,   G_B + mu N N^T = L_B L_B^T  (N: the exact null
               vectors that centring gives G_B, lifted by mu = mean eigenvalue so that the
               Cholesky factor exists; A^T N = 0, so the lift never reaches C)
               V_R = B^T (L_B^-T q),   V_L = A^T (L_B q) / (sigma dof)
         PCA:  S = G_A / dof,   V = A^T q / sqrt(lambda dof)
      direct side (S_short <= T):  MCA: S = C_s C_s^T with C_s = X_s^T X_l / dof,
               V_short = q, V_long = C_s^T q / sigma;   PCA: S = A^T A / dof, V = q.
    S = Q T Q^T (xmca_sytrd), all eigenvalues by bisection (xmca_stebz); the first m
    eigenvectors q by inverse iteration on T (xmca_stein) and x = Q z (xmca_ormtr) when
    `vectors(m)` is called.  The reference computes all `rank` vectors in solve()
    (array.py:584); here requests beyond TRIDIAG_MAX_VECTORS run the Jacobi route once."""

    route = "tridiag"

    def __init__(self, A, B, null_basis=None, dof=None, defer=False):
        T, S1 = A.shape
        self.A, self.B = A, B
        self.pca = B is None
        S2 = S1 if self.pca else B.shape[1]
        self.dof = float(T - 1) if dof is None else float(dof)
        self.out_dtype = A.dtype
        self.null_basis = null_basis
        self.gram_side = T < min(S1, S2)
        self.sweeps = []
        self.frob2 = None
        self._cache = (0, None)
        self._full = None
        self.n_null = 0 if isinstance(null_basis, int) or not self.gram_side else \
            (null_basis.shape[1] if null_basis is not None else 1)
        S = self._build_S()
        self.n = S.shape[0]
        self._S = S
        if not defer:                  # (deferred: `solve_real_pair` reduces the two models of a pair itself)
            self.reduce()

    def _build_S(self):
        A, B, null_basis, dof = self.A, self.B, self.null_basis, self.dof
        T, S1 = A.shape
        S2 = S1 if self.pca else B.shape[1]

        def gram(X, alpha=1.0):
            """X X^T (T x T, fp64): tcgen05 3xTF32 with fp64 chunk accumulation for fp32 fields,
            fp64 DMMA product otherwise."""
            if X.dtype == D.f32() and X.shape[1] >= 64:
                return D.gram_tc(X, alpha)
            return D.matmul(X, X, trans_b=True, alpha=alpha, symmetric=True)

        if self.gram_side:
            if isinstance(null_basis, int):          # 0: the Gram matrices have no structural null vector
                Nb = None
            else:
                Nb = null_basis if null_basis is not None else D.to_device(np.full((T, 1), 1.0 / np.sqrt(T)))
            if self.pca:
                S = gram(A, 1.0 / dof)
            else:
                GA = gram(A)
                GB = gram(B)
                tr = float(D.to_host(D.col_sumsq(B)).sum())
                if not np.isfinite(tr) or tr <= 0.0:
                    raise np.linalg.LinAlgError("empty or non-finite field")
                if Nb is not None:
                    D.matmul(Nb, Nb, trans_b=True, alpha=tr / T, out=GB, accumulate=True)
                self.LB, self.invB = D.cholesky(GB, min_pivot=1e-11 * tr / T)
                W = D.matmul(GA, self.LB, b_lower=True)
                del GA
                S = D.matmul(self.LB, W, trans_a=True, alpha=1.0 / dof ** 2, symmetric=True, a_lower_t=True)
                del W
        else:
            self.left_short = S1 <= S2
            if self.pca:
                S = D.matmul(A, A, trans_a=True, alpha=1.0 / dof, symmetric=True)
            else:
                Xs, Xl = (A, B) if self.left_short else (B, A)
                self.C = D.matmul(Xs, Xl, trans_a=True, alpha=1.0 / dof)          # S_short x S_long
                S = D.matmul(self.C, self.C, trans_b=True, symmetric=True)
        return S

    def reduce(self):
        """S = Q T Q^T.  Two-stage reduction (xmca_sytrd2); if one of its panel factorisations reports a breakdown
        (exactly rank-deficient panel, non-finite data) S is formed again and reduced by the one-stage xmca_sytrd."""
        S = self._S
        if SYTRD_MODE == "two_stage":
            try:
                d, e, tfac = D.sytrd2(S)
                return self._finish(d, e, None, S, tfac)
            except np.linalg.LinAlgError:
                S = self._S = self._build_S()
        self._finish(*D.sytrd(S), S)

    def reduce_begin(self):
        """Enqueue the two-stage reduction on the CURRENT stream without synchronising (`reduce_end` finishes)."""
        S = self._S
        d, e, tfac, pending = D.sytrd2(S, sync=False)
        self._pending = (d, e, tfac, pending, S)

    def reduce_end(self):
        d, e, tfac, pending, S = self._pending
        self._pending = None
        try:
            D.sytrd2_check(pending)
            return self._finish(d, e, None, S, tfac)
        except np.linalg.LinAlgError:
            S = self._S = self._build_S()
        self._finish(*D.sytrd(S), S)

    def _finish(self, d, e, tau, Q, tfac=None):
        """Spectrum from the tridiagonal form (d, e) of S; Q: the matrix that holds the reflectors (one-stage: rows +
        tau; two-stage: panel reflectors below the band, sweep reflectors above the diagonal, + tfac)."""
        self._S = None
        self.d, self.e, self.tau, self.tfac = d, e, tau, tfac
        self.Q = Q
        lam = D.to_host(D.stebz(self.d, self.e))
        if not np.isfinite(lam).all():
            raise np.linalg.LinAlgError("non-finite spectrum")
        lam = np.maximum(lam, 0.0)
        if self.n_null:
            lam[self.n - self.n_null:] = 0.0           # the centring null directions: structurally zero
        self.lam = lam
        self.sigma = lam.copy() if self.pca else np.sqrt(lam)

    # -- eigenvectors of S for the leading m eigenvalues, as ROWS (m x n, fp64)
    def _eigvec_rows(self, m):
        lam = self.lam[:m]
        tnorm = float(self.lam[0]) if self.lam.size else 0.0
        # eigenvalues closer than `gap` share a cluster (mutual re-orthogonalisation, run
        # sequentially); vectors of different clusters are computed independently and are then
        # orthogonal to ~eps * tnorm / gap: 1e-7 for fp32 fields, 1e-10 for fp64 fields
        gap = 2.2e-16 * tnorm / (1e-7 if self.out_dtype == D.f32() else 1e-10)
        starts = [0] + [i for i in range(1, m) if lam[i - 1] - lam[i] > gap] + [m]
        Z = D.stein(self.d, self.e, lam, np.asarray(starts), tnorm, iterations=2)
        if self.tfac is not None:
            return D.ormtr2(self.Q, self.tfac, Z)
        return D.ormtr(self.Q, self.tau, Z)

    def vectors(self, m):
        m = int(min(max(m, 0), self.sigma.size))
        have, V = self._cache
        if V is not None and have >= m:
            return {k: v[:, :m] for k, v in V.items()}
        if m > TRIDIAG_MAX_VECTORS:
            return {k: v[:, :m] for k, v in self.V.items()}
        want = int(min(max(m, 16), self.sigma.size, TRIDIAG_MAX_VECTORS))
        Z = self._eigvec_rows(want)
        A, B, dof, od = self.A, self.B, self.dof, self.out_dtype
        sig = self.sigma[:want]
        if self.gram_side:
            if self.pca:
                floor = self.lam[0] * self.n * 2.3e-16 * 64 if self.lam.size else 0.0
                inv = np.sqrt(_inv_or_zero(self.lam[:want] * dof, floor * dof))
                Zs = D.scale_copy(Z, row_scale=D.to_device(inv))
                V = {"left": D.matmul(A, Zs, trans_a=True, trans_b=True, out_dtype=od)}
            else:
                # left: V_L = A^T (L_B q) / (sigma dof); right: V_R = C^T V_L / sigma = B^T (A V_L) / (sigma dof) -- one
                # more skinny product instead of the triangular solve L_B^-T q (128 dependent block steps, 5.5 ms at
                # T = 8192); its error ~ u sigma_1 / sigma_k is that of the reference's direct SVD.  Null modes
                # (sigma below the floor) get zero vectors on both sides.
                inv = _inv_or_zero(sig, 1e-7 * sig[0] if sig.size else 0.0)
                scale = D.to_device(inv / dof)
                Zs = D.scale_copy(Z, row_scale=scale)
                Yt = D.matmul(Zs, self.LB, trans_b=True)                      # rows (L_B q)^T / (sigma dof)
                VL64 = D.matmul(A, Yt, trans_a=True, trans_b=True)            # S1 x m, fp64
                P = D.scale_copy(D.matmul(A, VL64), col_scale=scale)          # T x m: A V_L / (sigma dof)
                VR = D.matmul(B, P, trans_a=True, out_dtype=od)
                VL = VL64 if od == VL64.dtype else VL64.to(od)
                V = {"left": VL, "right": VR}
        else:
            Vs = D.transpose(Z, out_dtype=od)                                 # S_short x m
            if self.pca:
                V = {"left": Vs}
            else:
                inv = _inv_or_zero(sig, 1e-7 * sig[0] if sig.size else 0.0)
                Zs = D.scale_copy(Z, row_scale=D.to_device(inv))
                Vl = D.matmul(self.C, Zs, trans_a=True, trans_b=True, out_dtype=od)
                V = {"left": Vs, "right": Vl} if self.left_short else {"left": Vl, "right": Vs}
        self._cache = (want, V)
        return {k: v[:, :m] for k, v in V.items()}

    @property
    def V(self):
        """All `rank` singular vectors (the reference's `_V`): one Jacobi solve, cached."""
        if self.sigma.size <= TRIDIAG_MAX_VECTORS:
            return self.vectors(self.sigma.size)
        if self._full is None:
            T, S1 = self.A.shape
            S2 = S1 if self.pca else self.B.shape[1]
            route = "gram" if T < min(S1, S2) else "direct"
            res = solve_real(self.A, self.B, want_vectors=True, force_route=route, null_basis=self.null_basis,
                             dof=self.dof)
            self.sweeps = res.sweeps
            self._full = res.V
        return self._full


# ---------------------------------------------------------------- Cholesky-QR
def _solve_cholqr(A, B, want_vectors, null_basis=None, dof=None):
    """T < min(S1, S2): ONE T x T Jacobi SVD instead of three decompositions.

    Centring makes n = 1/sqrt(T) an exact null vector of X X^T (for the real
    embedding of a complex field there are two, ``null_basis``).  With
    X~ = [X, sqrt(mu) n] (one column appended per null vector; it lifts the zero
    eigenvalue to mu = 4 trace(X X^T), above every genuine one) and the Cholesky
    factor X~ X~^T = L L^T:   X~^T = Q L^T with Q = X~^T L^-T orthonormal, hence
        C~ = A~^T B~ / dof = Q_A (L_A^T L_B / dof) Q_B^T = (Q_A P) Sigma (Q_B Q)^T.
    C~ = diag(C, sqrt(mu_A mu_B) / dof): the appended directions do not couple with
    the data (X^T n = 0) and show up as extra singular values, the largest by
    construction, which are dropped.  V_X = X^T (L_X^-T P).
    PCA: C = A^T A / dof = Q_A (L_A^T L_A / dof) Q_A^T -> SVD of L_A itself."""
    T, S1 = A.shape
    dof = float(T - 1) if dof is None else float(dof)
    pca = B is None
    out_dtype = A.dtype
    if isinstance(null_basis, int):              # 0: no structural null vector, nothing to lift
        Nb, k = None, 0
    else:
        Nb = null_basis if null_basis is not None else D.to_device(np.full((T, 1), 1.0 / np.sqrt(T)))
        k = Nb.shape[1]

    def factor(X):
        G = D.matmul(X, X, trans_b=True)                                  # T x T fp64 Gram
        tr = float(D.to_host(D.col_sumsq(X)).sum())                        # trace(G) = ||X||_F^2
        if not np.isfinite(tr) or tr <= 0.0:
            raise np.linalg.LinAlgError("empty or non-finite field")
        mu = 4.0 * tr
        if Nb is not None:
            D.matmul(Nb, Nb, trans_b=True, alpha=mu, out=G, accumulate=True)   # + mu N N^T
        # pivots below ~1e-11 of the mean diagonal: rank deficient beyond the centring null
        # vector (e.g. repeated time steps) -> LinAlgError -> eigen route
        Lm, inv = D.cholesky(G, min_pivot=1e-11 * tr / T)
        return Lm, inv, mu

    LA, invA, muA = factor(A)
    if pca:
        M = LA
        extra = muA / dof                        # eigenvalue of the appended directions, / dof
    else:
        LB, invB, muB = factor(B)
        M = D.matmul(LA, LB, trans_a=True, alpha=1.0 / dof)               # T x T
        extra = np.sqrt(muA * muB) / dof
    Mr, sig, _, sw = D.jacobi_svd(M, want_v=False, tol=_jacobi_tol(A))      # rows of Mr: sigma_j q_j^T (right singular vectors of M)
    s = D.to_host(sig)
    order = np.argsort(-s, kind="stable")[:T]
    sv = s[order]
    sigma_all = sv * sv / dof if pca else sv
    if k and (np.abs(sigma_all[:k] - extra).max() > 1e-6 * extra or (T > k and sigma_all[k] > 0.5 * extra)):
        raise np.linalg.LinAlgError("appended directions not isolated (fields not centred?)")
    sigma = np.concatenate([sigma_all[k:], np.zeros(k)])  # rank = T modes, the last k are the centring null modes
    if not want_vectors:
        return SolveResult(sigma, {}, "cholqr", [sw])
    real = order[k:]
    nreal = real.size
    floor = sv[k] * T * 2.3e-16 if nreal else 0.0
    inv_s = _inv_or_zero(sv[k:], floor)
    Qt = D.gather_rows(Mr, _idx(real), row_scale=D.to_device(inv_s), cols=T)      # nreal x T, unit rows q_j^T

    def back_project(X, Lm, inv, Wt):
        """V = X^T (L^-T W), W = Wt^T (T x nreal)."""
        W = D.transpose(Wt)                                                # T x nreal
        D.trsm_lt(Lm, inv, W)
        V = D.zeros((X.shape[1], T), out_dtype)
        D.matmul(X, W, trans_a=True, out=V[:, :nreal])
        return V

    if pca:
        return SolveResult(sigma, {"left": back_project(A, LA, invA, Qt)}, "cholqr", [sw])
    # P = M Q Sigma^-1  ->  Pt = Sigma^-1 Q^T M^T
    Pt = D.matmul(Qt, M, trans_b=True)
    Pt = D.scale_copy(Pt, row_scale=D.to_device(inv_s))
    VL = back_project(A, LA, invA, Pt)
    del LA, invA, Pt
    VR = back_project(B, LB, invB, Qt)
    return SolveResult(sigma, {"left": VL, "right": VR}, "cholqr", [sw])


# ------------------------------------------------------------------ complex
def embed_complex_field(Z: np.ndarray):
    """Host helper (tests): real embedding [[X, -Y], [Y, X]] of a complex T x S field."""
    X, Y = np.ascontiguousarray(Z.real), np.ascontiguousarray(Z.imag)
    return np.block([[X, -Y], [Y, X]])


def spectrum_embedding(X, F=None):
    """Real embedding (2 Tp x 2 S) of the one-sided spectrum Z^ = F X of a centred real field
    (array.py:429-472 `_complexify` in the frequency domain, see csrc/hilbert.cu)."""
    if F is None:
        F = D.dft_matrix(X.shape[0], X.dtype)
    ZZ, _ = D.apply_time_operator(F, X)
    return D.embed_complex(ZZ)


_PAIR_STREAMS = {}


def _pair_streams():
    """Two side streams per device for the paired surrogate runs (created once)."""
    t = D.torch()
    dev = t.cuda.current_device()
    if dev not in _PAIR_STREAMS:
        _PAIR_STREAMS[dev] = (t.cuda.Stream(), t.cuda.Stream())
    return _PAIR_STREAMS[dev]


def solve_real_pair(A0, B0, A1, B1, null_basis=None, dof=None, want_vectors=True):
    """Two independent real solves of the same shape (the surrogate runs of rule_n, array.py:1753-1765).
    On the tridiagonal route the two models are processed on two streams with asynchronous reductions (two-stage
    mode), or -- one-stage mode -- their matrices are reduced by ONE batched call (xmca_sytrd_batched: half of the SMs
    each); otherwise two plain `solve_real` calls.  Returns the two results."""
    T, S1 = A0.shape
    pca = B0 is None
    S2 = S1 if pca else B0.shape[1]
    rank = min(T, S1, S2)
    same = A0.shape == A1.shape and A0.dtype == A1.dtype and (pca == (B1 is None)) and (pca or B0.shape == B1.shape)
    single = lambda A, B: solve_real(A, B, want_vectors=want_vectors, null_basis=null_basis, dof=dof)
    if not (same and TRIDIAG_MIN_N <= rank <= D.sytrd_max_n()):
        return single(A0, B0), single(A1, B1)
    if SYTRD_MODE == "two_stage":
        # Two streams: the reduction of the first model is only ENQUEUED (xmca_sytrd2 with the no-sync flag), so its
        # latency-bound parts (panel factorisations, bulge chasing: a third of the SMs at most) run under the
        # GEMM-bound construction of the second model's matrix, and the two reductions under each other.
        t = D.torch()
        cur = t.cuda.current_stream()
        s0, s1 = _pair_streams()
        for x in (A0, B0, A1, B1, null_basis):
            if x is not None and hasattr(x, "record_stream"):
                x.record_stream(s0)
                x.record_stream(s1)
        s0.wait_stream(cur)
        s1.wait_stream(cur)
        try:
            with t.cuda.stream(s0):
                r0 = TridiagResult(A0, B0, null_basis, dof, defer=True)
                r0.reduce_begin()
            with t.cuda.stream(s1):
                r1 = TridiagResult(A1, B1, null_basis, dof, defer=True)
                r1.reduce_begin()
            with t.cuda.stream(s0):
                r0.reduce_end()
            with t.cuda.stream(s1):
                r1.reduce_end()
        except np.linalg.LinAlgError:
            s0.synchronize()
            s1.synchronize()
            return single(A0, B0), single(A1, B1)
        finally:
            cur.wait_stream(s0)
            cur.wait_stream(s1)
        return r0, r1
    try:
        r0 = TridiagResult(A0, B0, null_basis, dof, defer=True)
        r1 = TridiagResult(A1, B1, null_basis, dof, defer=True)
    except np.linalg.LinAlgError:
        return single(A0, B0), single(A1, B1)
    n = r0.n
    Sp = D.empty((2, n, n), D.f64())
    Sp[0].copy_(r0._S)
    Sp[1].copy_(r1._S)
    r0._S = r1._S = None
    d, e, tau = D.sytrd_pair(Sp)
    r0._finish(d[0], e[0, :max(n - 1, 1)], tau[0], Sp[0])
    r1._finish(d[1], e[1, :max(n - 1, 1)], tau[1], Sp[1])
    return r0, r1


def solve_complex(XA, XB, want_vectors=True):
    """Complex MCA/PCA (solve(complexify=True), array.py:546-547) of the CENTRED REAL device fields
    XA, XB (T x S1, T x S2; XB may be None).  The analytic fields Z = X + i H X satisfy
    Z_A^H Z_B = Z^_A^H Z^_B with the one-sided spectra Z^ (T/2 rows, full row rank), so the real
    solver runs on the embeddings of Z^; every singular value of the embedding is double.
    Returns (sigma (rank,), ComplexVectors, embedded result)."""
    T, S1 = XA.shape
    pca = XB is None
    S2 = S1 if pca else XB.shape[1]
    rank = min(T, S1, S2)
    F = D.dft_matrix(T, XA.dtype)
    Ae = spectrum_embedding(XA, F)
    Be = None if pca else spectrum_embedding(XB, F)
    del F
    res = solve_real(Ae, Be, want_vectors=want_vectors, null_basis=0, dof=T - 1)
    s = res.sigma[0::2]
    sigma = np.zeros(rank)
    n = min(rank, s.size)
    sigma[:n] = s[:n]               # modes beyond the T/2 positive frequencies are exactly zero
    return sigma, ComplexVectors(res, n, rank), res


def solve_complex_pair(XA0, XB0, XA1, XB1, want_vectors=True):
    """Two independent complex solves of the same shape (see `solve_complex`) sharing one batched
    tridiagonalisation (`solve_real_pair`).  Returns two (sigma, ComplexVectors, embedded result) triples."""
    T, S1 = XA0.shape
    pca = XB0 is None
    S2 = S1 if pca else XB0.shape[1]
    rank = min(T, S1, S2)
    F = D.dft_matrix(T, XA0.dtype)
    emb = [(spectrum_embedding(XA, F), None if pca else spectrum_embedding(XB, F)) for XA, XB in ((XA0, XB0), (XA1, XB1))]
    del F
    out = []
    for res in solve_real_pair(emb[0][0], emb[0][1], emb[1][0], emb[1][1], null_basis=0, dof=T - 1,
                               want_vectors=want_vectors):
        s = res.sigma[0::2]
        sigma = np.zeros(rank)
        n = min(rank, s.size)
        sigma[:n] = s[:n]
        out.append((sigma, ComplexVectors(res, n, rank), res))
    return out


def solve_complex_time(XA, YA, XB, YB):
    """Complex solve from explicit analytic fields Z = X + iY (time domain), used when the Hilbert transform
    is taken on a fore/back-cast extension (extend='exp'): the cropped signal is no longer confined to the
    positive frequencies, so the real solver runs on the 2T x 2S embeddings [[X,-Y],[Y,X]] (both parts are
    centred: two exact null vectors)."""
    t = D.torch()
    T, S1 = XA.shape
    pca = XB is None
    S2 = S1 if pca else XB.shape[1]
    rank = min(T, S1, S2)
    Ae = D.embed_complex(t.cat([XA, YA], dim=0).contiguous())
    Be = None if pca else D.embed_complex(t.cat([XB, YB], dim=0).contiguous())
    nb = np.zeros((2 * T, 2))
    nb[:T, 0] = nb[T:, 1] = 1.0 / np.sqrt(T)
    res = solve_real(Ae, Be, null_basis=D.to_device(nb), dof=T - 1)
    s = res.sigma[0::2]
    sigma = np.zeros(rank)
    n = min(rank, s.size)
    sigma[:n] = s[:n]
    return sigma, ComplexVectors(res, n, rank), res


class ComplexVectors:
    """Complex singular vectors out of the real-embedded solve: each pair of equal singular
    values {x, Jx} spans one complex vector, so the even-numbered embedded vectors are taken;
    `vectors(m)` -> {field: (re, im)} device pairs, each S x m (zero columns for null modes)."""

    def __init__(self, res, n_live, rank):
        self.res, self.n_live, self.rank = res, n_live, rank
        self._cache = (0, None)

    def vectors(self, m):
        m = int(min(max(m, 0), self.rank))
        have, V = self._cache
        if V is None or have < m:
            live = min(m, self.n_live)
            Ve_all = self.res.vectors(2 * live)
            pick = _idx(np.arange(0, 2 * live, 2))
            V = {}
            for k, Ve in Ve_all.items():
                S = Ve.shape[0] // 2
                Vt = D.transpose(Ve)                                          # 2 live x 2S rows
                rows = D.gather_rows(Vt, pick)                                # live x 2S
                cols = D.transpose(rows)                                      # 2S x live
                if live < m:
                    full = D.zeros((2 * S, m), cols.dtype)
                    full[:, :live].copy_(cols)
                    cols = full
                V[k] = (cols[:S], cols[S:])
            self._cache = (m, V)
        return {k: (v[0][:, :m], v[1][:, :m]) for k, v in V.items()}


# ------------------------------------------------------------------ rotation
# The fused rotation kernels keep the p x p state of one iteration in shared memory: p <= 64 (real) / 32 (complex).
# The reference accepts any n_rot (array.py:810-813): wider rotations run the same fixed point as device products
# (n x p x p) with the p x p SVD of each iteration on the host -- correct for any p, a few hundred microseconds per
# iteration instead of tens.
VARIMAX_FUSED_MAX_P = 64
VARIMAX_FUSED_MAX_P_COMPLEX = 32


def varimax_wide(Ld, gamma=1.0, max_iter=1000, tol=1e-8):
    """Varimax (rotation.py:15-78) for any p: B = A R and A^T B^3 as device products, the polar factor of the
    p x p criterion matrix by numpy (rotation.py:59-61).  Uses A^T (B diag c) = G R diag c with G = A^T A,
    c = diag(R^T G R).  Returns (B fp64 device n x p, R host p x p, iterations)."""
    n, p = Ld.shape
    h = np.sqrt(D.to_host(D.row_sumsq(Ld)))                               # rotation.py:46
    inv_h = D.to_device(np.where(h > 0, 1.0 / np.where(h > 0, h, 1.0), 0.0))
    A = D.scale_copy(Ld, out_dtype=D.f64(), row_scale=inv_h)              # rotation.py:48
    G = D.to_host(D.matmul(A, A, trans_a=True))
    ones_r, ones_c = D.to_device(np.ones(n)), D.to_device(np.ones(p))
    R = np.eye(p)
    d = 0.0
    for it in range(1, max_iter + 1):
        d_old = d
        B = D.matmul(A, D.to_device(R))                                   # rotation.py:54
        _, B3 = D.promax_target(B, ones_r, ones_c, 3)                     # B |B|^2
        T1 = D.to_host(D.matmul(A, B3, trans_a=True))
        GR = G @ R
        c = np.einsum("ij,ij->j", R, GR)                                  # column sums of B^2
        crit = T1 - (gamma / n) * GR * c
        u, sv, vh = np.linalg.svd(crit)                                   # rotation.py:59
        R = u @ vh
        d = sv.sum()
        if abs(d - d_old) / d < tol:                                      # rotation.py:62
            Bout = D.matmul(D.scale_copy(A, row_scale=D.to_device(h)), D.to_device(R))    # rotation.py:74-77
            return Bout, R, it
    raise L_NotConverged("Rotation process did not converge.")


def varimax_complex_wide(Lr, Li, gamma=1.0, max_iter=1000, tol=1e-8):
    """Complex Varimax for any p (planar loadings): the same fixed point with complex products as four real ones."""
    t = D.torch()
    n, p = Lr.shape
    h = np.sqrt(D.to_host(D.row_sumsq(t.cat([Lr, Li], dim=1).contiguous())))
    inv_h = D.to_device(np.where(h > 0, 1.0 / np.where(h > 0, h, 1.0), 0.0))
    Ar = D.scale_copy(Lr, out_dtype=D.f64(), row_scale=inv_h)
    Ai = D.scale_copy(Li, out_dtype=D.f64(), row_scale=inv_h)
    G = _cgemm_tn(Ar, Ai, Ar, Ai)
    ones_r, ones_c = D.to_device(np.ones(n)), D.to_device(np.ones(p))
    R = np.eye(p, dtype=np.complex128)
    d = 0.0
    for it in range(1, max_iter + 1):
        d_old = d
        Br, Bi = _cgemm_right(Ar, Ai, R)
        _, _, Pr, Pi = D.promax_target_complex(Br, Bi, ones_r, ones_c, 3)           # B |B|^2
        T1 = _cgemm_tn(Ar, Ai, Pr, Pi)
        GR = G @ R
        c = np.einsum("ij,ij->j", R.conj(), GR).real
        crit = T1 - (gamma / n) * GR * c
        u, sv, vh = np.linalg.svd(crit)
        R = u @ vh
        d = sv.sum()
        if abs(d - d_old) / d < tol:
            hd = D.to_device(h)
            Br, Bi = _cgemm_right(D.scale_copy(Ar, row_scale=hd), D.scale_copy(Ai, row_scale=hd), R)
            return Br, Bi, R, it
    raise L_NotConverged("Rotation process did not converge.")


def promax(Ld, power=1, max_iter=1000, tol=1e-8):
    """Device Promax (rotation.py:84-149), real loadings Ld (n x p).
    Returns (B fp64 device n x p, R host p x p, Phi host p x p, iterations)."""
    n, p = Ld.shape
    if p > VARIMAX_FUSED_MAX_P:
        B, R, iters = varimax_wide(Ld, 1.0, max_iter, tol)
    else:
        B, Rdev, iters = D.varimax(Ld, 1.0, max_iter, tol)
        R = D.to_host(Rdev)
    if power == 1:
        # the Promax step is the identity for power 1 (L = I up to rounding)
        return B, R, np.eye(p), iters
    h2 = D.to_host(D.row_sumsq(B))
    h = np.sqrt(h2)
    inv_h = D.to_device(1.0 / h)
    colmax = D.col_absmax(B, row_scale=inv_h)
    X, Pm = D.promax_target(B, inv_h, colmax, power)
    XtX = D.to_host(D.matmul(X, X, trans_a=True))                         # p x p
    XtP = D.to_host(D.matmul(X, Pm, trans_a=True))
    Lm = np.linalg.inv(XtX) @ XtP                                         # rotation.py:128
    try:
        dinv = np.diag(np.linalg.inv(Lm.T @ Lm))                          # rotation.py:131-134
    except np.linalg.LinAlgError:
        dinv = np.diag(np.linalg.pinv(Lm.T @ Lm))
    Lm = Lm @ np.sqrt(np.diag(dinv))                                      # rotation.py:137
    XL = D.matmul(X, D.to_device(Lm))                                     # n x p
    Bp = D.scale_copy(XL, row_scale=D.to_device(h))                       # rotation.py:141
    Li = np.linalg.inv(Lm)
    return Bp, R @ Lm, Li @ Li.T, iters


def _cgemm_tn(Ar, Ai, Br, Bi):
    """A^H B for planar complex n x p operands -> host complex p x p (four real products)."""
    rr = D.to_host(D.matmul(Ar, Br, trans_a=True)) + D.to_host(D.matmul(Ai, Bi, trans_a=True))
    ii = D.to_host(D.matmul(Ar, Bi, trans_a=True)) - D.to_host(D.matmul(Ai, Br, trans_a=True))
    return rr + 1j * ii


def _cgemm_right(Xr, Xi, M):
    """X M for planar complex X (n x p, device) and a host complex p x p matrix M -> planar device."""
    Mr, Mi = D.to_device(np.ascontiguousarray(M.real)), D.to_device(np.ascontiguousarray(M.imag))
    Yr = D.matmul(Xr, Mr)
    D.matmul(Xi, Mi, alpha=-1.0, out=Yr, accumulate=True)
    Yi = D.matmul(Xr, Mi)
    D.matmul(Xi, Mr, out=Yi, accumulate=True)
    return Yr, Yi


def rotate_complex(V, sigma, keys, n_rot, power=1, max_iter=1000, tol=1e-8):
    """Varimax / Promax rotation of complex loadings L = [V_L; V_R] sqrt(sigma) (array.py:815-833 and
    rotation.py:84-149 with complex dtype).  V: {field: (re, im)} device pairs (S x >= n_rot).
    Returns (Br, Bi, s_left, R, Phi, iterations); norms follow from the column sums of |B|^2."""
    t = D.torch()
    root = D.to_device(np.sqrt(np.asarray(sigma[:n_rot], dtype=np.float64)))
    re = [D.scale_copy(V[k][0][:, :n_rot], col_scale=root) for k in keys]
    im = [D.scale_copy(V[k][1][:, :n_rot], col_scale=root) for k in keys]
    s_left = re[0].shape[0]
    Lr = t.cat(re, dim=0).contiguous() if len(re) > 1 else re[0]
    Li = t.cat(im, dim=0).contiguous() if len(im) > 1 else im[0]
    if n_rot > VARIMAX_FUSED_MAX_P_COMPLEX:
        Br, Bi, R, iters = varimax_complex_wide(Lr, Li, 1.0, max_iter, tol)
    else:
        Br, Bi, R, iters = D.varimax_complex(Lr, Li, 1.0, max_iter, tol)
    p = n_rot
    if power == 1:
        return Br, Bi, s_left, R, np.eye(p), iters
    # Promax step (rotation.py:115-147): streaming passes on the device, p x p algebra on the host
    h = np.sqrt(D.to_host(D.row_sumsq(t.cat([Br, Bi], dim=1).contiguous())))
    inv_h = D.to_device(1.0 / h)
    colmax = D.col_absmax_complex(Br, Bi, row_scale=inv_h)
    Xr, Xi, Pr, Pi = D.promax_target_complex(Br, Bi, inv_h, colmax, power)
    XhX = _cgemm_tn(Xr, Xi, Xr, Xi)
    XhP = _cgemm_tn(Xr, Xi, Pr, Pi)
    Lm = np.linalg.inv(XhX) @ XhP                                         # rotation.py:128
    try:
        dinv = np.diag(np.linalg.inv(Lm.conj().T @ Lm))                   # rotation.py:131-134
    except np.linalg.LinAlgError:
        dinv = np.diag(np.linalg.pinv(Lm.conj().T @ Lm))
    Lm = Lm @ np.sqrt(np.diag(dinv))                                      # rotation.py:137
    Yr, Yi = _cgemm_right(Xr, Xi, Lm)
    hd = D.to_device(h)
    Br, Bi = D.scale_copy(Yr, row_scale=hd), D.scale_copy(Yi, row_scale=hd)   # rotation.py:141
    Li_ = np.linalg.inv(Lm)
    return Br, Bi, s_left, R @ Lm, Li_ @ Li_.conj().T, iters


def complex_col_norms(Br, Bi, row0, row1):
    """sqrt(sum_rows |B|^2) per column over rows [row0, row1) (array.py:826-830)."""
    return np.sqrt(D.to_host(D.col_sumsq(Br, row0, row1)) + D.to_host(D.col_sumsq(Bi, row0, row1)))
