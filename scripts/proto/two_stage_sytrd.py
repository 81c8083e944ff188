"""Design prototype (numpy, NOT part of the shipped path, imported by nothing): two-stage reduction of a
symmetric matrix to tridiagonal form, the planned replacement of the one-stage `xmca_sytrd` for single solves.

  stage 1  dense -> band (bandwidth b): blocked Householder QR of each (m x b) sub-diagonal panel, two-sided
           compact-WY update of the trailing matrix -- everything but the panel QR is GEMM shaped
           (W = A22 V T, X = W - 1/2 V (T^T (V^T W)), A22 -= X V^T + V X^T): 4/3 n^3 flops on the fp64 DMMA pipe
           instead of n^3 8/6 bytes of latency-ridden matrix-vector passes.
  stage 2  band -> tridiagonal by bulge chasing (Bischof / Lang / Sun): per column one length-b reflector and a
           chain of (n - j) / b chase steps on b x 3b windows: 6 n^2 b flops, L2 / shared-memory resident,
           parallel as a wavefront of sweeps (sweep j + 1 may run two windows behind sweep j).

Planned GPU mapping (round 2):
  stage 1  panel QR by one cooperative kernel (TSQR-free: the panel is m x 64, column norms by grid reduction as in
           today's sytrd_panel_kernel but on 64 columns of a TALL panel only -- O(m b) bytes per column instead of
           O(m^2)); W = A22 V T and the rank-2b update on xmca_gemm_ex (DMMA), A22 kept tile-major / lower only.
  stage 2  band in a (2b + 1) x n array (8.4 MB at n = 8192, b = 64: L2 resident).  One persistent CTA per sweep
           (sweeps handed out in order from an atomic counter); task k of sweep j waits on a progress flag until
           sweep j - 1 has finished task k + 2, then generates its length-b reflector (warp reduction) and applies it
           to its b x 3b window from both sides in shared memory, publishes its own progress (release store).
           ~n / (3b) sweeps are in flight; ~64 tasks of 1-2 us per sweep.  Eigenvalues-only callers (rule_n, the
           unrotated spectrum) stop here; for vectors the reflectors are stored (n^2 / 2 doubles) and applied to
           the m requested vectors as diamond-shaped compact-WY blocks (b sweeps x one chase position each).

Run:  python scripts/proto/two_stage_sytrd.py [n] [b]   -> checks the eigenvalues of both stages against numpy and
prints the flop / task counts the DESIGN.md estimate uses.
"""
import sys
import numpy as np


def house(x):
    """v (v[0] = 1), tau, beta with (I - tau v v^T) x = beta e_0  (LAPACK dlarfg convention)."""
    alpha, xnorm = x[0], np.linalg.norm(x[1:])
    if xnorm == 0.0:
        v = np.zeros_like(x); v[0] = 1.0
        return v, 0.0, alpha
    beta = -np.copysign(np.hypot(alpha, xnorm), alpha)
    v = x / (alpha - beta); v[0] = 1.0
    return v, (beta - alpha) / beta, beta


def panel_qr(P):
    """Householder QR of the m x b panel: returns V (m x b, unit lower trapezoidal), T (b x b upper, compact WY:
    Q = I - V T V^T) and R (b x b upper triangular)."""
    m, b = P.shape
    P = P.copy()
    V = np.zeros((m, b)); taus = np.zeros(b)
    for j in range(min(b, m)):
        v, tau, beta = house(P[j:, j])
        V[j:, j] = v; taus[j] = tau
        P[j:, j:] -= tau * np.outer(v, v @ P[j:, j:])
    T = np.zeros((b, b))
    for j in range(b):                       # dlarft, forward / columnwise
        T[j, j] = taus[j]
        if j:
            T[:j, j] = -taus[j] * (T[:j, :j] @ (V[:, :j].T @ V[:, j]))
    return V, T, np.triu(P[:b, :])


def stage1_dense_to_band(A, b):
    """Returns the band matrix (full storage, entries beyond bandwidth b zero) and the flop count of the GEMMs."""
    A = A.copy(); n = A.shape[0]; flops = 0
    for k in range(0, n - b - 1, b):
        m = n - k - b
        V, T, R = panel_qr(A[k + b:, k:k + b])
        A[k + b:, k:k + b] = 0.0
        A[k + b:k + b + R.shape[0], k:k + b] = R[:min(m, b), :]
        A[k:k + b, k + b:] = A[k + b:, k:k + b].T
        A22 = A[k + b:, k + b:]
        W = A22 @ V @ T                                           # symm: 2 m^2 b
        X = W - 0.5 * V @ (T.T @ (V.T @ W))
        A22 -= X @ V.T + V @ X.T                                  # syr2k: 2 m^2 b
        flops += 4 * m * m * b
    return A, flops


def stage2_band_to_tridiag(B, b):
    """Bulge chasing on the full-storage band matrix; returns (d, e) and the number of chase tasks (windows)."""
    B = B.copy(); n = B.shape[0]; tasks = 0
    for j in range(n - 2):
        # eliminate column j below the sub-diagonal, then chase the bulge down the band
        col, r0 = j, j + 1
        while r0 < n - 1:
            r1 = min(r0 + b, n)
            x = B[r0:r1, col]
            if np.all(x[1:] == 0.0):
                break
            v, tau, beta = house(x)
            # two-sided application on the window touched by rows / columns r0:r1 (everything else is zero there)
            lo, hi = max(col, r0 - b), min(n, r1 + b)
            Hw = B[r0:r1, lo:hi]
            Hw -= tau * np.outer(v, v @ Hw)
            Hc = B[lo:hi, r0:r1]
            Hc -= tau * np.outer(Hc @ v, v)
            tasks += 1
            # the reflector filled a bulge below the band in columns r0:r1: its first column is chased next
            col, r0 = r0, r1
    return np.diag(B).copy(), np.diag(B, -1).copy(), tasks


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 192
    b = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    rng = np.random.default_rng(0)
    S = rng.standard_normal((n, n)); S = S @ S.T / n
    ref = np.linalg.eigvalsh(S)
    Bm, flops = stage1_dense_to_band(S, b)
    i, j = np.indices((n, n))
    assert np.abs(Bm[np.abs(i - j) > b]).max() < 1e-12 * np.abs(S).max(), "stage 1 left entries outside the band"
    print("stage 1: band eigenvalues vs numpy %.2e, GEMM flops %.3g (4/3 n^3 = %.3g)"
          % (np.abs(np.linalg.eigvalsh(Bm) - ref).max() / ref.max(), flops, 4 / 3 * n ** 3))
    d, e, tasks = stage2_band_to_tridiag(Bm, b)
    Tm = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    print("stage 2: tridiagonal eigenvalues vs numpy %.2e, chase tasks %d (n^2 / (2 b) = %d)"
          % (np.abs(np.linalg.eigvalsh(Tm) - ref).max() / ref.max(), tasks, n * n // (2 * b)))


# ----------------------------------------------------------------------------------------------------------------
# Stage 2 again, on the storage the kernel will use: AB[d, j] = A[j + d, j], d = 0 .. 2b (lower band + bulge room).
# One task = three dense sub-blocks around the reflector's index range I = [r0, r1):
#   left   A[I, lo:r0]   <- H A[I, lo:r0]        (b x <= b; its first column is the one being eliminated)
#   diag   A[I, I]       <- H A[I, I] H          (symmetric b x b, lower part stored)
#   below  A[r1:hi, I]   <- A[r1:hi, I] H        (<= b x b; fills the next bulge)
def to_band_storage(B, b):
    n = B.shape[0]
    AB = np.zeros((2 * b + 1, n))
    for d in range(2 * b + 1):
        AB[d, :n - d] = np.diag(B, -d)
    return AB


def stage2_band_storage(AB, b):
    AB = AB.copy(); n = AB.shape[1]

    def get(r, c):            # r >= c
        return AB[r - c, c]

    for j in range(n - 2):
        col, r0 = j, j + 1
        while r0 < n - 1:
            r1 = min(r0 + b, n); m = r1 - r0
            x = np.array([get(r0 + i, col) for i in range(m)])
            if np.all(x[1:] == 0.0):
                break
            v, tau, beta = house(x)
            lo, hi = max(col, r0 - b), min(n, r1 + b)
            # left block: rows I, columns lo .. r0 - 1
            for c in range(lo, r0):
                colv = np.array([get(r0 + i, c) for i in range(m)])
                colv -= tau * v * (v @ colv)
                for i in range(m):
                    AB[r0 + i - c, c] = colv[i]
            # diagonal block (symmetric, two-sided): D <- D - v w^T - w v^T,  w = tau D v - (tau^2 v^T D v / 2) v
            D = np.zeros((m, m))
            for i in range(m):
                for k in range(i + 1):
                    D[i, k] = D[k, i] = get(r0 + i, r0 + k)
            p = tau * (D @ v)
            w = p - 0.5 * tau * (v @ p) * v
            D -= np.outer(v, w) + np.outer(w, v)
            for i in range(m):
                for k in range(i + 1):
                    AB[i - k, r0 + k] = D[i, k]
            # block below: rows r1 .. hi - 1, columns I
            for r in range(r1, hi):
                rowv = np.array([get(r, r0 + k) for k in range(m)])
                rowv -= tau * (rowv @ v) * v
                for k in range(m):
                    AB[r - r0 - k, r0 + k] = rowv[k]
            col, r0 = r0, r1
    return AB[0].copy(), AB[1, :n - 1].copy()


if __name__ == "__main__":
    d2, e2 = stage2_band_storage(to_band_storage(Bm, b), b)
    T2 = np.diag(d2) + np.diag(e2, 1) + np.diag(e2, -1)
    print("stage 2 on band storage (left / diagonal / below blocks): eigenvalues vs numpy %.2e; max |d - d_full| %.2e"
          % (np.abs(np.linalg.eigvalsh(T2) - ref).max() / ref.max(), np.abs(d2 - d).max()))


# ----------------------------------------------------------------------------------------------------------------
# Eigenvectors: S = Q1 B Q1^T (stage 1, block reflectors per panel), B = Q2 T Q2^T (stage 2, one reflector per
# task, in task order).  An eigenvector z of T goes back as x = Q1 (Q2 z): the stage-2 reflectors are applied in
# REVERSE task order (each touches rows r0:r1 only -- reflectors of one chase position from b consecutive sweeps
# touch overlapping ranges and can be merged into one compact-WY block: the "diamond" blocking), then the stage-1
# block reflectors in reverse panel order (GEMMs).
def two_stage_with_vectors(S, b, m):
    n = S.shape[0]
    A = S.copy(); q1 = []
    for k in range(0, n - b - 1, b):
        V, T, R = panel_qr(A[k + b:, k:k + b])
        q1.append((k + b, V, T))
        A[k + b:, k:k + b] = 0.0
        A[k + b:k + b + R.shape[0], k:k + b] = R[:min(n - k - b, b), :]
        A[k:k + b, k + b:] = A[k + b:, k:k + b].T
        A22 = A[k + b:, k + b:]
        W = A22 @ V @ T
        X = W - 0.5 * V @ (T.T @ (V.T @ W))
        A22 -= X @ V.T + V @ X.T
    q2 = []
    for j in range(n - 2):
        col, r0 = j, j + 1
        while r0 < n - 1:
            r1 = min(r0 + b, n)
            x = A[r0:r1, col]
            if np.all(x[1:] == 0.0):
                break
            v, tau, beta = house(x.copy())
            lo, hi = max(col, r0 - b), min(n, r1 + b)
            Hw = A[r0:r1, lo:hi]; Hw -= tau * np.outer(v, v @ Hw)
            Hc = A[lo:hi, r0:r1]; Hc -= tau * np.outer(Hc @ v, v)
            q2.append((r0, v, tau))
            col, r0 = r0, r1
    d, e = np.diag(A).copy(), np.diag(A, -1).copy()
    lam, Z = np.linalg.eigh(np.diag(d) + np.diag(e, 1) + np.diag(e, -1))
    lam, Z = lam[::-1][:m], Z[:, ::-1][:, :m].copy()
    for r0, v, tau in reversed(q2):                       # x <- H x  on rows r0 : r0 + len(v)
        blk = Z[r0:r0 + v.size]
        blk -= tau * np.outer(v, v @ blk)
    for r0, V, T in reversed(q1):                         # x <- (I - V T V^T) x  on rows r0 :
        blk = Z[r0:]
        blk -= V @ (T @ (V.T @ blk))
    return lam, Z


if __name__ == "__main__":
    lam, X = two_stage_with_vectors(S, b, 6)
    print("vectors through both stages: max |S x - lambda x| / lambda_max = %.2e, |X^T X - I| = %.2e"
          % (np.abs(S @ X - X * lam).max() / lam[0], np.abs(X.T @ X - np.eye(6)).max()))
