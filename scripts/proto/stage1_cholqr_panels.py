"""Design prototype (numpy, NOT shipped): can the panel factorisation of stage 1 (dense -> band) be made entirely
GEMM shaped?  Panel P (m x b) -> orthonormal Q by CholeskyQR2 (two Gram + Cholesky + triangular-solve passes), then
the compact-WY form (Y, T) of an orthogonal W = I - Y T Y^T with W^T P = [R; 0] by Householder RECONSTRUCTION
(Ballard, Demmel, Grigori, Jacquelin, Knight, Nguyen 2015): LU without pivoting of [I; 0] - Q S with the signs S
chosen on the fly so that every pivot is >= 1.  The question is numerical: CholeskyQR needs cond(P)^2 < 1/eps.
This script runs stage 1 that way on the engine's own kind of matrix (S = L_B^T G_A L_B / dof^2 of synthetic
low-rank-plus-noise fields) and on a graded SPD matrix, and reports breakdowns and the accuracy of the band's spectrum.

Run:  python scripts/proto/stage1_cholqr_panels.py
"""
import numpy as np


def cholqr2(P):
    """Returns (Q, R, ok).  ok = False if a Cholesky factorisation broke down."""
    Q, Rt = P, np.eye(P.shape[1])
    for _ in range(2):
        G = Q.T @ Q
        try:
            L = np.linalg.cholesky(G)
        except np.linalg.LinAlgError:
            return None, None, False
        Q = np.linalg.solve(L, Q.T).T                  # Q <- Q L^-T
        Rt = L.T @ Rt
    return Q, Rt, True


def reconstruct_wy(Q):
    """Y (m x b unit lower trapezoidal), T (b x b upper), S (signs) with (I - Y T Y^T)[I; 0] = Q S."""
    m, b = Q.shape
    A = -Q.copy()
    S = np.ones(b)
    for j in range(b):                                  # LU without pivoting of [I; 0] - Q S, sign chosen per column
        # current pivot candidate of column j is s_j * (-Q~)_jj + 1 ; choose s_j so that it is >= 1
        S[j] = 1.0 if A[j, j] >= 0.0 else -1.0          # A holds -Q (eliminated so far); pivot = 1 + s_j A_jj s.t. >= 1
        A[:, j] *= S[j]
        A[j, j] += 1.0
        A[j + 1:, j] /= A[j, j]
        A[j + 1:, j + 1:] -= np.outer(A[j + 1:, j], A[j, j + 1:])
    Y = np.tril(A, -1)
    Y[np.arange(b), np.arange(b)] = 1.0
    U = np.triu(A[:b])
    T = U @ np.linalg.inv(Y[:b].T)
    return Y, T, S


def stage1(Sm, b, verbose=False):
    A = Sm.copy(); n = A.shape[0]; breakdowns = 0; worst_orth = 0.0
    for k in range(0, n - b - 1, b):
        P = A[k + b:, k:k + b]
        Q, R, ok = cholqr2(P)
        if not ok:
            breakdowns += 1
            Q, R = np.linalg.qr(P)                      # (the fallback a kernel would need: Householder panel QR)
        Y, T, S = reconstruct_wy(Q)
        W = np.eye(P.shape[0]) - Y @ T @ Y.T
        worst_orth = max(worst_orth, np.abs(W.T @ W - np.eye(P.shape[0])).max())
        Rt = (S[:, None] * R)
        A[k + b:, k:k + b] = 0.0
        A[k + b:k + b + min(b, P.shape[0]), k:k + b] = Rt[:min(b, P.shape[0])]
        A[k:k + b, k + b:] = A[k + b:, k:k + b].T
        A22 = A[k + b:, k + b:]
        Wm = A22 @ Y @ T
        X = Wm - 0.5 * Y @ (T.T @ (Y.T @ Wm))
        A22 -= X @ Y.T + Y @ X.T
    return A, breakdowns, worst_orth


def engine_matrix(T_, S_, seed=0, k=16):
    r = np.random.default_rng(seed)
    ts = r.standard_normal((T_, k)); a = 3 * np.sqrt(S_) * 0.9 ** np.arange(k) / np.sqrt(k)
    A = (ts * a) @ (r.standard_normal((k, S_)) / np.sqrt(S_)) + r.standard_normal((T_, S_))
    B = (ts * a) @ (r.standard_normal((k, S_)) / np.sqrt(S_)) + r.standard_normal((T_, S_))
    A -= A.mean(0); B -= B.mean(0)
    GA, GB = A @ A.T, B @ B.T
    GB += np.trace(GB) / T_ * np.ones((T_, T_)) / T_
    LB = np.linalg.cholesky(GB)
    return LB.T @ GA @ LB / (T_ - 1) ** 2


if __name__ == "__main__":
    b = 16
    for name, M in (("engine S (T 320, S 640)", engine_matrix(320, 640)),
                    ("graded SPD, cond 1e12", (lambda Q, lam: (Q * lam) @ Q.T)(
                        np.linalg.qr(np.random.default_rng(1).standard_normal((320, 320)))[0], np.logspace(0, -12, 320)))):
        M = 0.5 * (M + M.T)
        ref = np.linalg.eigvalsh(M)
        Bm, bd, orth = stage1(M, b)
        i, j = np.indices(M.shape)
        print("%s: CholeskyQR2 breakdowns %d of %d panels, worst |W^T W - I| %.1e, band spectrum error %.1e, outside band %.1e"
              % (name, bd, len(range(0, M.shape[0] - b - 1, b)), orth,
                 np.abs(np.linalg.eigvalsh(Bm) - ref).max() / ref.max(), np.abs(Bm[np.abs(i - j) > b]).max() / ref.max()))
