"""numpy statement of the polar-factor step that varimax.cu runs inside the Varimax iteration (NOT on the product
path itself: the kernel is the product, this is its arithmetic written down for review and CPU tests).

For X with nearly orthogonal columns (cosines <= ~1e-3):  polar(X) = X G^{-1/2},  G = X^T X = D + E,
D = diag(s_j^2), and G^{-1/2} to second order in E through the divided differences of f(x) = x^{-1/2}:
    f[a, b]    = -1 / (ra rb (ra + rb))
    f[a, b, c] = (ra + rb + rc) / (ra rb rc (ra + rb)(rb + rc)(ra + rc)),       r. = sqrt(.)
    Z_ij = [i = j] / s_i  -  Et_ij / (s_i s_j)  +  sum_k Et_ik Et_kj (s_i + s_j + s_k) / (s_k s_i s_j (s_i + s_j)),
    Et_ij = E_ij / (s_i + s_j)
The error is third order in the cosines; trace(Z G) = sum of the singular values to the same order.
"""
import numpy as np


def gram_inv_sqrt2(G):
    s = np.sqrt(np.diag(G))
    E = G - np.diag(np.diag(G))
    S = np.add.outer(s, s)
    Et = E / S
    F = (Et / s[None, :]) @ Et                    # sum_k Et_ik Et_kj / s_k
    H = Et @ Et
    rij = 1.0 / np.outer(s, s)
    Z = np.diag(1.0 / s) - Et * rij + (S * F + H) * rij / S
    return Z, float(np.sum(Z * G))


def polar_by_expansion(X):
    Z, d = gram_inv_sqrt2(X.T @ X)
    return X @ Z, d


def nearly_orthogonal(p, eps, rng, decades=3.0):
    """p x p matrix whose columns have pairwise cosines of order eps and norms spread over `decades`."""
    U, _ = np.linalg.qr(rng.standard_normal((p, p)))
    Y = U + eps * rng.standard_normal((p, p)) / np.sqrt(p)
    Y /= np.linalg.norm(Y, axis=0)
    return Y * np.logspace(0, -decades, p)


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    p = 50
    for eps in (1e-2, 1e-3, 1e-4):
        X = nearly_orthogonal(p, eps, rng)
        G = X.T @ X
        sn = np.sqrt(np.diag(G))
        C = G / np.outer(sn, sn) - np.eye(p)
        P, d = polar_by_expansion(X)
        u, sv, vt = np.linalg.svd(X)
        print("largest cosine %.1e: |P - polar(X)| = %.1e, |P^T P - I| = %.1e, d rel err %.1e"
              % (np.abs(C).max(), np.abs(P - u @ vt).max(), np.abs(P.T @ P - np.eye(p)).max(), abs(d - sv.sum()) / sv.sum()))
