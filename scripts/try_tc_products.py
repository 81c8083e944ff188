"""Experiment: S = L_B^T G_A L_B / dof^2 with the two n^3 products on the tensor cores (3xTF32, fp64 chunk sums) instead of
the fp64 DMMA kernel -- accuracy of the spectrum and time.  python scripts/try_tc_products.py [T S]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from xmca_b200 import device as D, _lib as L

T, S = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (8192, 16384)
A, B = bench.synthetic_fields(T, S, S, seed=1000)
A = A - A.mean(axis=0); B = B - B.mean(axis=0)
Ad, Bd = D.to_device(A), D.to_device(B)
dof = T - 1.0
lib = L.load()

def tc_nt_f64(Ah, Al, Bh, Bl, M, N, K, alpha, symmetric):
    G = D.empty((M, N), D.f64())
    rc = lib.xmca_tc_gemm_nt_f64(M, N, K, float(alpha), L.ptr(Ah), L.ptr(Al), Ah.stride(0), L.ptr(Bh), L.ptr(Bl),
                                 Bh.stride(0), L.ptr(G), N, 1 if symmetric else 0, L.stream_ptr())
    L.check(rc, "xmca_tc_gemm_nt_f64")
    return G

def timed(fn):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = fn(); e1.record(); torch.cuda.synchronize(); return out, e0.elapsed_time(e1)

GA = D.gram_tc(Ad); GB = D.gram_tc(Bd)
tr = float(D.to_host(D.col_sumsq(Bd)).sum())
Nb = D.to_device(np.full((T, 1), 1.0 / np.sqrt(T)))
D.matmul(Nb, Nb, trans_b=True, alpha=tr / T, out=GB, accumulate=True)
LB, invB = D.cholesky(GB, min_pivot=1e-11 * tr / T)

def ref_path():
    W = D.matmul(GA, LB, b_lower=True)
    return D.matmul(LB, W, trans_a=True, alpha=1.0 / dof ** 2, symmetric=True, a_lower_t=True)

def tc_path():
    gh, gl, _ = D.split_tf32(GA)                       # G_A rows (M x K)
    lth, ltl, _ = D.split_tf32(LB, transpose=True)     # L_B^T rows: [n][k] = L_B[k][n]
    W = tc_nt_f64(gh, gl, lth, ltl, T, T, T, 1.0, False)          # W = G_A L_B
    wth, wtl, _ = D.split_tf32(W, transpose=True)      # W^T rows
    return tc_nt_f64(lth, ltl, wth, wtl, T, T, T, 1.0 / dof ** 2, True)   # S = L_B^T W

for _ in range(2):
    S_ref, t_ref = timed(ref_path)
    S_tc, t_tc = timed(tc_path)
print("products: fp64 DMMA %.1f ms, tensor cores (incl. 3 operand splits) %.1f ms" % (t_ref, t_tc))
print("max |S_tc - S_ref| / max|S| = %.2e" % float((S_tc - S_ref).abs().max() / S_ref.abs().max()))
def spectrum(Sm):
    d, e, tf = D.sytrd2(Sm.clone(), want_vectors=False)
    return np.sqrt(np.maximum(D.to_host(D.stebz(d, e)), 0.0))
s_ref, s_tc = spectrum(S_ref), spectrum(S_tc)
rel = np.abs(s_tc - s_ref) / s_ref[0]
relk = np.abs(s_tc - s_ref) / np.maximum(s_ref, 1e-300)
print("sigma: max rel diff of modes 0..49: %.2e; modes 50..999: %.2e; all: %.2e (relative to sigma_1: %.2e)"
      % (relk[:50].max(), relk[50:1000].max(), relk[:T - 2].max(), rel.max()))
# against numpy on a float64 direct SVD? too large here; the reference point is the fp64-product path itself
print("sigma_1 %.6e sigma_50 %.6e sigma_1000 %.6e" % (s_ref[0], s_ref[49], s_ref[999]))
