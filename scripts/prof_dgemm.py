"""Time the fp64 DMMA GEMM paths (xmca_gemm_ex): symmetric Gram X X^T and a general product.  python scripts/prof_dgemm.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xmca_b200 import device as D
g = torch.Generator(device="cuda").manual_seed(0)
def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return out, best
for T, S in ((4096, 2048), (8192, 2048), (8192, 16384)):
    X = torch.randn((T, S), device="cuda", dtype=torch.float64, generator=g)
    G, ms = timed(lambda: D.matmul(X, X, trans_b=True, symmetric=True))
    print("symmetric Gram X X^T  T=%d S=%d f64 (X: %d MB): %.2f ms = %.1f TFLOP/s" % (T, S, T * S * 8 >> 20, ms, T * T * S / ms / 1e9))
    G2, ms2 = timed(lambda: D.matmul(X, X, trans_b=True))
    print("   same product without the symmetric shortcut: %.2f ms = %.1f TFLOP/s" % (ms2, 2.0 * T * T * S / ms2 / 1e9))
ref = X[:256] @ X[:256].T
print("   max rel err of a corner vs torch: %.2e" % float((G[:256, :256] - ref).abs().max() / ref.abs().max()))
n = 8192
A = torch.randn((n, n), device="cuda", dtype=torch.float64, generator=g)
B = torch.randn((n, n), device="cuda", dtype=torch.float64, generator=g)
C, ms = timed(lambda: D.matmul(A, B))
print("general A B  n=%d f64: %.2f ms = %.1f TFLOP/s" % (n, ms, 2.0 * n ** 3 / ms / 1e9))
C2, ms2 = timed(lambda: A @ B)
print("torch (cuBLAS) A B  n=%d f64: %.2f ms = %.1f TFLOP/s; max rel diff %.2e" % (n, ms2, 2.0 * n ** 3 / ms2 / 1e9, float((C - C2).abs().max() / C2.abs().max())))
