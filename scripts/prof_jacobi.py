"""Profiling driver: blocked-Jacobi SVD of (a) the Gram matrix G = A A^T of a synthetic
field (T = n, S = 2n; symmetric PSD, no rotations accumulated) or (b) the kernel
K = F_A^T F_B-like general matrix.  torch is used only to BUILD the input.
Usage: python scripts/prof_jacobi.py N [max_sweeps] [kind: gram|general|wishart] [want_v]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xmca_b200 import device as D, _lib
from bench import synthetic_fields

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
kind = sys.argv[3] if len(sys.argv) > 3 else "gram"
want_v = (sys.argv[4] != "0") if len(sys.argv) > 4 else (kind != "gram")
if kind == "wishart":
    g = torch.Generator(device="cuda").manual_seed(0)
    X = torch.randn((n, n), dtype=torch.float64, device="cuda", generator=g)
    X = X @ X.T / n
else:
    A, B = synthetic_fields(n, 2 * n, 2 * n, seed=3)
    A -= A.mean(axis=0); B -= B.mean(axis=0)
    Ad = torch.from_numpy(A).cuda().double()
    X = Ad @ Ad.T
    if kind == "general":
        Bd = torch.from_numpy(B).cuda().double()
        X = (Ad @ Bd.T) / n            # general (non-symmetric) n x n test matrix with a planted spectrum
        del Bd
    del Ad
torch.cuda.synchronize()
t0 = time.perf_counter()
try:
    Xr, sig, Jt, sw = D.jacobi_svd(X, want_v=want_v, max_sweeps=sweeps)
except _lib.NotConvergedError:
    sw = sweeps
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("n=%d kind=%s sweeps=%d want_v=%s inner=%s wall %.3f s (%.3f s/sweep)" %
      (n, kind, sw, want_v, os.environ.get("XMCA_JACOBI_INNER", "default"), dt, dt / max(sw, 1)))
