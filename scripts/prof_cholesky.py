"""Time xmca_cholesky (fp64, n x n SPD) and check it against numpy: python scripts/prof_cholesky.py [n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xmca_b200 import device as D
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
g = torch.Generator(device="cuda").manual_seed(0)
X = torch.randn((n, n + 64), device="cuda", dtype=torch.float64, generator=g)
G = (X @ X.T) / n
for rep in range(3):
    Gc = G.clone()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    Ld, inv = D.cholesky(Gc)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
Lt = torch.tril(Ld)
err = float((Lt @ Lt.T - G).abs().max() / G.abs().max())
print("cholesky n=%d: %.2f ms (%.1f TFLOP/s), |L L^T - G| / |G| = %.2e, mode=%s" % (n, ms, n ** 3 / 3 / ms / 1e9, err, os.environ.get("XMCA_CHOL_DIAG", "blocked")))
