"""Time the tridiagonal eigen-solver kernels alone (for ncu captures and phase traces)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xmca_b200 import device as D

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
g = torch.Generator(device="cuda").manual_seed(0)
X = torch.randn((n, 2 * n), dtype=torch.float64, device="cuda", generator=g)
S0 = D.matmul(X, X, trans_b=True, symmetric=True)
for r in range(reps):
    S = S0.clone()
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    d, e, tau = D.sytrd(S)
    e1.record()
    w = D.stebz(d, e)
    e2.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1)
    print("n=%d: sytrd %.1f ms (%.0f GB/s on n^3*8/3 algorithmic bytes), stebz %.1f ms" %
          (n, t, n ** 3 * 8 / 3 / t / 1e6, e1.elapsed_time(e2)), flush=True)
if os.environ.get("XMCA_PROF_CHECK"):
    ref = torch.linalg.eigvalsh(S0).flip(0)
    print("variant %s: max |lambda - eigvalsh| / lambda_max = %.3e" %
          (os.environ.get("XMCA_SYTRD_VARIANT", "default"), float((w - ref).abs().max() / ref.abs().max())), flush=True)
if os.environ.get("XMCA_PROF_PAIR"):
    # two problems per batched call against two single calls
    for r in range(reps):
        Sp = torch.stack([S0, S0.flip(0).flip(1).contiguous()]).contiguous()
        torch.cuda.synchronize()
        e0, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(2))
        e0.record()
        d2, e2_, tau2 = D.sytrd_pair(Sp)
        e1.record()
        torch.cuda.synchronize()
        t2 = e0.elapsed_time(e1)
        w2 = D.stebz(d2[0], e2_[0, :n - 1])
        print("n=%d: batched pair %.1f ms = %.1f ms per problem (single: %.1f ms); max |lambda_pair - lambda_single| / lambda_max = %.2e"
              % (n, t2, t2 / 2, t, float((w2 - w).abs().max() / w.abs().max())), flush=True)
