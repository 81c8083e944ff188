"""Time the fused Varimax kernel per iteration (fixed iteration count, tol = 0)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xmca_b200 import device as D, _lib as L

def run(n, p, dtype, iters=200):
    rng = np.random.default_rng(0)
    k = p
    Lh = (rng.standard_normal((n, k)) @ rng.standard_normal((k, p)) * 0.3 + rng.standard_normal((n, p))).astype(dtype)
    Ld = D.to_device(Lh)
    for it in (iters,):
        for rep in range(3):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            try:
                D.varimax(Ld, 1.0, it, 0.0)
            except L.NotConvergedError:
                pass
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
        st = D.to_host(D.last_varimax_stats)
        print("   svd sweeps/iter %.2f; clocks/iter: stream %.0f sync1 %.0f reduce+sync2 %.0f matmuls %.0f jacobi %.0f rest %.0f"
              % ((st[3] / st[0],) + tuple(st[4:10] / st[0])))
        print("n=%d p=%d %s: %d iterations %.2f ms -> %.1f us/iteration, %.1f GB/s algorithmic"
              % (n, p, np.dtype(dtype).name, it, ms, ms * 1e3 / it, n * p * np.dtype(dtype).itemsize * it / ms / 1e6), flush=True)

if __name__ == "__main__":
    run(32768, 50, np.float32)
    run(131072, 50, np.float32)
    run(98304, 50, np.float64)
    run(65536, 20, np.float32)
