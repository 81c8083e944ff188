"""Varimax statistics on the config-2 workload (solve + rotate(50) of the bench's synthetic model)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from xmca_b200 import MCA, device as D
T, S = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (8192, 16384)
A, B = bench.synthetic_fields(T, S, S, seed=1000)
m = MCA(A, B)
m.solve()
for rep in range(2):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    m.rotate(50, 1)
    e1.record()
    torch.cuda.synchronize()
    st = D.to_host(D.last_varimax_stats)
    it = st[0]
    print("rotate(50): %.1f ms; iterations %d, sweeps %d (%.2f/iter), rotations %d (%.1f per sweep), pairs > 1e-4: %d, > 1e-3: %d (per guard pass: %.1f / %.1f)"
          % (e0.elapsed_time(e1), it, st[3], st[3] / it, st[10], st[10] / max(st[3], 1), st[11], st[12], st[11] / (it + st[3]), st[12] / (it + st[3])))
    print("   clocks/iter: stream %.0f sync1 %.0f reduce+sync2 %.0f matmuls %.0f jacobi %.0f rest %.0f" % tuple(st[4:10] / it), flush=True)
