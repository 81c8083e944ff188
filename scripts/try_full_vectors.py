"""Experiment: all `rank` singular vectors through the tridiagonal route (stein for every eigenvalue + ormtr2), with
per-phase times.  python scripts/try_full_vectors.py T S"""
import sys, time, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synthetic_fields
from xmca_b200 import MCA, engine as E, device as D
T, S = int(sys.argv[1]), int(sys.argv[2])
A, B = synthetic_fields(T, S, S, seed=3)
m = MCA(A, B); m.solve()
res = m._dV[1]
lam = res.lam[:T]
tnorm = float(res.lam[0])
gap = 2.2e-16 * tnorm / 1e-7
starts = [0] + [i for i in range(1, T) if lam[i - 1] - lam[i] > gap] + [T]
sizes = np.diff(starts)
print("clusters %d, largest %d, members in clusters > 1: %d" % (len(sizes), sizes.max(), sizes[sizes > 1].sum()), flush=True)
def timed(fn, name):
    torch.cuda.synchronize(); t0 = time.time(); out = fn(); torch.cuda.synchronize()
    print("%s: %.3f s" % (name, time.time() - t0), flush=True); return out
mm = int(sys.argv[3]) if len(sys.argv) > 3 else T
Z = timed(lambda: D.stein(res.d, res.e, lam[:mm], np.asarray([s for s in starts if s < mm] + [mm]), tnorm, iterations=2), "stein %d" % mm)
Z = timed(lambda: D.ormtr2(res.Q, res.tfac, Z), "ormtr2 %d" % mm)
