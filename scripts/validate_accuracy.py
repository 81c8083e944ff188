"""Large-size accuracy check against the numpy oracle (too slow for the test suite):
    python scripts/validate_accuracy.py [T] [S] [f32|f64] [n_rot]
GPU MCA.solve()/rotate() vs the oracle on the same synthetic fields; prints the
singular-value errors (leading 50, all modes above 1e-3 sigma_1, all modes), the
principal-subspace angle of the leading well-separated modes and the rotated-variance error."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import mca_oracle as orc
from bench import synthetic_fields
from xmca_b200 import MCA

T = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
S = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
dt = np.float64 if (len(sys.argv) > 3 and sys.argv[3] == "f64") else np.float32
n_rot = int(sys.argv[4]) if len(sys.argv) > 4 else 20
A, B = synthetic_fields(T, S, S, seed=42, dtype=dt)
t0 = time.perf_counter(); m = MCA(A, B); m.solve(); tg = time.perf_counter() - t0
sv = m.singular_values().astype(np.float64)
# ground truth: the oracle on the SAME values promoted to fp64 (the reference's fp32 LAPACK path carries ~eps32*sigma_1)
t0 = time.perf_counter(); ref = orc.solve(orc.make_model(A.astype(np.float64), B.astype(np.float64))); tc = time.perf_counter() - t0
rs = ref.sigma
rel = np.abs(sv - rs) / rs
lead = rs > 1e-3 * rs[0]
print("T=%d S=%d %s route=%s sweeps=%s  gpu solve %.2f s, oracle(f64) %.2f s" % (T, S, dt.__name__, m._solve_info["route"], m._solve_info["sweeps"], tg, tc))
print("sigma rel err: top50 max %.3e | modes > 1e-3 sigma1 (%d) max %.3e | all-but-last max %.3e | abs/sigma1 max %.3e" %
      (rel[:50].max(), lead.sum(), rel[lead].max(), rel[:-1].max(), (np.abs(sv - rs) / rs[0]).max()))
if dt == np.float32:
    r32 = orc.solve(orc.make_model(A, B)).sigma.astype(np.float64)
    print("reference-style fp32 LAPACK path vs fp64 truth: top50 max %.3e | lead max %.3e" %
          ((np.abs(r32 - rs) / rs)[:50].max(), (np.abs(r32 - rs) / rs)[lead].max()))
V = m._get_V(30, rotated=False)
for k in (5, 10, 20):
    print("subspace angle of leading %d modes: left %.3e right %.3e" %
          (k, orc.subspace_angle(ref.V["left"][:, :k], V["left"][:, :k]), orc.subspace_angle(ref.V["right"][:, :k], V["right"][:, :k])))
G = V["left"][:, :30].T.astype(np.float64) @ V["left"][:, :30].astype(np.float64)
print("orthonormality of the leading 30 left vectors: max |V^T V - I| = %.3e" % np.abs(G - np.eye(30)).max())
m.rotate(n_rot, 1)
orc.rotate(ref, n_rot, 1)
print("rotated variance rel err (n_rot=%d): %.3e, varimax iterations %s" %
      (n_rot, np.max(np.abs(m.variance(n_rot) - orc.get_variance(ref, n_rot)) / orc.get_variance(ref, n_rot)), m._solve_info.get("varimax_iterations")))
