"""One two-stage tridiagonalisation at n = 8192 (for `ncu --metrics gpu__time_duration.sum`): python scripts/prof_sytrd2.py [n]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xmca_b200 import device as D

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
r = np.random.default_rng(0)
X = r.standard_normal((n, n + 3))
S = D.to_device(X @ X.T / n)
S2 = S.clone()
d, e, tf = D.sytrd2(S)           # warm-up
torch.cuda.synchronize()
d, e, tf = D.sytrd2(S2)
Z = D.to_device(r.standard_normal((50, n)))
D.ormtr2(S2, tf, Z)
torch.cuda.synchronize()
print("done", float(d[0]))
