"""One Varimax call (fixed 60 iterations) for ncu: python scripts/prof_varimax_one.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xmca_b200 import device as D, _lib as L
rng = np.random.default_rng(0)
n, p = 32768, 50
Lh = (rng.standard_normal((n, p)) @ rng.standard_normal((p, p)) * 0.3 + rng.standard_normal((n, p))).astype(np.float32)
Ld = D.to_device(Lh)
try:
    D.varimax(Ld, 1.0, 60, 0.0)
except L.NotConvergedError:
    pass
torch.cuda.synchronize()
