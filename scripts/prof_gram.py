"""Time the tensor-core Gram matrix G = X X^T (3xTF32 tcgen05, fp64 chunk sums): python scripts/prof_gram.py [T S]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xmca_b200 import device as D
T, S = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (8192, 16384)
g = torch.Generator(device="cuda").manual_seed(0)
X = torch.randn((T, S), device="cuda", dtype=torch.float32, generator=g)
for rep in range(4):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    G = D.gram_tc(X, 1.0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
ref = X[:512].double() @ X[:512].double().T
err = float((G[:512, :512] - ref).abs().max() / ref.abs().max())
print("gram_tc T=%d S=%d: %.2f ms (incl. the operand split), %.1f TFLOP/s algorithmic (symmetric half), max rel err of a corner %.2e"
      % (T, S, ms, T * T * S / ms / 1e9, err))
