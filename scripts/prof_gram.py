"""Run the tensor-core Gram kernel (fp32 field -> T x T fp64) and the fp64 DMMA product at config-2 size."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xmca_b200 import device as D

T, S = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (8192, 16384)
g = torch.Generator(device="cuda").manual_seed(0)
X = torch.randn((T, S), dtype=torch.float32, device="cuda", generator=g)
X64 = X.double()
for r in range(3):
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    G = D.gram_tc(X)
    e[1].record()
    G64 = D.matmul(X64, X64, trans_b=True, symmetric=True)
    e[2].record()
    torch.cuda.synchronize()
    fl = 2.0 * T * T * S
    print("Gram T=%d S=%d: tcgen05 %.2f ms (%.0f TFLOP/s full-square equivalent), fp64 DMMA %.2f ms (%.1f TFLOP/s computed half)"
          % (T, S, e[0].elapsed_time(e[1]), fl / e[0].elapsed_time(e[1]) / 1e9, e[1].elapsed_time(e[2]),
             fl / 2 / e[1].elapsed_time(e[2]) / 1e9), "max rel diff %.2e" % float(((G - G64).abs().max() / G64.abs().max()).item()),
          flush=True)
