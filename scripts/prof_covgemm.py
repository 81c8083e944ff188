"""Run the tensor-core cross-covariance GEMM (C = A^T B / dof, 3xTF32 tcgen05) at config-2 size."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xmca_b200 import device as D

T, S = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (8192, 16384)
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.randn((T, S), dtype=torch.float32, device="cuda", generator=g)
B = torch.randn((T, S), dtype=torch.float32, device="cuda", generator=g)
planes = [D.split_tf32(A, transpose=True), D.split_tf32(B, transpose=True)]
C = D.empty((S, S), torch.float32)
for r in range(4):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    D.tc_gemm_nt(planes[0][0], planes[0][1], planes[1][0], planes[1][1], T, alpha=1.0 / (T - 1), out=C)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("cov-GEMM T=%d S=%d: %.2f ms, %.1f TFLOP/s algorithmic (x3 TF32 issued)" % (T, S, ms, 2.0 * T * S * S / ms / 1e9),
          flush=True)
