"""cov-GEMM C = A^T B / dof (tcgen05 3xTF32) at the north_star shape T = 16384, S1 = S2 = 65536, fp32.
python scripts/prof_covgemm_northstar.py [T S]   (prints algorithmic TFLOP/s; run under ncu for the tensor-pipe %)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xmca_b200 import device as D
T = int(sys.argv[1]) if len(sys.argv) > 2 else 16384
S = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.randn((T, S), device="cuda", dtype=torch.float32, generator=g)
B = torch.randn((T, S), device="cuda", dtype=torch.float32, generator=g)
pa = D.split_tf32(A, transpose=True)
pb = D.split_tf32(B, transpose=True)
del A, B
C = D.empty((S, S), torch.float32)
flops = 2.0 * T * S * S
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    D.tc_gemm_nt(pa[0], pa[1], pb[0], pb[1], T, alpha=1.0 / (T - 1), out=C)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("cov-GEMM T=%d S=%d: %.1f ms = %.1f TFLOP/s algorithmic (x3 TF32 MMAs issued = %.1f TFLOP/s)" % (T, S, ms, flops / ms / 1e9, 3 * flops / ms / 1e9), flush=True)
# spot check against fp64 on a corner
idx = torch.arange(0, 256, device="cuda")
ref = (pa[0][:256, :T].double() + pa[1][:256, :T].double()) @ (pb[0][:256, :T].double() + pb[1][:256, :T].double()).T / (T - 1)
print("max abs err of a 256 x 256 corner vs fp64: %.2e (|C| ~ %.2e)" % (float((C[:256, :256].double() - ref).abs().max()), float(ref.abs().max())))
