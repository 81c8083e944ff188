"""One config-2 hot-path step (solve + rotate(50) + getters) between cudaProfilerStart/Stop, after a warm-up step:
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv python scripts/prof_step_c2.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from xmca_b200 import MCA
T, S = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (8192, 16384)
A, B = bench.synthetic_fields(T, S, S, seed=1000)
m = MCA(A, B)
m._device_fields()
bench.hot_path_step(m, 50, 50)
torch.cuda.synchronize()
torch.cuda.profiler.start()
bench.hot_path_step(m, 50, 50)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
