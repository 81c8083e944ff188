"""Where the device idles inside one config-2 solve + rotate step: gaps between consecutive C-ABI calls (CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from xmca_b200 import MCA, _lib
T, S = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (8192, 16384)
A, B = bench.synthetic_fields(T, S, S, seed=1000)
m = MCA(A, B)
m._device_fields()
for _ in range(2):
    bench.hot_path_step(m, 50, 50)
torch.cuda.synchronize()
e_begin, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
_lib.profile_begin()
e_begin.record()
bench.hot_path_step(m, 50, 50)
e_end.record()
rec = list(_lib._profile)
_lib.profile_end()
torch.cuda.synchronize()
tot = e_begin.elapsed_time(e_end)
busy = sum(e0.elapsed_time(e1) for _, e0, e1, _ in rec)
print("step %.1f ms, inside C-ABI calls %.1f ms, outside %.1f ms (%d calls)" % (tot, busy, tot - busy, len(rec)))
gaps = [("<start> -> " + rec[0][0], e_begin.elapsed_time(rec[0][1]))]
for (n0, _, a1, _), (n1, b0, _, _) in zip(rec[:-1], rec[1:]):
    gaps.append((n0 + " -> " + n1, a1.elapsed_time(b0)))
gaps.append((rec[-1][0] + " -> <end>", rec[-1][2].elapsed_time(e_end)))
for name, g in sorted(gaps, key=lambda x: -x[1])[:16]:
    print("  %7.2f ms  %s" % (g, name))
print("ordered calls (ms in call / gap before it):")
prev = e_begin
for name, a0, a1, nl in rec:
    print("   %-22s %8.2f   gap %6.2f" % (name, a0.elapsed_time(a1), prev.elapsed_time(a0)))
    prev = a1

# host-side view of the tail: wall time of each public call of the step, then a cProfile of the accessors
import time, cProfile, pstats
def wall(label, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize()
    print("  wall %-22s %7.2f ms" % (label, (time.perf_counter() - t0) * 1e3)); return out
wall("solve", lambda: m.solve())
wall("rotate(50)", lambda: m.rotate(50, 1))
wall("singular_values(50)", lambda: m.singular_values(50))
wall("pcs(50)", lambda: m.pcs(50))
wall("eofs(50)", lambda: m.eofs(50))
pr = cProfile.Profile()
m.solve()
pr.enable(); m.rotate(50, 1); m.singular_values(50); m.pcs(50); m.eofs(50); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)

# per-call device time of three more steps
for rep in range(3):
    _lib.profile_begin()
    bench.hot_path_step(m, 50, 50)
    pr2 = _lib.profile_end()
    print("  step calls:", {k: round(v["ms"], 1) for k, v in pr2.items() if v["ms"] > 0.5})
