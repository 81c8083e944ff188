"""Design prototype (numpy, NOT shipped): which progress rule lets the bulge-chasing sweeps of stage 2
(scripts/proto/two_stage_sytrd.py) run concurrently?  Task (j, k) = k-th chase step of sweep j.  The planned kernel
gives every sweep to one CTA and lets task (j, k) start once sweep j - 1 has finished task k + LAG.  This script
executes the tasks in RANDOM order among those that are ready under the rule and compares the result bit for bit
with the sequential order: the smallest LAG that reproduces it is the rule the kernel has to implement.

Run:  python scripts/proto/bulge_chase_schedule.py [n] [b]
"""
import sys
import numpy as np
from two_stage_sytrd import house, stage1_dense_to_band


def n_tasks(n, b, j):
    """number of chase steps of sweep j (static: rows r0 = j + 1 + k b < n - 1)"""
    k = 0
    while j + 1 + k * b < n - 1:
        k += 1
    return k


def run_task(B, n, b, j, k):
    col = j if k == 0 else j + 1 + (k - 1) * b
    r0 = j + 1 + k * b
    r1 = min(r0 + b, n)
    v, tau, beta = house(B[r0:r1, col].copy())
    lo, hi = max(col, r0 - b), min(n, r1 + b)
    Hw = B[r0:r1, lo:hi]
    Hw -= tau * np.outer(v, v @ Hw)
    Hc = B[lo:hi, r0:r1]
    Hc -= tau * np.outer(Hc @ v, v)


def chase(Bm, b, lag=None, seed=0):
    """lag None: sequential.  Otherwise random ready-order under the rule (j, k) after (j, k-1) and
    (j-1, min(k + lag, last))."""
    B = Bm.copy(); n = B.shape[0]
    nt = [n_tasks(n, b, j) for j in range(n - 2)]
    if lag is None:
        for j in range(n - 2):
            for k in range(nt[j]):
                run_task(B, n, b, j, k)
        return B
    rng = np.random.default_rng(seed)
    done = [0] * (n - 2)                      # tasks finished per sweep
    active = 0
    while True:
        ready = []
        for j in range(n - 2):
            k = done[j]
            if k >= nt[j]:
                continue
            if j == 0 or done[j - 1] >= min(k + lag + 1, nt[j - 1]):
                ready.append(j)
            if j > 0 and done[j - 1] == 0:
                break                          # later sweeps cannot be ready either
        if not ready:
            break
        j = ready[rng.integers(len(ready))]
        run_task(B, n, b, j, done[j])
        done[j] += 1
        active = max(active, len(ready))
    return B, active


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    b = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    rng = np.random.default_rng(1)
    S = rng.standard_normal((n, n)); S = S @ S.T / n
    Bm, _ = stage1_dense_to_band(S, b)
    seq = chase(Bm, b)
    for lag in (0, 1, 2, 3):
        worst, width = 0.0, 0
        for seed in range(5):
            par, act = chase(Bm, b, lag, seed)
            worst = max(worst, np.abs(par - seq).max()); width = max(width, act)
        print("LAG %d: max |random-order - sequential| = %.2e over 5 random schedules, up to %d sweeps ready at once"
              % (lag, worst, width))
