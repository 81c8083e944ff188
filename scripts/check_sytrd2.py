"""GPU diagnostics of the two-stage tridiagonalisation (xmca_sytrd2 / xmca_ormtr2): each stage against numpy.
Run on a B200:  python scripts/check_sytrd2.py [quick]"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xmca_b200 import _lib as L, device as D

lib = L.load()
raw = lib._raw
raw.xmca_dbg_band_chase.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p]
raw.xmca_dbg_band_chase_ll_bytes.restype = C.c_size_t
raw.xmca_dbg_band_chase_ll_bytes.argtypes = [C.c_int64]


def ll_buffer(n):
    return torch.zeros(int(raw.xmca_dbg_band_chase_ll_bytes(n)), dtype=torch.uint8, device="cuda")
B = 64


def spd(n, seed, cond=None):
    r = np.random.default_rng(seed)
    if cond is None:
        X = r.standard_normal((n, n + 7))
        return X @ X.T / n
    Q = np.linalg.qr(r.standard_normal((n, n)))[0]
    return (Q * np.logspace(0, -np.log10(cond), n)) @ Q.T


def band_of(S, b=B):
    n = S.shape[0]
    i, j = np.indices((n, n))
    return np.where(np.abs(i - j) <= b, S, 0.0)


def to_ab(Bm):
    n = Bm.shape[0]
    AB = np.zeros((n, 2 * B))
    for d in range(min(B, n - 1) + 1):
        AB[:n - d, d] = np.diag(Bm, -d)
    return AB


def from_ab(AB, n):
    Bm = np.zeros((n, n))
    for d in range(min(2 * B - 1, n - 1) + 1):
        v = AB[:n - d, d]
        Bm += np.diag(v, -d)
        if d:
            Bm += np.diag(v, d)
    return Bm


def tri_eigs(d, e):
    n = d.size
    if n == 1:
        return d.copy()
    T = np.diag(d) + np.diag(e[:n - 1], 1) + np.diag(e[:n - 1], -1)
    return np.linalg.eigvalsh(T)


def check_stage2(n, seed=0, vectors=True):
    Bm = band_of(spd(n, seed))
    ref = np.linalg.eigvalsh(Bm)
    AB = D.to_device(to_ab(Bm))
    d, e = D.empty((n,), D.f64()), D.zeros((max(n - 1, 1),), D.f64())
    V2 = D.zeros((n, n), D.f64())
    cnt = torch.zeros(n + 8, dtype=torch.int32, device="cuda")
    rc = raw.xmca_dbg_band_chase(n, L.ptr(AB), L.ptr(d), L.ptr(e), L.ptr(V2) if vectors else None, n, L.ptr(cnt),
                                 None, L.ptr(ll_buffer(n)), L.stream_ptr())
    torch.cuda.synchronize()
    dh, eh = D.to_host(d), D.to_host(e)
    lam = tri_eigs(dh, eh)
    err = np.abs(lam - ref).max() / max(np.abs(ref).max(), 1e-300)
    ABh = D.to_host(AB)
    off = np.abs(ABh[:, 2:]).max() if n > 2 else 0.0
    msg = "stage2 n=%5d rc=%d eig err %.2e residual beyond tridiagonal %.2e" % (n, rc, err, off)
    if vectors and n > 2:
        # eigenvectors of T -> Q2 z must be eigenvectors of the band matrix
        T = np.diag(dh) + np.diag(eh[:n - 1], 1) + np.diag(eh[:n - 1], -1)
        w, Zt = np.linalg.eigh(T)
        k = min(5, n)
        Z = np.ascontiguousarray(Zt[:, -k:].T)
        Zd = D.to_device(Z)
        tf = D.zeros((max(lib.xmca_sytrd2_tfac_bytes(1) // 8, 4096),), D.f64())
        # Q2 only: call the apply through ormtr2 with n small enough that there are no stage-1 panels? not in general ->
        # emulate on the host from V2 instead
        V2h = D.to_host(V2)
        X = Z.copy()
        for j in range(n - 3, -1, -1):
            k0 = 0
            while True:
                r0 = j + 1 + k0 * B
                if r0 > n - 2:
                    break
                ln = min(B, n - r0)
                v = V2h[j, r0:r0 + ln].copy()
                tau = v[0]
                v[0] = 1.0
                X[:, r0:r0 + ln] -= tau * np.outer(X[:, r0:r0 + ln] @ v, v)
                k0 += 1
        res = np.abs(Bm @ X.T - X.T * w[-k:]).max() / np.abs(w).max()
        msg += " | host-applied Q2 residual %.2e" % res
    print(msg, flush=True)
    return err


def check_full(n, seed=1, cond=None, nvec=6, lda_pad=0):
    S = spd(n, seed, cond)
    S = 0.5 * (S + S.T)
    ref = np.linalg.eigvalsh(S)
    # stage 1 only
    Sd = D.to_device(S)
    d, e = D.empty((n,), D.f64()), D.zeros((max(n - 1, 1),), D.f64())
    tfac = D.zeros((lib.xmca_sytrd2_tfac_bytes(n) // 8,), D.f64())
    wsb = lib.xmca_sytrd2_workspace_bytes(n)
    ws = torch.zeros(wsb, dtype=torch.uint8, device="cuda")
    rc = raw.xmca_sytrd2(n, L.ptr(Sd), n, L.ptr(d), L.ptr(e), L.ptr(tfac), 4, L.ptr(ws), wsb, L.stream_ptr())
    torch.cuda.synchronize()
    ABh = ws[: n * 2 * B * 8].view(torch.float64).reshape(n, 2 * B).cpu().numpy()
    Bm = from_ab(ABh, n)
    e1 = np.abs(np.linalg.eigvalsh(Bm) - ref).max() / np.abs(ref).max()
    beyond = np.abs(ABh[:, B + 1:]).max()
    # full
    Sd = D.to_device(S)
    t0 = time.time()
    try:
        d, e, tfac = D.sytrd2(Sd, want_vectors=True)
        rc2 = 0
    except Exception as ex:           # noqa: BLE001
        print("  sytrd2 raised", repr(ex))
        rc2 = -1
    torch.cuda.synchronize()
    dh, eh = D.to_host(d), D.to_host(e)
    lam = tri_eigs(dh, eh)
    e2 = np.abs(lam - ref).max() / np.abs(ref).max()
    msg = "full   n=%5d cond=%s rc=%d/%d | stage1 band eig err %.2e (beyond band %.1e) | tridiag eig err %.2e" % (
        n, cond, rc, rc2, e1, beyond, e2)
    if nvec and n > 2:
        T = np.diag(dh) + np.diag(eh[:n - 1], 1) + np.diag(eh[:n - 1], -1)
        w, Zt = np.linalg.eigh(T)
        k = min(nvec, n)
        Zd = D.to_device(np.ascontiguousarray(Zt[:, -k:].T))
        D.ormtr2(Sd, tfac, Zd)
        X = D.to_host(Zd).T
        res = np.abs(S @ X - X * w[-k:]).max() / np.abs(w).max()
        orth = np.abs(X.T @ X - np.eye(k)).max()
        msg += " | vectors: residual %.2e, orth %.2e" % (res, orth)
    print(msg, flush=True)
    return e2


def timing(n, reps=3):
    S = spd(n, 3)
    S = 0.5 * (S + S.T)
    Sd0 = D.to_device(S)
    wsb = lib.xmca_sytrd2_workspace_bytes(n)
    ws = torch.zeros(wsb, dtype=torch.uint8, device="cuda")
    d, e = D.empty((n,), D.f64()), D.zeros((n,), D.f64())
    tfac = D.zeros((lib.xmca_sytrd2_tfac_bytes(n) // 8,), D.f64())
    out = {}
    for label, flag in (("stage1", 4), ("values", 0), ("with-reflectors", 1)):
        best = 1e9
        for _ in range(reps):
            Sd = Sd0.clone()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            raw.xmca_sytrd2(n, L.ptr(Sd), n, L.ptr(d), L.ptr(e), L.ptr(tfac), flag, L.ptr(ws), wsb, L.stream_ptr())
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out[label] = best
    # old one-stage for comparison
    Sd = Sd0.clone()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    D.sytrd(Sd)
    e1.record()
    torch.cuda.synchronize()
    out["one-stage"] = e0.elapsed_time(e1)
    # ormtr2 for 50 vectors
    Sd = Sd0.clone()
    dd, ee, tf = D.sytrd2(Sd)
    Z = D.to_device(np.random.default_rng(0).standard_normal((50, n)))
    torch.cuda.synchronize()
    e0.record()
    D.ormtr2(Sd, tf, Z)
    e1.record()
    torch.cuda.synchronize()
    out["ormtr2(50)"] = e0.elapsed_time(e1)
    lam = D.to_host(D.stebz(dd, ee))
    ref = np.linalg.eigvalsh(S)[::-1]
    out["eig err"] = float(np.abs(lam - ref).max() / ref.max())
    print("timing n=%d: %s" % (n, {k: (round(v, 3) if v > 1e-3 else v) for k, v in out.items()}), flush=True)


def chase_profile(n):
    """phase clocks of the bulge-chasing kernel (thread 0 of every CTA, summed over tasks)"""
    Bm = band_of(spd(n, 5))
    AB0 = D.to_device(to_ab(Bm))
    for vec in (False, True):
        AB = AB0.clone()
        d, e = D.empty((n,), D.f64()), D.zeros((n,), D.f64())
        V2 = D.zeros((n, n), D.f64()) if vec else None
        cnt = torch.zeros(n + 8, dtype=torch.int32, device="cuda")
        prof = torch.zeros(8, dtype=torch.int64, device="cuda")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        llb = ll_buffer(n)
        torch.cuda.synchronize()
        e0.record()
        raw.xmca_dbg_band_chase(n, L.ptr(AB), L.ptr(d), L.ptr(e), L.ptr(V2), n, L.ptr(cnt), L.ptr(prof), L.ptr(llb), L.stream_ptr())
        e1.record()
        torch.cuda.synchronize()
        p = prof.cpu().numpy()
        nt = max(int(p[6]), 1)
        names = ["wait", "load", "matvec", "update+larfg", "store", "publish"]
        print("chase n=%d vectors=%s: %.2f ms, %d tasks; clocks/task: %s; LL mismatches %d" % (
            n, vec, e0.elapsed_time(e1), nt, ", ".join("%s %.0f" % (a, p[i] / nt) for i, a in enumerate(names)), int(p[7])),
            flush=True)


if __name__ == "__main__":
    if "prof" in sys.argv:
        for nn in ([int(a) for a in sys.argv if a.isdigit()] or [8192]):
            chase_profile(nn)
        sys.exit(0)
    quick = "quick" in sys.argv
    if "tiny" in sys.argv:             # for compute-sanitizer
        for n in (3, 66, 130):
            check_stage2(n)
        for n in (70, 130, 200):
            check_full(n)
        sys.exit(0)
    for n in (2, 3, 5, 40, 64, 65, 66, 129, 200, 1000):
        check_stage2(n)
    for n in (66, 70, 127, 128, 129, 130, 191, 257, 700, 1000, 2050):
        check_full(n)
    check_full(1024, cond=1e12)
    check_full(1500, cond=1e6)
    if not quick:
        timing(2048, reps=2)
        timing(4096, reps=2)
        timing(8192, reps=3)

