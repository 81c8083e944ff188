#!/usr/bin/env python
"""Benchmark of the MCA hot path (BASELINE.json metric) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl product|reference]
                    [--workload c2|half|small] [--rule-n-runs R]

One *step* = one pass of the hot path over one model: ``solve()`` +
``rotate(n_rot=50, power=1)`` + ``singular_values(50)``/``pcs(50)``/``eofs(50)``
on the config-2 workload (two synthetic fp32 fields T=8192, S1=S2=16384,
low-rank + noise generator of SURVEY.md 8d).  With N > 1 ranks every rank runs
its own model of that shape (weak scaling, independent replicas: a single
solve/rotate does not shard) and -- the part of the path that
does shard -- ``rule_n`` surrogates are block-partitioned over the ranks with
one all-gather at the end.

Printed JSON line (rank 0):
  value            models/s, whole job, fields resident in HBM before the timed region
  ms_per_step      wall milliseconds of one solve()+rotate()+getters
  e2e              the same through ``MCA(host arrays)``: pinned host fields ->
                   ctor -> upload -> solve -> rotate -> getters -> host results
  solve_rotate_wall_s, cov_gemm, rule_n : the three numbers BASELINE.json names
  roofline         dominant C-ABI call class of the step (CUDA events per call)
  cpu_baseline     the reference's CPU path on this box's host cores, bounded sample (half / quarter size,
                   extrapolation labelled as such)
  config.rule_n    strong-scaling rule_n: a FIXED total of surrogates (--rule-n-total) split over the ranks
``--impl reference`` times the LIVE reference (oracle/_ref: the unmodified xmca package, copied there by
``__graft_entry__.build()``; kind "reference") -- or the numpy oracle port when that copy is absent (kind "port")
-- with all host cores: one half-size step, then ONE FULL-SIZE step of the workload (measured, not extrapolated)
whenever the half-size time predicts that it fits the time budget (XMCA_REF_BUDGET_S, default 480 s).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (T, S1, S2, n_rot, n_modes)   -- extra settings in WORKLOAD_OPTS
    "c2": (8192, 16384, 16384, 50, 50),
    "half": (4096, 8192, 8192, 50, 50),
    "small": (1024, 2048, 2048, 20, 20),
    "c3": (8192, 32768, 32768, 20, 20),        # complex MCA (Hilbert) + Varimax n_rot = 20
    "c3half": (4096, 16384, 16384, 20, 20),
    "c5": (16384, 65536, 32768, 50, 50),       # fp64, Promax power 4
    "c5half": (8192, 32768, 16384, 50, 50),
}
WORKLOAD_OPTS = {
    "c3": {"complexify": True}, "c3half": {"complexify": True},
    "c5": {"dtype": np.float64, "power": 4}, "c5half": {"dtype": np.float64, "power": 4},
}


def workload_string(workload):
    T, S1, S2, n_rot, n_modes = WORKLOADS[workload]
    o = opts(workload)
    return "%s: %sMCA T=%d S1=%d S2=%d %s, rotate(n_rot=%d, power=%d), getters n=%d" % (
        workload, "complex " if o["complexify"] else "", T, S1, S2, np.dtype(o["dtype"]).name, n_rot, o["power"], n_modes)


# one `ncu --set full` capture of the two DMMA kernels of xmca_sytrd2 (profiles/r2_ncu_summary.md), per launch
SYTRD2_NCU = {"traffic": 323.1e6, "traffic_algorithmic_same_launch": 315.0e6,
              "traffic_note": "ONE launch of sbr_symm_kernel (ncu --set full, n = 8192, panel ~29, m = 6272, 184 us): 312 MB read + "
                              "11 MB written for m^2 * 8 = 315 MB of algorithmic bytes, 79 % DMMA-pipe active (27.5 TFLOP/s); "
                              "sbr_syr2k_kernel: 412 MB for 453 MB algorithmic, 63 % (profiles/r2_ncu_summary.md); `achieved` "
                              "averages the WHOLE call (panel factorisations and the latency-bound bulge chasing included)",
              "kernels_ncu": {"sbr_symm_kernel": {"tflops": 27.5, "dmma_pipe_active_pct": 79.3},
                              "sbr_syr2k_kernel": {"tflops": 21.5, "dmma_pipe_active_pct": 63.5},
                              "sb_chase_kernel": {"ms": 68.0, "bound": "latency (L2 round trips; 8 MB band, L2 resident)"}}}

# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of sytrd_panel_kernel<true> (ncu --set full, n = 8192,
# panel 9 of 128) next to the algorithmic bytes of that launch (64 columns x n'^2 x 4 B)
SYTRD_NCU_TRAFFIC = {"traffic": 15.36e9, "traffic_algorithmic_same_launch": 14.95e9,
                     "traffic_note": "ONE panel launch (ncu --set full, n = 8192, panel 9/128, 3.84 ms): 15.21 GB read + 0.15 GB "
                                     "written for 14.95 GB of algorithmic bytes = 4.0 TB/s; `achieved` averages all panel "
                                     "launches, barriers and trailing updates of a call (profiles/r1_ncu_summary.md)"}


def opts(workload):
    o = {"complexify": False, "dtype": np.float32, "power": 1}
    o.update(WORKLOAD_OPTS.get(workload, {}))
    return o


def synthetic_fields(T, S1, S2, seed, k=64, dtype=np.float32):
    """Low-rank-plus-noise fields: k shared time series with
    geometrically decaying amplitudes + unit white noise."""
    rng = np.random.default_rng(seed)
    ts = rng.standard_normal((T, k), dtype=np.float32)
    amp = (3.0 * np.sqrt(max(S1, S2)) * 0.9 ** np.arange(k) / np.sqrt(k)).astype(np.float32)
    out = []
    for S in (S1, S2):
        pat = rng.standard_normal((k, S), dtype=np.float32) / np.float32(np.sqrt(S))
        X = rng.standard_normal((T, S), dtype=np.float32)
        X += (ts * amp) @ pat
        out.append(X.astype(dtype, copy=False))
    return out


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d["bf16_tflops"]),
                "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
            "source": "fallback (B200_PROFILING.md)"}


def measure_gemm_peaks():
    """cuBLAS GEMM throughput of THIS box for the two pipes MEASURED_PEAKS.json does not cover: fp64 (DGEMM: the DMMA
    pipe the two-stage tridiagonalisation and the fp64 Gram matrices run on) and TF32 (the pipe of the 3xTF32 cov-GEMM).
    torch.matmul 8192^3, best of 5, CUDA events -- library calls used as roofline denominators only."""
    import torch
    out = {}
    n = 8192
    for name, dt, tf32 in (("fp64_tflops", torch.float64, False), ("tf32_tflops", torch.float32, True)):
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        try:
            a = torch.randn((n, n), dtype=dt, device="cuda")
            b = torch.randn((n, n), dtype=dt, device="cuda")
            best = 1e30
            for i in range(6):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                c = a @ b
                e1.record()
                torch.cuda.synchronize()
                if i:
                    best = min(best, e0.elapsed_time(e1))
            out[name] = 2.0 * n ** 3 / best / 1e9
            del a, b, c
        finally:
            torch.backends.cuda.matmul.allow_tf32 = old
    torch.cuda.empty_cache()
    out["how"] = "torch.matmul (cuBLAS) %d^3, best of 5, CUDA events, measured in this run" % n
    return out


class ClockSampler:
    """nvidia-smi sampled every 200 ms during the timed region (profiling recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, "/tmp/xmca_clocks_%d.csv" % os.getpid()

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as fh:
            for line in fh:
                parts = [x.strip() for x in line.split(",")]
                if len(parts) < 7:
                    continue
                try:
                    sm.append(float(parts[0])); mx.append(float(parts[1])); power.append(float(parts[2]))
                except ValueError:
                    continue
                for nm, val in zip(names, parts[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": float(max(power))}


# ------------------------------------------------------------------ product arm
def hot_path_step(m, n_rot, n_modes, complexify=False, power=1):
    m.solve(complexify=complexify)
    m.rotate(n_rot, power)
    sv = m.singular_values(n_modes)
    pcs = m.pcs(n_modes)
    eofs = m.eofs(n_modes)
    return sv, pcs, eofs


def result_bytes(res):
    sv, pcs, eofs = res
    return int(sv.nbytes + sum(v.nbytes for v in pcs.values()) + sum(v.nbytes for v in eofs.values()))


def run_product(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from xmca_b200 import MCA, _lib
    from xmca_b200 import device as D

    torch.cuda.set_device(local_rank)
    _lib.load()
    T, S1, S2, n_rot, n_modes = WORKLOADS[args.workload]
    wo = opts(args.workload)
    peaks = load_peaks()
    peaks.update(measure_gemm_peaks())

    # pinned host fields (the e2e leg copies from these every step)
    A0, B0 = synthetic_fields(T, S1, S2, seed=1000 + rank, dtype=wo["dtype"])
    Ap = torch.from_numpy(A0).pin_memory()
    Bp = torch.from_numpy(B0).pin_memory()
    A, B = Ap.numpy(), Bp.numpy()
    del A0, B0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = _lib.launch_count()
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = reduce_max(e0.elapsed_time(e1))
        return ms / steps, _lib.launch_count() - n0, out

    # ---- leg 1: device-resident (value) -------------------------------------------------
    model = MCA(A, B)
    model._device_fields()                     # upload + keep resident
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    step = lambda mm: hot_path_step(mm, n_rot, n_modes, wo["complexify"], wo["power"])
    ms_step, launches, res = timed(lambda: step(model), args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    info = dict(model._solve_info)

    # ---- per-call-class device time of one more step (CUDA events on the launch stream) -
    _lib.profile_begin()
    step(model)
    prof = _lib.profile_end()
    vm_iters = int(model._solve_info.get("varimax_iterations", 0))

    # ---- leg 2: end to end through the public class, host buffers ----------------------
    def e2e_step():
        m = MCA(A, B)
        return step(m)
    ms_e2e, _, res_e2e = timed(e2e_step, args.steps, min(args.warmup, 1))
    h2d = int(A.nbytes + B.nbytes)
    d2h = result_bytes(res_e2e)

    # ---- cov-GEMM on the tensor cores (C = A^T B / dof, 3xTF32 tcgen05) -----------------
    dA, dB = model._dev["left"], model._dev["right"]
    cov = None
    if dA.dtype == torch.float32 and not wo["complexify"]:
        planes = [D.split_tf32(dA, transpose=True), D.split_tf32(dB, transpose=True)]
        Cbuf = D.empty((S1, S2), torch.float32)

        def gemm_only():
            D.tc_gemm_nt(planes[0][0], planes[0][1], planes[1][0], planes[1][1], T, alpha=1.0 / (T - 1), out=Cbuf)
        ms_gemm, _, _ = timed(gemm_only, 5, 3)
        ms_cov, _, _ = timed(lambda: D.cov_gemm_tc(dA, dB, 1.0 / (T - 1)), 3, 1)
        flops = 2.0 * T * S1 * S2
        cov = {"tflops_gemm_kernel": flops / ms_gemm / 1e9, "tflops_with_operand_split": flops / ms_cov / 1e9,
               "ms_gemm_kernel": ms_gemm, "ms_with_operand_split": ms_cov, "flops": flops,
               "note": "algorithmic 2*T*S1*S2 flops; the kernel issues 3 TF32 products per flop pair (3xTF32)"}
        del planes, Cbuf
        torch.cuda.empty_cache()

    # ---- rule_n: STRONG scaling -- a fixed total of surrogates block-partitioned over the ranks, one all-gather ------
    # Default surrogates are float64 like the reference's (array.py:1756); the fp32 fast mode (Gram matrices on the
    # tensor cores) is timed next to it on a quarter of the runs.
    rn = None
    if args.rule_n_total > 0:
        model.solve(complexify=wo["complexify"])      # rule N of the unrotated model
        n_runs = args.rule_n_total

        def timed_rule_n(total, **kw):
            model.rule_n(2 * world, n_modes, seed=99, **kw)      # warm-up (allocator, kernel attributes)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            sp = model.rule_n(total, n_modes, seed=1234, **kw)
            e1.record()
            barrier()
            ms = reduce_max(e0.elapsed_time(e1))
            return {"surrogates_per_s": total / (ms / 1e3), "n_runs": total, "ms_total": ms, "shape": list(sp.shape)}

        rn = timed_rule_n(n_runs)
        rn.update({"surrogate_dtype": "float64 (reference default, array.py:1756)", "scaling": "strong",
                   "runs_per_rank": n_runs / world, "dtype": "f64",
                   "collective": "1 all_gather (nccl)" if world > 1 else "none (1 rank)"})
        if np.dtype(wo["dtype"]) == np.float32:
            fast = timed_rule_n(max(n_runs // 4, 2 * world), surrogate_dtype="float32")
            fast["surrogate_dtype"] = "float32 (opt-in fast mode: Gram matrices on the tensor cores)"
            rn["float32_surrogates"] = fast
        _lib.profile_begin()
        model.rule_n(world, n_modes, seed=4321)
        rn_prof = _lib.profile_end()
        rn["call_ms_one_surrogate"] = {k: round(v["ms"], 2) for k, v in sorted(rn_prof.items(), key=lambda kv: -kv[1]["ms"])}

    if rank != 0:
        return None

    # ---- roofline of the dominant call class -------------------------------------------
    total_prof = sum(v["ms"] for v in prof.values()) or 1.0
    shares = {k: {"ms": round(v["ms"], 3), "share": round(v["ms"] / total_prof, 4), "calls": v["calls"],
                  "launches": v["launches"]} for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
    dom = next(iter(shares))
    n_load = S1 + S2
    e_store = np.dtype(wo["dtype"]).itemsize * (2 if wo["complexify"] else 1)
    roof_list = {}
    if "xmca_varimax" in prof:
        v = prof["xmca_varimax"]
        vbytes = (vm_iters + 5) * n_load * n_rot * e_store
        roof_list["xmca_varimax"] = {"bound": "hbm", "achieved": vbytes / (v["ms"] / v["calls"]) / 1e6,
                                     "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                     "bytes_per_launch": vbytes, "iterations": vm_iters}
    if "xmca_jacobi_svd" in prof:
        v = prof["xmca_jacobi_svd"]
        sw = info.get("sweeps", [])
        n = T
        # blocked sweep: panel Gram 4 m n^2 + panel update 4 m n^2 (+ 4 n^3 when rotations are accumulated)
        per_sweep = 8.0 * n ** 3 if info.get("route") == "cholqr" else 12.0 * n ** 3
        fl = sum(s * per_sweep for s in sw)
        roof_list["xmca_jacobi_svd"] = {"bound": "fp64", "achieved": fl / v["ms"] / 1e9, "peak": 37.0,
                                        "unit": "TFLOP/s", "sweeps": sw,
                                        "note": "fp64 CUDA-core path (tcgen05 has no fp64 kind); peak = 37 TFLOP/s DFMA "
                                                "measured with scripts/ubench_fp64.cu (not in MEASURED_PEAKS.json)"}
    if "xmca_sytrd" in prof:
        v = prof["xmca_sytrd"]
        n_eig = min(T, S1, S2)
        # every element of the symmetric trailing matrix once per Householder column (y = A v): n^3 * 8 / 6 bytes;
        # the tile-major pass reads exactly that while the trailing matrix exceeds the L2, the L2-resident tail
        # (n' <= 4096) reads full rows from the cache
        sbytes = n_eig ** 3 * 8.0 / 6.0
        roof_list["xmca_sytrd"] = {"bound": "hbm", "achieved": sbytes * v["calls"] / v["ms"] / 1e6,
                                   "peak": peaks["hbm_gbs"], "unit": "GB/s", "bytes_per_call": sbytes, "n": n_eig,
                                   "kernel": "sytrd_panel_kernel<tiled> (+ tiled rank-128 trailing update)",
                                   "note": "algorithmic bytes n^3*8/6 = one triangle of the trailing matrix per Householder "
                                           "column; the time is the whole xmca_sytrd call (panel kernels, grid barriers, "
                                           "trailing updates, layout conversions)"}
    if "xmca_sytrd2" in prof:
        v = prof["xmca_sytrd2"]
        n_eig = v_n = int(info.get("eigen_n") or min(T, S1, S2))
        fl = 4.0 / 3.0 * float(v_n) ** 3
        roof_list["xmca_sytrd2"] = {
            "bound": "tensor", "achieved": fl * v["calls"] / v["ms"] / 1e9, "peak": peaks["fp64_tflops"],
            "unit": "TFLOP/s", "flops_per_call": fl, "n": n_eig,
            "kernel": "xmca_sytrd2: sbr_symm_kernel + sbr_syr2k_kernel (fp64 DMMA, stage 1) + sb_chase_kernel (stage 2)",
            "note": "algorithmic 4/3 n^3 flops of the dense -> band reduction over the time of the WHOLE call (stage 1 "
                    "GEMMs and panel factorisations, bulge chasing); peak = cuBLAS DGEMM measured in this run "
                    "(MEASURED_PEAKS.json has no fp64 entry)"}
        roof_list["xmca_sytrd2"].update(SYTRD2_NCU)
    for name in ("xmca_gemm_ex", "xmca_gemm"):
        if name in prof:
            v = prof[name]
            roof_list[name] = {"bound": "fp64-simt", "ms": v["ms"], "calls": v["calls"]}
    if cov:
        roof_list["xmca_tc_gemm_nt"] = {"bound": "tensor", "achieved": cov["tflops_gemm_kernel"] * 3,
                                        "peak": peaks["tf32_tflops"], "unit": "TFLOP/s (TF32 issued)",
                                        "note": "3 TF32 MMAs per algorithmic product; peak = cuBLAS TF32 GEMM measured "
                                                "in this run"}
    for r in roof_list.values():
        if "achieved" in r and "peak" in r:
            r["frac"] = r["achieved"] / r["peak"]
    roof = dict(roof_list.get(dom, {"bound": "hbm", "achieved": None, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                    "frac": None}))
    roof.update({"kernel": roof.get("kernel", dom), "call": dom, "share_of_step": shares[dom]["share"],
                 "peak_source": peaks["source"]})
    if dom == "xmca_sytrd":
        # one `ncu --set full` capture of the tiled panel kernel (profiles/r1_ncu_summary.md)
        roof.update(SYTRD_NCU_TRAFFIC)
    else:
        roof.setdefault("traffic", None)
    peaks_out = {k: peaks[k] for k in ("hbm_gbs", "bf16_tflops", "fp64_tflops", "tf32_tflops", "source", "how") if k in peaks}

    out = {
        "metric": "solve()+rotate() throughput (models/s; wall-sec in solve_rotate_wall_s, cov-GEMM TFLOP/s in "
                  "cov_gemm, rule_n surrogates/s in rule_n)",
        "value": world / (ms_step / 1e3), "unit": "models/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "%s fields; f64 accumulation/eigen-solver/rotation" % np.dtype(wo["dtype"]).name, "data": "synthetic",
        "config": {"workload": workload_string(args.workload),
                   "l2": "inputs (2 x %.0f MB) larger than the 126 MB L2" % (A.nbytes / 1e6),
                   "parallelism": "replicas x%d (solve/rotate); rule_n surrogates block-sharded" % world,
                   "route": info.get("route"), "jacobi_sweeps": info.get("sweeps"),
                   "varimax_iterations": vm_iters,
                   "varimax_polar_jacobi_sweeps": info.get("varimax_svd_sweeps"),
                   "solve_rotate_wall_s": ms_step / 1e3,
                   "rule_n": rn, "cov_gemm": cov, "roofline_kernels": roof_list, "call_shares": shares,
                   "peaks": peaks_out},
        "solve_rotate_wall_s": ms_step / 1e3,
        "e2e": {"value": world / (ms_e2e / 1e3), "unit": "models/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "cov_gemm": cov,
        "rule_n": rn,
        "roofline": roof,
        "roofline_kernels": roof_list,
        "call_shares": shares,
    }
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args.workload)
    return out


# ------------------------------------------------------------ CPU baseline / reference arm
def _cpu_step_fn():
    """(callable(A, B, n_rot, n_modes, wo) -> seconds, kind): the live reference (oracle/_ref or /root/reference)
    when importable, else the numpy oracle port."""
    try:
        from oracle import ref_harness as rh
        if rh.reference_available():
            MCAref = rh.import_reference_mca()

            def step(A, B, n_rot, n_modes, wo):
                t0 = time.perf_counter()
                m = MCAref(A, B)
                m.solve(complexify=wo["complexify"])
                try:
                    m.rotate(n_rot, wo["power"])
                except RuntimeError:                    # Varimax did not converge (rotation.py:66-71)
                    pass
                m.singular_values(n_modes)
                m.pcs(n_modes)
                m.eofs(n_modes)
                return time.perf_counter() - t0
            return step, "reference"
    except Exception as exc:                            # noqa: BLE001 -- fall back to the port, say why
        sys.stderr.write("bench.py: live reference not importable (%s: %s); timing the oracle port\n"
                         % (type(exc).__name__, exc))
    from oracle import mca_oracle as orc

    def step(A, B, n_rot, n_modes, wo):
        t0 = time.perf_counter()
        m = orc.solve(orc.make_model(A, B), complexify=wo["complexify"])
        try:
            orc.rotate(m, n_rot, wo["power"])
        except orc.NotConverged:
            pass
        orc.pcs(m, n_modes)
        orc.eofs(m, n_modes)
        return time.perf_counter() - t0
    return step, "port"


def _ref_location():
    try:
        from oracle import ref_harness as rh
        return os.path.relpath(rh.reference_location(), ROOT) if rh.reference_location().startswith(ROOT) else rh.reference_location()
    except Exception:                                   # noqa: BLE001
        return "?"


def _all_cores():
    cores = os.cpu_count() or 1
    try:        # torchrun exports OMP_NUM_THREADS=1: give the reference all host cores anyway
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=cores)
    except Exception:                                   # noqa: BLE001 -- threadpoolctl is optional
        pass
    return cores


def _sample(step, T, S1, S2, n_rot, n_modes, seed, wo):
    A, B = synthetic_fields(T, S1, S2, seed=seed, dtype=wo["dtype"])
    return step(A, B, n_rot, n_modes, wo)


def cpu_baseline(workload):
    """Bounded CPU sample for the product line (about 10-30 s): the reference path at 1/2 and 1/4 of every axis
    (c3/c5: 1/4 and 1/8), with the measured per-doubling factor and the extrapolated full-size time LABELLED as
    extrapolated.  The measured full-size number is the reference arm's (`--impl reference`)."""
    T, S1, S2, n_rot, n_modes = WORKLOADS[workload]
    wo = opts(workload)
    div = {"c2": 2, "half": 1, "small": 1, "c3": 4, "c3half": 2, "c5": 8, "c5half": 4}.get(workload, 2)
    cores = _all_cores()
    step, kind = _cpu_step_fn()
    _sample(step, 256, 512, 512, 10, 10, 5, opts("c2"))          # BLAS/LAPACK warm-up
    t_hi = _sample(step, T // div, S1 // div, S2 // div, n_rot, n_modes, 77, wo)
    out = {"unit": "models/s", "cores": cores, "kind": kind}
    what = "%s %ssolve+rotate(%d, %d)+pcs/eofs(%d)" % (
        "live reference (xmca 1.4.2)" if kind == "reference" else "numpy oracle port",
        "complex " if wo["complexify"] else "", n_rot, wo["power"], n_modes)
    if div == 1:
        out.update({"value": 1.0 / t_hi, "sample": "%s at the full size T=%d,S1=%d: %.2f s measured" % (what, T, S1, t_hi),
                    "extrapolated": False, "seconds_per_model": t_hi})
        return out
    t_lo = _sample(step, T // (2 * div), S1 // (2 * div), S2 // (2 * div), n_rot, n_modes, 177, wo)
    f = max(t_hi / t_lo, 1.0)
    full = t_hi * f ** int(np.log2(div))
    out.update({"value": 1.0 / full, "extrapolated": True, "seconds_per_model": full,
                "seconds_sample": t_hi,
                "sample": "%s at T=%d,S1=%d: %.2f s measured; at T=%d,S1=%d: %.2f s; per-doubling factor %.2f -> "
                          "EXTRAPOLATED %.1f s at the full size (the measured full-size step is the reference arm's)"
                          % (what, T // div, S1 // div, t_hi, T // (2 * div), S1 // (2 * div), t_lo, f, full)})
    return out


def run_reference(args, rank, world):
    """The reference arm: ONE full-size step of the workload through the reference's own CPU path, measured."""
    if rank != 0:
        return None
    T, S1, S2, n_rot, n_modes = WORKLOADS[args.workload]
    wo = opts(args.workload)
    cores = _all_cores()
    step, kind = _cpu_step_fn()
    budget = float(os.environ.get("XMCA_REF_BUDGET_S", "480"))
    t_start = time.perf_counter()
    _sample(step, 256, 512, 512, 10, 10, 5, opts("c2"))          # BLAS/LAPACK warm-up
    t_half = _sample(step, T // 2, S1 // 2, S2 // 2, n_rot, n_modes, 77, wo)
    predicted = 8.0 * t_half                                     # O(T^2 S): x8 per doubling at most
    what = "%s %ssolve+rotate(%d, %d)+pcs/eofs(%d), %d host cores" % (
        "live reference (xmca 1.4.2 imported from %s)" % _ref_location() if kind == "reference" else "numpy oracle port",
        "complex " if wo["complexify"] else "", n_rot, wo["power"], n_modes, cores)
    if predicted <= budget:
        t_full = _sample(step, T, S1, S2, n_rot, n_modes, 78, wo)
        measured, extrapolated = 1, False
        sample = "%s: 1 full-size step measured: %.1f s (half-size step: %.2f s); steps/warmup of the command line are " \
                 "not repeated (one step is minutes of CPU time)" % (what, t_full, t_half)
    else:
        t_quarter = _sample(step, T // 4, S1 // 4, S2 // 4, n_rot, n_modes, 177, wo)
        f = max(t_half / t_quarter, 1.0)
        t_full = t_half * f
        measured, extrapolated = 0, True
        sample = "%s: half-size step %.1f s measured, quarter-size %.2f s; a full-size step (predicted > %.0f s budget) " \
                 "is EXTRAPOLATED with the measured per-doubling factor %.2f: %.1f s" % (what, t_half, t_quarter, budget, f, t_full)
    wall = time.perf_counter() - t_start
    cb = {"value": 1.0 / t_full, "unit": "models/s", "cores": cores, "kind": kind, "sample": sample,
          "extrapolated": extrapolated, "measured_full_size_steps": measured, "seconds_per_model": t_full}
    return {
        "impl": "reference",
        "metric": "solve()+rotate() throughput (models/s; wall-sec in solve_rotate_wall_s)",
        "value": cb["value"], "unit": "models/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "measured_steps": measured, "ms_per_step": t_full * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "%s fields (numpy/LAPACK gesdd in the field dtype), f64 rotation"
                                      % np.dtype(wo["dtype"]).name, "data": "synthetic",
        "config": {"workload": workload_string(args.workload)},
        "solve_rotate_wall_s": t_full,
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "models/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "bench_wall_s": wall,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--rule-n-total", type=int, default=128,
                    help="strong-scaling rule_n: total number of float64 surrogates split over the ranks (0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        # torchrun exports OMP_NUM_THREADS=1, which caps the BLAS thread pool when numpy is loaded: give the
        # reference's CPU path all host cores by re-running this arm in a child with a clean thread environment
        capped = [v for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS") if v in os.environ]
        if capped and not os.environ.get("XMCA_BENCH_CHILD"):
            env = {k: v for k, v in os.environ.items() if k not in capped}
            env["XMCA_BENCH_CHILD"] = "1"
            env["RANK"], env["WORLD_SIZE"] = "0", str(world)
            res = subprocess.run([sys.executable, os.path.abspath(__file__)] + sys.argv[1:], env=env,
                                 stdout=subprocess.PIPE, text=True)
            sys.stdout.write(res.stdout)
            sys.stdout.flush()
            return
        out = run_reference(args, rank, world)
        if out is not None:
            print(json.dumps(out), flush=True)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product arm has no CPU fallback")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    out = run_product(args, rank, world, local_rank)
    if out is not None:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
