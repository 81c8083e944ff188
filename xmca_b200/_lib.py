"""ctypes binding of ``libxmca_b200.so`` (the C ABI declared in
``include/xmca_b200.h``).

There is NO fallback: if the shared library is missing or a CUDA device is not
available, every compute entry point raises.  torch tensors are used purely as
device buffers -- their ``data_ptr()`` goes across the C ABI together with the
raw ``cudaStream_t`` of torch's current stream.
"""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libxmca_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "xmca_b200.h")

F32, F64 = 0, 1
GEMM_SYMMETRIC, GEMM_A_LOWER_T, GEMM_B_LOWER = 1, 2, 4
OK, BAD_ARG, CUDA_ERROR, NOT_CONVERGED, NUMERIC = 0, 1, 2, 3, 4


class XmcaLibraryError(RuntimeError):
    pass


class NotConvergedError(RuntimeError):
    pass


_lib = None


def declared_symbols():
    """Function names declared in include/xmca_b200.h (used by the CPU tests)."""
    with open(HEADER_PATH) as fh:
        text = fh.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(xmca_[a-z0-9_]+)\s*\(", text)))


def load():
    """dlopen the library (works without a GPU; compute calls need one)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise XmcaLibraryError(
            "libxmca_b200.so not found at %s -- run `python -c 'import __graft_entry__ as g; g.build()'`; "
            "there is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    i64, i32, dbl, vp, sz = C.c_int64, C.c_int, C.c_double, C.c_void_p, C.c_size_t
    lib.xmca_last_error.restype = C.c_char_p
    lib.xmca_version.restype = i32
    lib.xmca_launch_count.restype = C.c_longlong
    lib.xmca_gemm_workspace_bytes.restype = sz
    lib.xmca_gemm_workspace_bytes.argtypes = [i64, i64, i32, i32]
    lib.xmca_gemm.argtypes = [i32, i32, i64, i64, i64, dbl, vp, i32, i64, vp, i32, i64, vp, i32, i64,
                              i32, i32, i32, vp, sz, vp]
    lib.xmca_gemm_ex.argtypes = [i32, i32, i64, i64, i64, dbl, vp, i32, i64, vp, i32, i64, vp, i32, i64,
                                 i32, i32, i32, vp, sz, i32, vp]
    lib.xmca_split_tf32.argtypes = [vp, i32, i64, i64, i64, i32, vp, vp, i64, vp]
    lib.xmca_tc_gemm_nt.argtypes = [i64, i64, i64, C.c_float, vp, vp, i64, vp, vp, i64, vp, i64, vp, vp]
    lib.xmca_tc_gemm_nt_f64.argtypes = [i64, i64, i64, dbl, vp, vp, i64, vp, vp, i64, vp, i64, i32, vp]
    lib.xmca_jacobi_padded_cols.restype = i64
    lib.xmca_jacobi_padded_cols.argtypes = [i64]
    lib.xmca_jacobi_workspace_bytes.restype = sz
    lib.xmca_jacobi_workspace_bytes.argtypes = [i64, i64]
    lib.xmca_jacobi_svd.argtypes = [i64, i64, vp, i64, vp, i64, vp, i32, dbl, C.POINTER(i32),
                                    C.POINTER(dbl), vp, sz, vp]
    lib.xmca_cholesky_workspace_bytes.restype = sz
    lib.xmca_cholesky_workspace_bytes.argtypes = [i64]
    lib.xmca_cholesky_invdiag_bytes.restype = sz
    lib.xmca_cholesky_invdiag_bytes.argtypes = [i64]
    lib.xmca_cholesky.argtypes = [i64, vp, i64, vp, dbl, C.POINTER(i32), vp, sz, vp]
    lib.xmca_trsm_workspace_bytes.restype = sz
    lib.xmca_trsm_workspace_bytes.argtypes = [i64, i64]
    lib.xmca_trsm_lt.argtypes = [i64, i64, vp, i64, vp, vp, i64, vp, sz, vp]
    lib.xmca_sytrd_max_n.restype = i64
    lib.xmca_sytrd_max_n.argtypes = []
    lib.xmca_sytrd_workspace_bytes.restype = sz
    lib.xmca_sytrd_workspace_bytes.argtypes = [i64]
    lib.xmca_sytrd.argtypes = [i64, vp, i64, vp, vp, vp, vp, sz, vp]
    lib.xmca_sytrd_batched.argtypes = [i64, i32, vp, i64, i64, vp, vp, vp, i64, vp, sz, vp]
    lib.xmca_stebz.argtypes = [i64, vp, vp, vp, vp, vp]
    lib.xmca_stein_workspace_bytes.restype = sz
    lib.xmca_stein_workspace_bytes.argtypes = [i64, i64]
    lib.xmca_stein.argtypes = [i64, vp, vp, i64, vp, vp, i64, dbl, i32, vp, i64, vp, sz, vp]
    lib.xmca_ormtr.argtypes = [i64, vp, i64, vp, i64, vp, i64, vp]
    lib.xmca_sytrd2_workspace_bytes.restype = sz
    lib.xmca_sytrd2_workspace_bytes.argtypes = [i64]
    lib.xmca_sytrd2_tfac_bytes.restype = sz
    lib.xmca_sytrd2_tfac_bytes.argtypes = [i64]
    lib.xmca_sytrd2_info_offset.restype = sz
    lib.xmca_sytrd2_info_offset.argtypes = [i64]
    lib.xmca_sytrd2.argtypes = [i64, vp, i64, vp, vp, vp, i32, vp, sz, vp]
    lib.xmca_ormtr2_workspace_bytes.restype = sz
    lib.xmca_ormtr2_workspace_bytes.argtypes = [i64, i64]
    lib.xmca_ormtr2.argtypes = [i64, vp, i64, vp, i64, vp, i64, vp, sz, vp]
    lib.xmca_hilbert_matrix.argtypes = [i64, vp, i32, i64, vp, vp]
    lib.xmca_hilbert_block.argtypes = [i64, i64, i64, i64, i64, vp, i32, i64, vp, vp]
    lib.xmca_dft_rows.restype = i64
    lib.xmca_dft_rows.argtypes = [i64]
    lib.xmca_dft_matrix.argtypes = [i64, vp, i32, i64, vp]
    lib.xmca_embed_complex.argtypes = [vp, i32, i64, i64, i64, vp, i32, i64, vp]
    lib.xmca_scale_copy.argtypes = [vp, i32, i64, vp, i32, i64, i64, i64, vp, vp, vp]
    lib.xmca_transpose.argtypes = [vp, i32, i64, i64, i64, vp, i32, i64, vp]
    lib.xmca_col_sumsq.argtypes = [vp, i32, i64, i64, i64, i64, vp, vp]
    lib.xmca_center_columns.argtypes = [vp, i32, i64, i64, i64, vp, vp]
    lib.xmca_field_stats.argtypes = [vp, i32, i64, i64, i64, vp, vp, vp, vp, vp]
    lib.xmca_compact_center.argtypes = [vp, i32, i64, i64, vp, i64, vp, vp, i32, i64, vp]
    lib.xmca_fill_normal.argtypes = [vp, i32, i64, i64, i64, C.c_uint64, C.c_uint64, vp]
    lib.xmca_gather_rows.argtypes = [vp, i32, i64, vp, i64, i64, vp, vp, i32, i64, vp]
    lib.xmca_row_sumsq.argtypes = [vp, i32, i64, i64, i64, vp, vp]
    lib.xmca_col_absmax.argtypes = [vp, i32, i64, i64, i64, vp, vp, vp]
    lib.xmca_promax_target.argtypes = [vp, i64, i64, i64, vp, vp, dbl, vp, vp, i64, vp]
    lib.xmca_col_absmax_complex.argtypes = [vp, vp, i64, i64, i64, vp, vp, vp]
    lib.xmca_promax_target_complex.argtypes = [vp, vp, i64, i64, i64, vp, vp, dbl, vp, vp, vp, vp, i64, vp]
    lib.xmca_varimax_workspace_bytes.restype = sz
    lib.xmca_varimax_workspace_bytes.argtypes = [i64, i32]
    lib.xmca_varimax.argtypes = [vp, i32, i64, i32, i64, dbl, i32, dbl, vp, i64, vp, C.POINTER(i32), vp,
                                 vp, sz, vp]
    lib.xmca_varimax_complex_workspace_bytes.restype = sz
    lib.xmca_varimax_complex_workspace_bytes.argtypes = [i64, i32]
    lib.xmca_varimax_complex.argtypes = [vp, vp, i32, i64, i32, i64, dbl, i32, dbl, vp, vp, i64, vp, vp,
                                         C.POINTER(i32), vp, vp, sz, vp]
    for name in declared_symbols():
        fn = getattr(lib, name)          # raises AttributeError if a declared symbol is not exported
        if fn.restype is C.c_int and name not in ("xmca_version",):
            fn.restype = i32
    _lib = _Profiled(lib)
    return _lib


# ------------------------------------------------------------------ profiling
# bench.py asks for the device time of every C-ABI call class: CUDA events are
# recorded on torch's current stream (the stream every call is enqueued on)
# around each call while a profile is open.  Disabled -> one attribute lookup.
_NO_TIMING = ("xmca_last_error", "xmca_version", "xmca_launch_count", "xmca_gemm_workspace_bytes",
              "xmca_jacobi_padded_cols", "xmca_jacobi_workspace_bytes", "xmca_varimax_workspace_bytes",
              "xmca_cholesky_workspace_bytes", "xmca_cholesky_invdiag_bytes", "xmca_trsm_workspace_bytes",
              "xmca_sytrd_max_n", "xmca_sytrd_workspace_bytes", "xmca_sytrd2_workspace_bytes", "xmca_sytrd2_tfac_bytes", "xmca_sytrd2_info_offset", "xmca_ormtr2_workspace_bytes", "xmca_stein_workspace_bytes", "xmca_dft_rows",
              "xmca_varimax_complex_workspace_bytes")
_profile = None


class _Profiled:
    def __init__(self, lib):
        self._raw = lib
        self._cache = {}

    def __getattr__(self, name):
        fn = self._cache.get(name)
        if fn is None:
            raw = getattr(self._raw, name)
            if name in _NO_TIMING or not name.startswith("xmca_"):
                fn = raw
            else:
                def fn(*args, _raw=raw, _name=name):
                    if _profile is None:
                        return _raw(*args)
                    import torch
                    e0 = torch.cuda.Event(enable_timing=True)
                    e1 = torch.cuda.Event(enable_timing=True)
                    n0 = int(self._raw.xmca_launch_count())
                    e0.record()
                    rc = _raw(*args)
                    e1.record()
                    _profile.append((_name, e0, e1, int(self._raw.xmca_launch_count()) - n0))
                    return rc
            self._cache[name] = fn
        return fn


def profile_begin():
    """Start recording (name, start event, end event, launches) per C-ABI call."""
    global _profile
    _profile = []


def profile_end():
    """Stop recording; returns {name: {"calls", "launches", "ms"}} (synchronises)."""
    global _profile
    rec, _profile = _profile or [], None
    import torch
    torch.cuda.synchronize()
    out = {}
    for name, e0, e1, nl in rec:
        d = out.setdefault(name, {"calls": 0, "launches": 0, "ms": 0.0})
        d["calls"] += 1
        d["launches"] += nl
        d["ms"] += e0.elapsed_time(e1)
    return out


def cuda_available() -> bool:
    try:
        import torch
        return bool(torch.cuda.is_available())
    except Exception:
        return False


def launch_count() -> int:
    return int(load().xmca_launch_count())


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise XmcaLibraryError("xmca_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.")
    return torch


def dtype_code(t) -> int:
    torch = _torch()
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.float64:
        return F64
    raise TypeError("unsupported device dtype %s" % t.dtype)


def stream_ptr():
    return C.c_void_p(_torch().cuda.current_stream().cuda_stream)


def check(rc: int, what: str):
    if rc == OK:
        return
    msg = load().xmca_last_error().decode("utf-8", "replace")
    if rc == NOT_CONVERGED:
        raise NotConvergedError(msg)
    if rc == NUMERIC:
        raise np.linalg.LinAlgError(msg)
    if rc == BAD_ARG:
        raise ValueError("%s: %s" % (what, msg))
    raise XmcaLibraryError("%s failed (code %d): %s" % (what, rc, msg))


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
