"""numpy restatement of the xmca solve / rotate / getters / rule_n hot path.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

Parity status
-------------
* solve (sigma, unrotated EOFs): PINNED against the reference's own golden
  fixtures ``tests/integration/fixtures/{std,cplx}`` (committed, re-packed, as
  ``tests/golden/fixtures.npz``) and against the live reference imported in the
  build container (``oracle/ref_harness.py``; outputs in
  ``tests/golden/live_*.npz``).
* rotate / pcs: the reference ships NO golden vectors for rotation (SURVEY.md
  section 8c).  Pinned against outputs of the live reference only.
* rule_n: "parity unpinned" by the reference (smoke test only); pinned here by
  replaying the identical numpy global-RNG stream against the live reference.

Every function cites the reference lines (``/root/reference/xmca/...``) whose
arithmetic it restates.  Third-party arithmetic the reference delegates to
(not vendored under /root/reference): ``numpy.linalg.svd/inv/pinv`` (LAPACK
gesdd, numpy unpinned ``>=1.19.2`` in requirements.txt:1; 2.3.5 here) and
``scipy.signal.hilbert`` (scipy 1.18.1 here) -- the latter is restated below
from its published algorithm and checked against scipy in the tests.
"""
from __future__ import annotations

import cmath
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

KEYS = ("left", "right")


# --------------------------------------------------------------------------
# pre-processing  (array.py:199-240, tools/array.py:26-73)
# --------------------------------------------------------------------------
def flatten_field(x: np.ndarray) -> np.ndarray:
    """(T, *space) -> (T, prod(space)); array.py:230-240."""
    return x.reshape(x.shape[0], int(np.prod(x.shape[1:])))


def valid_columns(x2d: np.ndarray) -> np.ndarray:
    """True where a column holds no NaN; tools/array.py:26-41 + array.py:217-220."""
    return ~np.isnan(x2d).any(axis=0)


def center(x2d: np.ndarray) -> np.ndarray:
    """Remove the time mean, dtype preserved; array.py:199-207."""
    return x2d - x2d.mean(axis=0)


def analytic_signal(x: np.ndarray) -> np.ndarray:
    """Analytic signal along axis 0 (what ``scipy.signal.hilbert(x, axis=0)``
    returns, array.py:464): FFT, zero the negative frequencies, double the
    positive ones (DC and -- for even N -- Nyquist kept once), inverse FFT.
    f32 in -> c64 out, f64 in -> c128 out (scipy.fft keeps single precision)."""
    n = x.shape[0]
    spec = np.fft.fft(x.astype(np.float64), axis=0)
    w = np.zeros(n)
    if n % 2 == 0:
        w[0] = w[n // 2] = 1.0
        w[1:n // 2] = 2.0
    else:
        w[0] = 1.0
        w[1:(n + 1) // 2] = 2.0
    out = np.fft.ifft(spec * w[:, None], axis=0)
    return out.astype(np.complex64 if x.dtype == np.float32 else np.complex128)


# --------------------------------------------------------------------------
# model state
# --------------------------------------------------------------------------
@dataclass
class OracleModel:
    keys: List[str]
    fields: Dict[str, np.ndarray]            # centred, NaN columns removed (T x S')
    spatial_shape: Dict[str, tuple]
    n_variables: Dict[str, int]              # full grid size incl. NaN columns
    keep: Dict[str, np.ndarray]              # bool mask of kept columns
    n_obs: int
    is_complex: bool = False
    # filled by solve()
    sigma: Optional[np.ndarray] = None
    V: Dict[str, np.ndarray] = field(default_factory=dict)
    rank: int = 0
    total_covariance: float = 0.0
    total_squared_covariance: float = 0.0
    # rotation state
    is_rotated: bool = False
    n_rot: int = 0
    power: int = 0
    R: Optional[np.ndarray] = None
    Phi: Optional[np.ndarray] = None
    norm: Dict[str, np.ndarray] = field(default_factory=dict)
    variance: Optional[np.ndarray] = None
    var_idx: Optional[np.ndarray] = None

    @property
    def bivariate(self) -> bool:
        return len(self.keys) == 2


def make_model(*arrays: np.ndarray) -> OracleModel:
    """Constructor semantics of array.py:39-143 (validation lives in the
    product class; the oracle assumes valid input)."""
    keys = list(KEYS[:len(arrays)])
    flds, shp, nvar, keep = {}, {}, {}, {}
    for k, a in zip(keys, arrays):
        shp[k] = a.shape[1:]
        nvar[k] = int(np.prod(a.shape[1:]))
        x = flatten_field(a)
        keep[k] = valid_columns(x)
        flds[k] = center(x[:, keep[k]])
    return OracleModel(keys, flds, shp, nvar, keep, arrays[0].shape[0])


# --------------------------------------------------------------------------
# solve  (array.py:509-603)
# --------------------------------------------------------------------------
def solve(m: OracleModel, complexify: bool = False) -> OracleModel:
    m.is_complex = complexify
    dof = m.n_obs - 1
    if complexify:                                   # array.py:546-547, :455-464
        m.fields = {k: analytic_signal(np.real(f)) for k, f in m.fields.items()}
    # per-field thin SVD, array.py:474-483
    U, s, Vt = {}, {}, {}
    for k in m.keys:
        U[k], s[k], Vt[k] = np.linalg.svd(m.fields[k], full_matrices=False)
    Rl = U["left"] * s["left"]                       # array.py:553
    Rr = U["right"] * s["right"] if m.bivariate else Rl
    kernel = Rl.conj().T @ Rr / dof                  # array.py:556-566
    P, sigma, Qh = np.linalg.svd(kernel, full_matrices=False)   # array.py:570
    small = {"left": P, "right": Qh.conj().T}
    m.V = {k: Vt[k].conj().T @ small[k] for k in m.keys}        # array.py:584
    m.sigma = sigma
    m.rank = len(sigma)
    m.total_covariance = float(sigma.sum())          # array.py:595
    m.total_squared_covariance = float((sigma ** 2).sum())
    m.is_rotated, m.n_rot, m.power = False, m.rank, 0           # array.py:598-600
    m.R = np.eye(m.rank)
    m.Phi = np.eye(m.rank)
    m.norm = {k: np.sqrt(sigma) for k in m.keys}     # array.py:593
    m.variance = sigma
    m.var_idx = np.argsort(sigma)[::-1]
    return m


# --------------------------------------------------------------------------
# mode slicing and getters  (array.py:145-173, :605-779)
# --------------------------------------------------------------------------
def mode_slice(m: OracleModel, n) -> slice:
    """int n -> modes [0,n); slice(a,b) -> 1-based inclusive; array.py:153-167."""
    if n is None:
        return slice(0, m.rank)
    if isinstance(n, (int, np.integer)):
        return slice(0, int(n))
    if isinstance(n, slice):
        lo = 0 if n.start is None else max(0, n.start - 1)
        hi = m.rank if n.stop is None else min(n.stop, m.rank)
        return slice(lo, hi, n.step)
    raise ValueError("Invalid type {:}. Must be either int or slice.".format(type(n)))


def get_norm(m: OracleModel, n=None, sorted: bool = True):
    out = {k: v[m.var_idx] if sorted else v for k, v in m.norm.items()}   # array.py:755-770
    return {k: v[mode_slice(m, n)] for k, v in out.items()}


def get_variance(m: OracleModel, n=None, sorted: bool = True) -> np.ndarray:
    nr = get_norm(m, n, sorted)                       # array.py:772-779
    return nr["left"] * nr["right"] if m.bivariate else nr["left"] ** 2


def rotation_matrix(m: OracleModel, inverse_transpose: bool = False) -> np.ndarray:
    R = m.R                                           # array.py:846-876
    if inverse_transpose and m.power > 1:
        R = np.linalg.pinv(R).conj().T
    return R


def get_V(m: OracleModel, n=None, rotated: bool = True):
    top = m.n_rot if rotated else (n.stop if isinstance(n, slice) else n)   # array.py:615-622
    out = {}
    for k in m.keys:
        v = m.V[k][:, :top]
        if rotated:                                   # array.py:634-642
            root = np.sqrt(m.sigma[mode_slice(m, top)])
            nrm = get_norm(m, top, sorted=False)[k]
            v = (v * root @ rotation_matrix(m) / nrm)[:, m.var_idx]
        out[k] = v[:, mode_slice(m, n)]
    return out


def get_U(m: OracleModel, n=None, rotated: bool = True):
    top = m.n_rot if rotated else (n.stop if isinstance(n, slice) else n)   # array.py:648-656
    Vun = get_V(m, top, rotated=False)
    root = np.sqrt(m.sigma[mode_slice(m, top)])
    Rit = rotation_matrix(m, inverse_transpose=True)
    out = {}
    for k in m.keys:
        u = m.fields[k] @ Vun[k] / root               # array.py:667
        if rotated:
            u = (u @ Rit)[:, m.var_idx]               # array.py:668-671
        out[k] = u[:, mode_slice(m, n)]
    return out


def _scale(m, arr, kind, k, n_for_norm, real_axes):
    if kind == "None":
        return arr
    if kind == "eigen":
        return arr * get_norm(m, n_for_norm, sorted=True)[k]
    if kind == "max":
        return arr / np.nanmax(abs(arr.real), axis=real_axes)
    if kind == "std":
        return arr / np.nanstd(arr.real, axis=real_axes)
    raise ValueError("The scaling option {:} is not valid.".format(kind))


def eofs(m: OracleModel, n=None, scaling="None", phase_shift=0, rotated=True):
    """array.py:676-721: scatter into the full grid (NaN at dropped columns),
    reshape to space + (mode,), phase shift, scaling."""
    V = get_V(m, n, rotated)
    out = {}
    for k in m.keys:
        nm = V[k].shape[1]
        full = np.zeros((m.n_variables[k], nm), dtype=V[k].dtype) * np.nan
        full[m.keep[k], :] = V[k]
        full = full.reshape(m.spatial_shape[k] + (nm,))
        if m.is_complex:
            full = full * cmath.rect(1, phase_shift)
        axes = tuple(range(full.ndim - 1))
        out[k] = _scale(m, full, scaling, k, V["left"].shape[1], axes)
    return out


def pcs(m: OracleModel, n=None, scaling="None", phase_shift=0, rotated=True):
    """array.py:723-753."""
    U = get_U(m, n, rotated)
    out = {}
    for k in m.keys:
        u = U[k]
        if m.is_complex:
            u = u * cmath.rect(1, phase_shift)
        out[k] = _scale(m, u, scaling, k, n, 0)
    return out


# --------------------------------------------------------------------------
# Varimax / Promax  (tools/rotation.py:15-149)
# --------------------------------------------------------------------------
class NotConverged(RuntimeError):
    pass


def varimax(A: np.ndarray, gamma: float = 1.0, max_iter: int = 1000, tol: float = 1e-8):
    """Kaiser-normalised Varimax fixed point; rotation.py:38-77.  Runs in
    f64/c128 whatever the input dtype because R starts as a float64 identity
    (rotation.py:41).  Returns (B, R, iterations)."""
    n, p = A.shape
    h = np.sqrt(np.sum(A * A.conj(), axis=1))         # rotation.py:46
    An = A / h[:, None]                               # rotation.py:48
    R = np.eye(p)
    d = 0.0
    for it in range(1, max_iter + 1):
        d_old = d
        B = An @ R                                    # rotation.py:54
        crit = An.conj().T @ (B ** 2 * B.conj()
                              - (gamma / n) * (B @ np.diag(np.sum(B * B.conj(), axis=0))))
        u, s, vh = np.linalg.svd(crit)                # rotation.py:59
        R = u @ vh
        d = s.sum()
        if abs(d - d_old) / d < tol:                  # rotation.py:62
            return (h[:, None] * An) @ R, R, it       # rotation.py:74-77
    raise NotConverged("Rotation process did not converge.")


def promax(A: np.ndarray, power: int = 1, max_iter: int = 1000, tol: float = 1e-8):
    """rotation.py:103-149.  Returns (B, R, Phi, varimax_iterations)."""
    X, R, iters = varimax(A, max_iter=max_iter, tol=tol)
    h = np.sqrt(np.sum(X * X.conj(), axis=1))         # rotation.py:115
    X = X / h[:, None]
    Xn = X / np.max(abs(X), axis=0)                   # rotation.py:121
    P = Xn * np.abs(Xn) ** (power - 1)                # rotation.py:124
    L = np.linalg.inv(X.conj().T @ X) @ X.conj().T @ P        # rotation.py:128
    try:
        d = np.diag(np.linalg.inv(L.conj().T @ L))    # rotation.py:131-134
    except np.linalg.LinAlgError:
        d = np.diag(np.linalg.pinv(L.conj().T @ L))
    L = L @ np.sqrt(np.diag(d))                       # rotation.py:137
    B = h[:, None] * (X @ L)                          # rotation.py:138-141
    Li = np.linalg.inv(L)
    return B, R @ L, Li @ Li.conj().T, iters          # rotation.py:143-147


def rotate(m: OracleModel, n_rot: int, power: int = 1, tol: float = 1e-8) -> OracleModel:
    """array.py:810-844."""
    if n_rot < 2:
        raise ValueError("`n_rot` must be > 1")
    if power < 1:
        raise ValueError("`power` must be >=1")
    root = np.sqrt(m.sigma[:n_rot])
    Vun = get_V(m, n_rot, rotated=False)
    s_left = Vun["left"].shape[0]
    L = np.concatenate([Vun[k] for k in m.keys]) * root        # array.py:821-822
    Lr, R, Phi, _ = promax(L, power, max_iter=1000, tol=tol)
    nl = np.linalg.norm(Lr[:s_left], axis=0)          # array.py:826-830
    nr = np.linalg.norm(Lr[s_left:], axis=0) if m.bivariate else nl
    m.norm = {"left": nl, "right": nr}
    m.variance = nl * nr
    m.var_idx = np.argsort(m.variance)[::-1]
    m.R, m.Phi = R, Phi
    m.is_rotated, m.n_rot, m.power = True, n_rot, power
    return m


# --------------------------------------------------------------------------
# Rule N  (array.py:1744-1771)
# --------------------------------------------------------------------------
def rule_n(m: OracleModel, n_runs: int, n_modes=None, rng=np.random) -> np.ndarray:
    """Gaussian surrogates of the FULL grid size (NaN columns included,
    array.py:1745), always float64, global numpy RNG (array.py:1756)."""
    spectra = []
    for _ in range(n_runs):
        data = [rng.standard_normal([m.n_obs, m.n_variables[k]]) for k in m.keys]
        s = solve(make_model(*data), complexify=m.is_complex)
        if m.is_rotated:
            try:
                rotate(s, m.n_rot, m.power)
            except NotConverged:
                continue                              # array.py:1759-1763
        spectra.append(get_variance(s))
    sv = np.array(spectra).T
    sv /= sv.sum(axis=0) / get_variance(m).sum()      # array.py:1768-1769
    return sv[mode_slice(m, n_modes)]


# --------------------------------------------------------------------------
# comparison helpers used by the parity tests
# --------------------------------------------------------------------------
def align_modes(ref_left, got_left, *others):
    """Joint sign (real) / phase (complex) alignment: per mode take <ref, got>
    on the LEFT field and apply the same unit factor to every array passed in
    ``others`` (and to got_left).  NaNs are treated as zeros."""
    r = np.nan_to_num(ref_left).reshape(-1, ref_left.shape[-1])
    g = np.nan_to_num(got_left).reshape(-1, got_left.shape[-1])
    ip = np.sum(r.conj() * g, axis=0)
    mag = np.abs(ip)
    fac = np.where(mag > 0, ip.conj() / np.where(mag > 0, mag, 1), 1)
    if not np.iscomplexobj(fac):
        fac = np.sign(fac) + (fac == 0)
    return [got_left * fac] + [o * fac for o in others]


def subspace_angle(X: np.ndarray, Y: np.ndarray) -> float:
    """Largest principal angle between span(X) and span(Y) (radians)."""
    qx, _ = np.linalg.qr(np.nan_to_num(X).reshape(-1, X.shape[-1]).astype(np.complex128))
    qy, _ = np.linalg.qr(np.nan_to_num(Y).reshape(-1, Y.shape[-1]).astype(np.complex128))
    # sin of the largest angle = || (I - qx qx^H) qy ||_2
    resid = qy - qx @ (qx.conj().T @ qy)
    return float(np.arcsin(min(1.0, np.linalg.norm(resid, 2))))


def synthetic_fields(T: int, S1: int, S2: int, seed: int = 0, k: int = 64,
                     dtype=np.float32):
    """Low-rank-plus-noise generator of SURVEY.md section 8(d)."""
    rng = np.random.default_rng(seed)
    ts = rng.standard_normal((T, k))
    amp = 3.0 * np.sqrt(max(S1, S2)) * 0.9 ** np.arange(k) / np.sqrt(k)
    pa = rng.standard_normal((k, S1)) / np.sqrt(S1)
    pb = rng.standard_normal((k, S2)) / np.sqrt(S2)
    A = (ts * amp) @ pa + rng.standard_normal((T, S1))
    B = (ts * amp) @ pb + rng.standard_normal((T, S2))
    return A.astype(dtype), B.astype(dtype)
