"""Callers immediately downstream of the hot path: ``predict``,
``reconstructed_fields``, homogeneous / heterogeneous patterns and ``bootstrapping``
(xmca/array.py:1188-1428, :1813-1952; xmca/tools/array.py:76-138).

They reuse the engine's device products: every T x S (or S x m) matrix product runs through the
C ABI, the host keeps the reference's bookkeeping (NaN scatter, scaling options, error messages)
and the O(S m) Beta-distribution p-values.  Bound to ``MCA`` in ``array.py``.
"""
from __future__ import annotations

import cmath

import numpy as np

from . import _lib as L
from . import device as D
from . import engine as E


# ------------------------------------------------------------------ scaling helpers
def scale_X(self, data_dict):
    """array.py:263-273 -- including its quirk: the division by the standard deviation sits
    OUTSIDE the loop, so only the last field of `data_dict` is normalised."""
    scaled = dict(data_dict)
    k = None
    for k in scaled:
        scaled[k] = scaled[k] - self._field_means[k]
    if self._analysis["is_normalized"] and k is not None:
        scaled[k] = scaled[k] / self._field_stds[k]
    return scaled


def scale_X_inverse(self, data_dict):
    """array.py:275-287."""
    out = {}
    for k, field in data_dict.items():
        if self._analysis["is_normalized"]:
            field = field * self._field_stds[k]
        out[k] = field + self._field_means[k]
    return out


def _scatter_space(self, k, arr):
    """(S' x m) -> space + (m,) with NaN at the dropped grid points (array.py:1204-1210)."""
    nm = arr.shape[1]
    full = np.zeros([self._n_variables[k], nm], dtype=arr.dtype) * np.nan
    full[self._no_nan_index[k], :] = arr
    return full.reshape(self._fields_spatial_shape[k] + (nm,))


# ------------------------------------------------------------------ predict
def predict(self, left=None, right=None, n=None, scaling="None", phase_shift=0):
    """Project new data onto the (rotated) singular vectors (array.py:1299-1428)."""
    self._require_solved("principal components")
    data_new = {k: d.copy() for k, d in zip(self._keys, [left, right]) if d is not None}
    R = self.rotation_matrix(inverse_transpose=True) if self._analysis["is_rotated"] else None
    n_rot = self._analysis["n_rot"]
    if n is None:
        n = n_rot
    # an unrotated model has R = I(rank) (array.py:1343): only the n modes returned are projected
    need = n_rot if R is not None else int(min(n, self._singular_values.size))
    sqrt_sv = np.sqrt(self._singular_values[:need].astype(np.float64))
    out = {}
    for k, x_new in data_new.items():
        try:
            x_new = x_new.reshape(x_new.shape[0], self._n_variables[k])
            x_new = x_new[:, self._no_nan_index[k]]
        except ValueError as err:
            if len(x_new.shape) != len(self._shape[k]):
                msg = ("Error in {:} field. Dimension of new data ({:}) and the original field ({:}) do not "
                       "match. Did you forget the time dimension?").format(k, len(x_new.shape), len(self._shape[k]))
            elif x_new.shape[1:] != self._field_means[k].shape:
                msg = ("Error in {:} field. Spatial dimensions of new data {:} and the original field {:} do "
                       "not match.").format(k, x_new.shape[1:], self._shape[k][1:])
            else:
                msg = "Dimension mismatch in {:} field.".format(k)
            raise ValueError(msg) from err
        x_new = scale_X(self, {k: x_new})[k]
        xd = D.to_device(np.ascontiguousarray(x_new))
        kind, _ = self._dV
        inv = D.to_device(1.0 / sqrt_sv)
        if kind == "real":
            pcs = D.to_host(D.scale_copy(D.matmul(xd, self._V_device_cols(k, need)), col_scale=inv))
        else:       # the reference multiplies the REAL new data with the complex vectors (no Hilbert transform)
            vr, vi = self._V_device_cols(k, need)
            pcs = (D.to_host(D.scale_copy(D.matmul(xd, vr), col_scale=inv))
                   + 1j * D.to_host(D.scale_copy(D.matmul(xd, vi), col_scale=inv)))
        if R is not None:
            pcs = (pcs @ R)[:, self._var_idx]
        pcs = pcs[:, :n]
        if self._analysis["is_complex"]:
            pcs = pcs * cmath.rect(1, phase_shift)
        if scaling == "None":
            pass
        elif scaling == "eigen":
            pcs = pcs * self._get_norm(n, sorted=True)[k]
        elif scaling == "max":
            pcs = pcs / np.nanmax(abs(self._get_pcs(n, "None", phase_shift)[k].real), axis=0)
        elif scaling == "std":
            pcs = pcs / np.nanstd(self._get_pcs(n, "None", phase_shift)[k].real, axis=0)
        else:
            raise ValueError("The scaling option {:} is not valid. Please choose one of the following: "
                             "None, eigen, std, max".format(scaling))
        out[k] = pcs
    return out


# ------------------------------------------------------------------ reconstruction
def _reconstruct_into(self, k, u, v, out=None, alpha=1.0):
    """out (+)= alpha Re(u v^H) on the device (T x S'); u: T x m, v: S' x m host arrays."""
    f64 = np.float64
    ur, vr = D.to_device(np.ascontiguousarray(u.real, dtype=f64)), D.to_device(np.ascontiguousarray(v.real, dtype=f64))
    acc = out is not None
    X = D.matmul(ur, vr, trans_b=True, alpha=alpha, out=out, accumulate=acc)
    if np.iscomplexobj(u) or np.iscomplexobj(v):                             # Re(u v^H) = ur vr^T + ui vi^T
        ui = D.to_device(np.ascontiguousarray(np.imag(u), dtype=f64))
        vi = D.to_device(np.ascontiguousarray(np.imag(v), dtype=f64))
        D.matmul(ui, vi, trans_b=True, alpha=alpha, out=X, accumulate=True)
    return X


def reconstructed_X(self, mode=None, original_scale=True):
    """Low-rank reconstruction U_eigen V^H, real part (array.py:1263-1276); host arrays."""
    V = self._get_V(n=mode, rotated=True)
    U = self._get_pcs(n=mode, scaling="eigen", rotated=True)
    Xrec = {}
    for k in self._keys:
        u, v = U[k], V[k]
        if u.shape[1] == 0:
            Xrec[k] = np.zeros((u.shape[0], v.shape[0]))
        else:
            Xrec[k] = D.to_host(_reconstruct_into(self, k, u, v))
    if original_scale:
        Xrec = scale_X_inverse(self, Xrec)
    return Xrec


def reconstructed_fields(self, mode=None, original_scale=True):
    """array.py:1278-1297."""
    Xrec = reconstructed_X(self, mode=mode, original_scale=original_scale)
    n_obs = self._n_observations["left"]
    out = {}
    for k, x in Xrec.items():
        full = np.zeros((n_obs, self._n_variables[k])) * np.nan
        full[:, self._no_nan_index[k]] = x
        out[k] = full.reshape((-1,) + self._fields_spatial_shape[k])
    return out


# ------------------------------------------------------------------ correlation patterns
def _pearson(self, k, y):
    """Correlation coefficients and two-sided p-values between every grid point of field k and the
    columns of y (T x m, real) -- tools/array.py:76-88 (np.corrcoef + Beta distribution).
    The S x T x m product and the column moments run on the device."""
    import scipy.stats
    X = self._device_fields()[k]                                          # T x S' (real, centred field)
    T = X.shape[0]
    yc = np.ascontiguousarray(y - y.mean(axis=0), dtype=np.float64)
    num = D.to_host(D.matmul(X, D.to_device(yc), trans_a=True))            # sum_t x (y - ybar)  (S' x m)
    ones = D.to_device(np.ones((T, 1)))
    colsum = D.to_host(D.matmul(X, ones, trans_a=True))[:, 0]
    sxx = D.to_host(D.col_sumsq(X)) - colsum ** 2 / T
    syy = (yc ** 2).sum(axis=0)
    r = num / np.sqrt(np.outer(sxx, syy))
    r = np.clip(r, -1.0, 1.0)
    dist = scipy.stats.beta(T / 2 - 1, T / 2 - 1, loc=-1, scale=2)
    p = 2 * dist.cdf(-abs(r))
    return r, p


def _patterns(self, n, phase_shift, hetero):
    pcs = self._get_pcs(n=n, phase_shift=phase_shift)
    reverse = dict(zip(self._keys, self._keys[::-1]))
    rvals, pvals = {}, {}
    for k in self._keys:
        src = reverse[k] if hetero else k
        if hetero and len(self._keys) < 2:
            raise KeyError("Key not found. Two fields needed for heterogenous maps.")
        r, p = _pearson(self, k, np.real(pcs[src]))
        rvals[k], pvals[k] = _scatter_space(self, k, r), _scatter_space(self, k, p)
    return rvals, pvals


def homogeneous_patterns(self, n=None, phase_shift=0):
    """array.py:1188-1221."""
    return _patterns(self, n, phase_shift, hetero=False)


def heterogeneous_patterns(self, n=None, phase_shift=0):
    """array.py:1223-1261."""
    return _patterns(self, n, phase_shift, hetero=True)


# ------------------------------------------------------------------ bootstrapping
def _resample(X, axis, block_size, replace):
    """Device version of tools/array.py:91-138 (moving-block bootstrap / permutation of a 2-D array).
    The block indices come from the GLOBAL numpy RNG exactly like the reference's np.random.choice."""
    if axis not in (0, 1):
        raise ValueError("{:} not a valid axis. either 0 or 1.".format(axis))
    arr = X if axis == 0 else D.transpose(X)
    n_obs = arr.shape[0]
    if n_obs % block_size != 0:
        raise ValueError("Length of data array ({:}) must be a multiple of block size {:}".format(n_obs, block_size))
    n_samples = n_obs // block_size
    idx = np.random.choice(n_samples, size=n_samples, replace=replace)
    rows = (idx[:, None] * block_size + np.arange(block_size)[None, :]).reshape(-1)
    out = D.gather_rows(arr, D.to_device(np.ascontiguousarray(rows, dtype=np.int64)))
    return out if axis == 0 else D.transpose(out)


def bootstrapping(self, n_runs, n_modes=20, axis=0, on_left=True, on_right=False, block_size=1, replace=True,
                  strategy="standard", disable_progress=False):
    """Monte-Carlo bootstrapping (array.py:1813-1952): every run resamples the (already resampled --
    the reference overwrites `X_surr` in place) fields, re-centres, solves [and rotates] on the GPU and
    records the variances.  Index draws replay the reference's use of the global numpy RNG."""
    from .rule_n import variance_of_fields
    self._require_solved("singular values")
    t = D.torch()
    complexify = self._analysis["is_complex"]
    extend, period = self._analysis["extend"], self._analysis["theta_period"]
    if extend == "theta":
        raise NotImplementedError("the Theta-model extension needs statsmodels")
    is_rotated = self._analysis["is_rotated"]
    n_rot, power = self._analysis["n_rot"], self._analysis["power"]
    n_modes_max = int(min(self._analysis["rank"], n_modes, n_rot))
    var_surr = np.zeros([n_modes_max, n_runs])
    dev = self._device_fields()
    for mode in range(n_modes):
        X = {k: dev[k].clone() for k in self._keys}                        # _get_X(original_scale=False, real=True)
        if strategy == "iterative":                                        # remove the first `mode` modes (array.py:1896-1901)
            V = self._get_V(n=mode, rotated=True)
            U = self._get_pcs(n=mode, scaling="eigen", rotated=True)
            for k in self._keys:
                if U[k].shape[1]:
                    _reconstruct_into(self, k, U[k], V[k], out=X[k], alpha=-1.0)
        for run in range(n_runs):
            if on_left and not on_right:
                X["left"] = _resample(X["left"], axis, block_size, replace)
            elif on_right and not on_left:
                if "right" not in X:
                    raise ValueError("No bootstrapping possible. There is no right field. Set `on_right=False`.")
                X["right"] = _resample(X["right"], axis, block_size, replace)
            elif on_left and on_right:
                s_left = X["left"].shape[1]
                cat = t.cat(list(X.values()), dim=1).contiguous()
                cat = _resample(cat, axis, block_size, replace)
                X["left"] = cat[:, :s_left].contiguous()
                if "right" in X:
                    X["right"] = cat[:, s_left:].contiguous()
            fields = []
            for k in self._keys:                                           # MCA(*X_surr): the constructor re-centres
                F = X[k].clone()
                D.center_columns(F)
                fields.append(F)
            var = variance_of_fields(fields, complexify, is_rotated, n_rot, power, extend, period)
            if var is None:
                continue                                                   # rotation did not converge (array.py:1939-1943)
            m = n_modes_max - mode
            var_surr[mode:, run] = var[:m]
        if strategy == "standard":
            break
    return var_surr
