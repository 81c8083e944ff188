"""``MCA`` -- drop-in for ``xmca.array.MCA`` on the solve / rotate / getters /
``rule_n`` hot path, computed by the B200 engine (``libxmca_b200.so``).

Host side (this file) mirrors the reference's constructor, metadata and NaN
handling (xmca/array.py:39-143, :145-240) and its getter semantics
(:605-779, :898-1063); every matrix product, SVD and rotation iteration runs
on the GPU through the C ABI.  There is no CPU fallback: without the shared
library or a CUDA device ``solve`` raises.

The callers either side of the path (``predict``, ``reconstructed_fields``, homogeneous /
heterogeneous patterns, ``bootstrapping``: SURVEY.md section 8f) live in ``downstream.py``.
Checkpointing (info file, ``load_analysis``) is in ``storage.py``.  Out of scope: plotting and the
Theta-model Hilbert extension (``NotImplementedError``).
"""
from __future__ import annotations

import cmath
import os
import warnings

import numpy as np
import yaml

from . import __version__
from . import _lib as L
from . import device as D
from . import engine as E
from . import downstream as DS
from . import storage as ST

_SIDES = ("left", "right")


class MCA:
    """Maximum Covariance Analysis / PCA of one or two ``numpy.ndarray`` fields
    (time on axis 0).  Same call surface as ``xmca.array.MCA`` (array.py:30)."""

    # ------------------------------------------------------------------ ctor
    def __init__(self, *fields):
        if len(fields) > 2:
            raise ValueError("Too many fields. Pass 1 or 2 fields.")
        if len(fields) == 2 and fields[0].shape[0] != fields[1].shape[0]:
            raise ValueError("Time dimensions of given fields are different. "
                             "Time series should have same time lengths.")
        if not all(isinstance(f, np.ndarray) for f in fields):
            raise TypeError("One or more fields are not `numpy.ndarray`. "
                            "Please provide `numpy.ndarray` only.")
        self._keys = list(_SIDES[:len(fields)])
        self._host = {}           # host mirror of the centred fields (filled lazily, see `_fields`)
        self._dev = {}            # device copies of the centred REAL fields (always real, see `solve`)
        self._devY = {}           # their Hilbert transforms (complex models, filled lazily)
        self._shape = {}
        self._field_names = {}
        self._field_means = {}
        self._field_stds = {}
        self._fields_spatial_shape = {}
        self._n_variables = {}
        self._no_nan_index = {}
        self._n_observations = {}
        on_device = L.cuda_available()
        for k, f in zip(self._keys, fields):
            self._shape[k] = f.shape
            self._n_observations[k] = f.shape[0]
            self._fields_spatial_shape[k] = f.shape[1:]
            self._n_variables[k] = int(np.prod(f.shape[1:]))      # incl. NaN columns (array.py:196)
            self._field_names[k] = k
            flat = f.reshape(f.shape[0], self._n_variables[k])
            if not np.issubdtype(flat.dtype, np.floating) or flat.dtype.itemsize < 4:
                flat = flat.astype(np.float64)
            if on_device:
                self._ingest_device(k, flat)
            else:
                self._ingest_host(k, flat)

        self._analysis = {
            "version": __version__,
            "is_bivariate": len(self._keys) > 1,
            "is_normalized": False,
            "is_coslat_corrected": False,
            "method": "pca",
            "is_complex": False,
            "extend": False,
            "theta_period": 365,
            "is_rotated": False,
            "n_rot": 0,
            "power": 0,
            "is_truncated": False,
            "is_truncated_at": 0,
            "rank": 0,
            "total_covariance": 0.0,
            "total_squared_covariance": 0.0,
        }
        self._analysis["method"] = self._get_method_id()
        self._dV = None           # device singular vectors
        self._solve_info = {}

    # ---------------------------------------------------------------- ingest
    _NAN_STEP_MSG = ("One or more fields contain NaN time steps. "
                     "Please remove these prior to analysis.")

    def _ingest_device(self, k, flat):
        """Constructor pre-processing of array.py:191-240 on the GPU: upload the raw field
        once, NaN scan + column mean/std in one kernel, then drop the NaN columns and centre
        in a second one.  Only the O(S) statistics come back to the host."""
        raw = D.to_device(flat)
        mean, std, col_nan, row_ok = D.field_stats(raw)
        if not row_ok.all():      # a NaN time step = a row that is NaN everywhere (tools/array.py:65-73)
            raise ValueError(self._NAN_STEP_MSG)
        keep = ~col_nan
        self._no_nan_index[k] = keep
        self._field_means[k] = mean[keep].astype(flat.dtype)
        self._field_stds[k] = std[keep].astype(flat.dtype)
        self._dev[k] = D.compact_center(raw, np.flatnonzero(keep), mean)

    def _ingest_host(self, k, flat):
        """The same bookkeeping in numpy, used only on a machine WITHOUT a CUDA device so that
        the class (validation, metadata, mode slicing) can still be constructed there; every
        numerical method (`solve`, `rotate`, getters, `rule_n`) raises without the GPU."""
        if np.isnan(flat).all(axis=1).any():
            raise ValueError(self._NAN_STEP_MSG)
        keep = ~np.isnan(flat).any(axis=0)
        self._no_nan_index[k] = keep
        flat = flat[:, keep]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", category=RuntimeWarning)
            self._field_means[k] = flat.mean(axis=0)
            self._field_stds[k] = flat.std(axis=0)
            self._host[k] = flat - flat.mean(axis=0)          # dtype preserved (array.py:199-207)

    @property
    def _fields(self):
        """Centred fields on the host, keyed 'left'/'right' (array.py `_fields`).  The device
        copy is the primary one; the host mirror is downloaded on first use."""
        for k in self._keys:
            if k not in self._host:
                x = D.to_host(self._dev[k])
                if self._analysis["is_complex"]:          # analytic field x + i H x (array.py:464)
                    x = (x + 1j * D.to_host(self._hilbert_dev(k))).astype(self._field_dtype(k))
                self._host[k] = x
        return {k: self._host[k] for k in self._keys}

    @_fields.setter
    def _fields(self, value):
        self._host = dict(value)
        self._dev = {}
        self._devY = {}

    def _hilbert_dev(self, k):
        """Device Hilbert transform H x of the centred real field k (imaginary part of the
        analytic signal, array.py:464): one GEMM with the circulant operator, cached."""
        if k not in self._devY:
            X = self._device_fields()[k]
            if self._analysis["extend"] == "exp":             # array.py:378-411, :455-472
                H = D.hilbert_matrix_exp_extension(X.shape[0], float(self._analysis["theta_period"]), X.dtype)
                Y, _ = D.apply_time_operator(H, X)
                D.center_columns(Y)                           # remove_mean of the cropped signal (array.py:471)
                self._devY[k] = Y
            else:
                H = D.hilbert_matrix(X.shape[0], X.dtype)
                self._devY[k], _ = D.apply_time_operator(H, X)
        return self._devY[k]

    def _n_kept(self, k):
        return int(self._no_nan_index[k].sum())

    def _field_dtype(self, k):
        """dtype of `_fields[k]` without materialising the host mirror."""
        real = self._field_means[k].dtype
        if k in self._dev:
            real = np.dtype(np.float32 if self._dev[k].dtype == D.f32() else np.float64)
        elif k in self._host:
            real = np.zeros(0, self._host[k].dtype).real.dtype
        if self._analysis["is_complex"]:
            return np.dtype(np.complex64 if real == np.float32 else np.complex128)
        return real

    # --------------------------------------------------------------- helpers
    def _get_method_id(self):
        return "mca" if self._analysis["is_bivariate"] else "pca"

    def set_field_names(self, left="left", right="right"):
        self._field_names["left"] = left
        self._field_names["right"] = right

    def _get_slice(self, n):
        """int -> modes [0, n); slice(a, b) -> 1-based inclusive (array.py:145-173)."""
        rank = self._analysis["rank"]
        if n is None:
            return slice(0, rank)
        if np.issubdtype(type(n), np.integer):
            return slice(0, n)
        if isinstance(n, slice):
            lo = max(0, n.start - 1) if isinstance(n.start, (int, np.integer)) else 0
            hi = min(n.stop, rank) if isinstance(n.stop, (int, np.integer)) else rank
            return slice(lo, hi, n.step)
        raise ValueError("Invalid type {:}. Must be either int or slice.".format(type(n)))

    def _scale_columns_device(self, scales, promote=True):
        """fields[k] *= scales[k] (per-column factors, broadcast over time) on the device copies."""
        dev = self._device_fields()
        for k, sc in scales.items():
            if k in dev and sc is not None:
                cs = D.to_device(np.array(np.broadcast_to(np.asarray(sc, dtype=np.float64), (dev[k].shape[1],))))
                # numpy promotion of `field * weights` (an fp64 weight ARRAY promotes an fp32 field)
                up = promote and isinstance(sc, np.ndarray) and sc.dtype == np.float64 and sc.ndim > 0
                dev[k] = D.scale_copy(dev[k], col_scale=cs, out_dtype=D.f64() if up else None)
        self._host, self._devY = {}, {}               # host mirror / Hilbert transforms are rebuilt lazily

    def apply_weights(self, left=None, right=None):
        """array.py:317-349: multiply the fields by weights.  Scalars and per-grid-point weights run on
        the device; anything else (weights varying in time) takes numpy broadcasting on the host mirror."""
        w = {"left": left, "right": right}
        simple = all(v is None or np.ndim(v) == 0 or (np.ndim(v) <= 2 and np.size(v) == self._n_kept(k)
                                                      and np.shape(v)[-1] == self._n_kept(k))
                     for k, v in w.items() if k in self._keys)
        if simple and L.cuda_available() and not self._analysis["is_complex"]:
            self._scale_columns_device({k: (None if v is None else np.reshape(v, -1) if np.ndim(v) else v)
                                        for k, v in w.items()})
            return
        w = {k: 1 if v is None else v for k, v in w.items()}
        self._fields = {k: f * w[k] for k, f in self._fields.items()}

    def normalize(self):
        """array.py:351-365: divide every grid point by its standard deviation."""
        if L.cuda_available() and not self._analysis["is_complex"]:
            self._scale_columns_device({k: 1.0 / self._field_stds[k].astype(np.float64) for k in self._keys},
                                       promote=False)       # std has the field dtype: no promotion
        else:
            self._fields = {k: f / self._field_stds[k] for k, f in self._fields.items()}
        self._analysis["is_normalized"] = True
        self._analysis["is_coslat_corrected"] = False
        self._analysis["method"] = self._get_method_id()

    def _get_X(self, original_scale=False, real=False):
        X = {k: f.copy() for k, f in self._fields.items()}
        if real:
            X = {k: x.real for k, x in X.items()}
        if original_scale:
            for k in X:
                if self._analysis["is_normalized"]:
                    X[k] *= self._field_stds[k]
                X[k] += self._field_means[k]
        return X

    def fields(self, original_scale=False):
        n_obs = self._n_observations["left"]
        out = {}
        for k, X in self._get_X(original_scale=original_scale).items():
            full = np.zeros([n_obs, self._n_variables[k]], dtype=X.dtype) * np.nan
            full[:, self._no_nan_index[k]] = X
            out[k] = full.reshape((n_obs,) + self._fields_spatial_shape[k])
        return out

    # ------------------------------------------------------------ device I/O
    def _device_fields(self):
        """Centred REAL fields on the device (uploaded once).  A complex model keeps the real
        field here; its imaginary part is `_hilbert_dev` (the reference re-takes `.real` before
        every Hilbert transform as well, array.py:456)."""
        for k in self._keys:
            if k not in self._dev:
                self._dev[k] = D.to_device(np.ascontiguousarray(np.real(self._host[k])))
        return self._dev

    # ----------------------------------------------------------------- solve
    def solve(self, complexify=False, extend=False, period=1):
        """Solve the MCA/PCA problem on the GPU (semantics of array.py:509-603)."""
        if len(self._keys) == 0 or any(self._n_kept(k) == 0 for k in self._keys):
            raise RuntimeError("Fields are empty. Did you forget to load data?")
        L._torch()                # no CUDA device -> XmcaLibraryError: there is no CPU fallback
        if extend == "theta":
            raise NotImplementedError("the Theta-model extension needs statsmodels (per-column model fits, "
                                      "array.py:367-376); use extend='exp' or extend=False")
        if extend not in (False, None, "exp"):
            raise ValueError("{:} is not a valid extension. Choose either `exp` or `theta`.".format(extend))
        if bool(complexify) != bool(self._analysis["is_complex"]) or extend != self._analysis["extend"] \
                or period != self._analysis["theta_period"]:
            self._devY = {}       # cached Hilbert transforms belong to the previous setting
            self._host = {k: v for k, v in self._host.items() if not np.iscomplexobj(v)}
        self._analysis["is_complex"] = bool(complexify)
        self._analysis["extend"] = extend
        self._analysis["theta_period"] = period

        dev = self._device_fields()
        if self._host and any(np.iscomplexobj(f) for f in self._host.values()) != bool(complexify):
            self._host = {}       # host mirror no longer matches the model (rebuilt lazily from the device)
        A = dev["left"]
        B = dev.get("right")
        real_dtype = np.float32 if A.dtype == D.f32() else np.float64
        try:
            if complexify and extend == "exp":
                sigma, vec, res = E.solve_complex_time(A, self._hilbert_dev("left"), B,
                                                       self._hilbert_dev("right") if B is not None else None)
                self._dV = ("complex", vec)
            elif complexify:
                sigma, vec, res = E.solve_complex(A, B)
                self._dV = ("complex", vec)
            else:
                res = E.solve_real(A, B)
                sigma = res.sigma
                self._dV = ("real", res)
        except np.linalg.LinAlgError:
            raise np.linalg.LinAlgError("SVD failed. NaN entries may be the problem.")
        self._solve_info = {"route": res.route, "sweeps": res.sweeps}
        self._Vhost = {}

        sv = sigma.astype(real_dtype)                      # singular values keep the field dtype
        self._singular_values = sv
        self._variance = sv
        self._var_idx = np.argsort(sv)[::-1]
        self._norm = {k: np.sqrt(sv) for k in self._keys}
        n = len(sv)
        self._analysis["total_covariance"] = sv.sum()
        self._analysis["total_squared_covariance"] = (sv ** 2).sum()
        self._analysis["rank"] = n
        self._analysis["is_rotated"] = False
        self._analysis["n_rot"] = n
        self._analysis["power"] = 0
        # array.py:601-602 store eye(rank) twice (2 x 512 MB at rank 8192); kept lazy here, the
        # getters below materialise the identity on demand
        self._rot_R = None
        self._rot_Phi = None
        self._analysis["is_truncated_at"] = n

    # --------------------------------------------------------------- getters
    def _require_solved(self, what):
        if not hasattr(self, "_singular_values"):
            raise RuntimeError("Cannot retrieve {}. Please call the method `solve` first.".format(what))

    def _get_svals(self, n=None):
        self._require_solved("singular values")
        return self._singular_values[self._get_slice(n)]

    def _complex_dtype(self):
        return np.complex64 if self._singular_values.dtype == np.float32 else np.complex128

    def _V_device_cols(self, k, m):
        """First m unrotated singular vectors of field k as device tensors:
        real -> (S x m); complex -> ((S x m) re, (S x m) im)."""
        _, provider = self._dV
        m = min(int(m), self._singular_values.size)
        return provider.vectors(m)[k]       # singular vectors are computed on demand (engine.TridiagResult)

    def _V_host(self, k, m):
        """Unrotated V[k][:, :m] on the host (downloaded once, cached)."""
        cached = self._Vhost.get(k)
        if cached is None or cached.shape[1] < m:
            kind, _ = self._dV
            v = self._V_device_cols(k, m)
            if kind == "real":
                part = D.to_host(v.contiguous())
            else:
                part = (D.to_host(v[0].contiguous())
                        + 1j * D.to_host(v[1].contiguous())).astype(self._complex_dtype())
            self._Vhost[k] = cached = part
        return cached[:, :m]

    @property
    def _V(self):
        self._require_solved("singular vectors")
        return {k: self._V_host(k, self._singular_values.size) for k in self._keys}

    def _get_V(self, n=None, rotated=True):
        self._require_solved("singular vectors")
        if rotated:
            top = self._analysis["n_rot"]
        else:
            top = n.stop if isinstance(n, slice) else n
        keep = self._get_slice(n)
        nsv = self._singular_values.size
        top = nsv if top is None else min(top, nsv)
        out = {}
        for k in self._keys:
            if rotated and self._analysis["is_rotated"]:
                v = self._rotated_loadings_host(k)[:, self._var_idx][:, keep]   # S x n_rot, / norm, fp64
            elif rotated:
                # unrotated model: R = I (array.py:636-642 multiplies by eye(rank)); only the
                # reordering and the promotion to fp64 remain -- download just the columns asked for
                cols = self._var_idx[:top][keep]
                need = int(cols.max()) + 1 if cols.size else 0
                v = self._V_host(k, need)[:, cols]
                v = v.astype(np.complex128 if np.iscomplexobj(v) else np.float64)
            else:
                v = self._V_host(k, top)[:, keep]
            out[k] = v
        return out

    def _rotated_loadings_host(self, k):
        """V sqrt(sigma) R / norm for field k (array.py:634-640), from the device
        product L_rot computed in rotate() (downloaded on first use)."""
        if k not in self._rot_eofs:
            self._rot_eofs[k] = D.to_host(self._rot_eofs_dev[k])
        return self._rot_eofs[k]

    def _no_nan_index_dev(self, k):
        cache = self.__dict__.setdefault("_no_nan_dev", {})
        if k not in cache:
            cache[k] = D.to_device(np.ascontiguousarray(self._no_nan_index[k]))
        return cache[k]

    def _get_U(self, n=None, rotated=True):
        self._require_solved("principal components")
        if rotated:
            top = self._analysis["n_rot"]
        else:
            top = n.stop if isinstance(n, slice) else n
        nsv = self._singular_values.size
        top = nsv if top is None else min(top, nsv)
        keep = self._get_slice(n)
        is_rot = rotated and self._analysis["is_rotated"]
        if is_rot:
            cols, need = None, top
        else:
            # unrotated: R = I (array.py:668-669 multiplies by eye(rank)); project only onto the
            # modes that are returned instead of onto all `rank` of them
            order = self._var_idx[:top] if rotated else np.arange(top)
            cols = order[keep]
            need = int(cols.max()) + 1 if cols.size else 0
        dev = self._device_fields()
        root = np.sqrt(self._singular_values[:need].astype(np.float64))
        inv_root = D.to_device(1.0 / root) if need else None
        Rit = self.rotation_matrix(inverse_transpose=True) if is_rot else None
        kind, _ = self._dV
        out = {}
        for k in self._keys:
            X = dev[k]
            if need == 0:
                T = self._n_observations[k]
                out[k] = np.zeros((T, 0), dtype=np.float64 if rotated else self._field_dtype(k))
                continue
            if kind == "real":
                Vd = self._V_device_cols(k, need)
                Ud = D.matmul(X, Vd)                                        # T x need (array.py:667)
                Ud = D.scale_copy(Ud, col_scale=inv_root)
                if is_rot:
                    Ud = D.matmul(Ud, D.to_device(np.ascontiguousarray(Rit)))
                u = D.to_host(Ud)
            else:
                # Z V = (X + iY)(Vr + iVi): two skinny products over the real field and its Hilbert transform
                vr, vi = self._V_device_cols(k, need)
                t = D.torch()
                Vri = t.cat([vr, vi], dim=1).contiguous()                   # S x 2 need
                sc2 = t.cat([inv_root, inv_root])
                P1 = D.to_host(D.scale_copy(D.matmul(X, Vri), col_scale=sc2))
                P2 = D.to_host(D.scale_copy(D.matmul(self._hilbert_dev(k), Vri), col_scale=sc2))
                u = (P1[:, :need] - P2[:, need:]) + 1j * (P1[:, need:] + P2[:, :need])
                if is_rot:
                    u = u @ Rit
            if is_rot:
                u = u[:, self._var_idx][:, keep]
            else:
                u = u[:, cols]
                if not rotated:
                    u = u.astype(self._field_dtype(k))
            out[k] = u
        return out

    def _get_norm(self, n=None, sorted=True):
        self._require_solved("field norms")
        norm = self._norm
        if sorted:
            norm = {k: v[self._var_idx] for k, v in norm.items()}
        sl = self._get_slice(n)
        return {k: v[sl] for k, v in norm.items()}

    def _get_variance(self, n=None, sorted=True):
        nrm = self._get_norm(n=n, sorted=sorted)
        if self._analysis["is_bivariate"]:
            return nrm["left"] * nrm["right"]
        return nrm["left"] ** 2

    @staticmethod
    def _apply_scaling(arr, scaling, norm_k, axes):
        if scaling == "None":
            return arr
        if scaling == "eigen":
            return arr * norm_k
        if scaling == "max":
            return arr / np.nanmax(abs(arr.real), axis=axes)
        if scaling == "std":
            return arr / np.nanstd(arr.real, axis=axes)
        raise ValueError("The scaling option {:} is not valid. Please choose one of the "
                         "following: None, eigen, std, max".format(scaling))

    def _get_eofs(self, n=None, scaling="None", phase_shift=0, rotated=True):
        if rotated and self._analysis["is_rotated"] and not self._analysis["is_complex"] \
                and all(k in getattr(self, "_rot_eofs_dev", {}) for k in self._keys):
            return self._get_eofs_rotated_dev(n, scaling)
        V = self._get_V(n, rotated=rotated)
        out = {}
        for k in self._keys:
            nm = V[k].shape[1]
            full = np.zeros([self._n_variables[k], nm], dtype=V[k].dtype) * np.nan
            full[self._no_nan_index[k], :] = V[k]
            full = full.reshape(self._fields_spatial_shape[k] + (nm,))
            if self._analysis["is_complex"]:
                full = full * cmath.rect(1, phase_shift)
            norm_k = self._get_norm(V["left"].shape[1], sorted=True)[k] if scaling == "eigen" else None
            out[k] = self._apply_scaling(full, scaling, norm_k, tuple(range(full.ndim - 1)))
        return out

    def _get_eofs_rotated_dev(self, n, scaling):
        """Rotated real model: the same result as the general path (`_get_V` -> mode order -> full grid with NaN at
        the removed points, array.py:634-642 / 1245-1262), with the reordering and the scatter done on the device and
        ONE download per field instead of three host-side copies of the S x n_rot array."""
        t = D.torch()
        cols = np.ascontiguousarray(self._var_idx[self._get_slice(n)])
        idx = D.to_device(cols.astype(np.int64))
        out = {}
        for k in self._keys:
            v = self._rot_eofs_dev[k].index_select(1, idx)
            nm = int(v.shape[1])
            mask = self._no_nan_index[k]
            if bool(np.all(mask)):
                full = v
            else:
                full = t.full((self._n_variables[k], nm), float("nan"), dtype=v.dtype, device=v.device)
                full[self._no_nan_index_dev(k)] = v
            arr = D.to_host(full).reshape(self._fields_spatial_shape[k] + (nm,))
            norm_k = self._get_norm(nm, sorted=True)[k] if scaling == "eigen" else None
            out[k] = self._apply_scaling(arr, scaling, norm_k, tuple(range(arr.ndim - 1)))
        return out

    def _get_pcs(self, n=None, scaling="None", phase_shift=0, rotated=True):
        U = self._get_U(n, rotated=rotated)
        out = {}
        for k in self._keys:
            u = U[k]
            if self._analysis["is_complex"]:
                u = u * cmath.rect(1, phase_shift)
            norm_k = self._get_norm(n, sorted=True)[k] if scaling == "eigen" else None
            out[k] = self._apply_scaling(u, scaling, norm_k, 0)
        return out

    # --------------------------------------------------------------- rotate
    def rotate(self, n_rot, power=1, tol=1e-8):
        """Varimax / Promax rotation of the first ``n_rot`` loaded EOFs on the
        GPU (semantics of array.py:781-844, tools/rotation.py)."""
        if n_rot < 2:
            raise ValueError("`n_rot` must be > 1")
        if power < 1:
            raise ValueError("`power` must be >=1")
        self._require_solved("singular values")
        sv = self._get_svals(n_rot)
        n_rot = sv.size
        if self._analysis["is_complex"]:
            return self._rotate_complex(sv, n_rot, power, tol)
        root = D.to_device(np.sqrt(sv.astype(np.float64)))
        t = D.torch()
        parts = [D.scale_copy(self._V_device_cols(k, n_rot), col_scale=root) for k in self._keys]
        s_left = parts[0].shape[0]
        Ld = t.cat(parts, dim=0).contiguous() if len(parts) > 1 else parts[0]     # (S1'+S2') x n_rot
        try:
            Lrot, R, Phi, iters = E.promax(Ld, power, max_iter=1000, tol=tol)
        except L.NotConvergedError:
            raise RuntimeError("Rotation process did not converge. Try decreasing the tolerance. "
                               "Invalid NaN entries also might be a problem.")
        n_all = Lrot.shape[0]
        nl = np.sqrt(D.to_host(D.col_sumsq(Lrot, 0, s_left)))                      # array.py:826-830
        nr = np.sqrt(D.to_host(D.col_sumsq(Lrot, s_left, n_all))) if self._analysis["is_bivariate"] else nl
        self._norm = {"left": nl, "right": nr} if self._analysis["is_bivariate"] else {"left": nl}
        self._variance = nl * nr
        self._var_idx = np.argsort(self._variance)[::-1]
        self._rot_R = R
        self._rot_Phi = Phi
        self._analysis["is_rotated"] = True
        self._analysis["n_rot"] = n_rot
        self._analysis["power"] = power
        self._solve_info["varimax_iterations"] = iters
        if n_rot <= E.VARIMAX_FUSED_MAX_P and D.last_varimax_stats is not None:
            st = D.to_host(D.last_varimax_stats)
            self._solve_info["varimax_svd_sweeps"] = int(st[3])
            self._solve_info["varimax_phase_clocks"] = [float(x) for x in st[4:10]]
        # rotated EOFs = L_rot / norm (array.py:640): one scaled copy per field; it stays on the device (`eofs` reorders
        # it, re-inserts the NaN grid points there and downloads once), a host copy is made only when asked for
        self._rot_eofs = {}
        self._rot_eofs_dev = {}
        bounds = {"left": (0, s_left), "right": (s_left, n_all)}
        for k in self._keys:
            lo, hi = bounds[k]
            inv = D.to_device(1.0 / self._norm[k])
            self._rot_eofs_dev[k] = D.scale_copy(Lrot[lo:hi], col_scale=inv)

    def _rotate_complex(self, sv, n_rot, power, tol):
        """Complex model: fused complex Varimax kernel on the planar loadings."""
        _, provider = self._dV
        V = provider.vectors(n_rot)
        try:
            Br, Bi, s_left, R, Phi, iters = E.rotate_complex(V, sv.astype(np.float64), self._keys, n_rot, power,
                                                             tol=tol)
        except L.NotConvergedError:
            raise RuntimeError("Rotation process did not converge. Try decreasing the tolerance. "
                               "Invalid NaN entries also might be a problem.")
        n_all = Br.shape[0]
        nl = E.complex_col_norms(Br, Bi, 0, s_left)
        nr = E.complex_col_norms(Br, Bi, s_left, n_all) if self._analysis["is_bivariate"] else nl
        self._norm = {"left": nl, "right": nr} if self._analysis["is_bivariate"] else {"left": nl}
        self._variance = nl * nr
        self._var_idx = np.argsort(self._variance)[::-1]
        self._rot_R = R
        self._rot_Phi = Phi
        self._analysis["is_rotated"] = True
        self._analysis["n_rot"] = n_rot
        self._analysis["power"] = power
        self._solve_info["varimax_iterations"] = iters
        self._rot_eofs = {}
        self._rot_eofs_dev = {}
        bounds = {"left": (0, s_left), "right": (s_left, n_all)}
        for k in self._keys:
            lo, hi = bounds[k]
            self._rot_eofs[k] = (D.to_host(Br[lo:hi]) + 1j * D.to_host(Bi[lo:hi])) / self._norm[k]

    @property
    def _rotation_matrix(self):
        self._require_solved("rotation matrix")
        return self._rot_R if self._rot_R is not None else np.eye(self._singular_values.size)

    @property
    def _correlation_matrix(self):
        self._require_solved("correlation matrix")
        return self._rot_Phi if self._rot_Phi is not None else np.eye(self._singular_values.size)

    def rotation_matrix(self, inverse_transpose=False):
        try:
            R = self._rotation_matrix
        except AttributeError:
            R = np.eye(len(self.singular_values()))
        if inverse_transpose and self._analysis["power"] > 1:
            R = np.linalg.pinv(R).conjugate().T
        return R

    def correlation_matrix(self):
        try:
            idx = self._var_idx
            return self._correlation_matrix[idx, :][:, idx]
        except AttributeError:
            return np.eye(len(self.singular_values()))

    # ----------------------------------------------------------- public API
    def singular_values(self, n=None):
        return self._get_svals(n)

    def norm(self, n=None, sorted=True):
        return self._get_norm(n=n, sorted=sorted)

    def variance(self, n=None, sorted=True):
        return self._get_variance(n=n, sorted=sorted)

    def scf(self, n=None):
        self._require_solved("squared covariance fraction")
        var = self._variance[self._var_idx][:n]                  # NB: plain [:n] (array.py:982)
        return var ** 2 / self._analysis["total_squared_covariance"] * 100

    def explained_variance(self, n=None):
        return self._get_variance(n=n, sorted=True) / self._analysis["total_covariance"] * 100

    def pcs(self, n=None, scaling="None", phase_shift=0, rotated=True):
        return self._get_pcs(n, scaling, phase_shift, rotated)

    def eofs(self, n=None, scaling="None", phase_shift=0, rotated=True):
        return self._get_eofs(n, scaling, phase_shift, rotated)

    def spatial_amplitude(self, n=None, scaling="None", rotated=True):
        e = self.eofs(n, scaling="None", rotated=rotated)
        amp = {k: np.sqrt(v * v.conjugate()).real for k, v in e.items()}
        if scaling == "max":
            amp = {k: a / np.nanmax(a, axis=tuple(range(a.ndim - 1))) for k, a in amp.items()}
        return amp

    def spatial_phase(self, n=None, phase_shift=0, rotated=True):
        e = self.eofs(n, phase_shift=phase_shift, rotated=rotated)
        return {k: np.arctan2(v.imag, v.real).real for k, v in e.items()}

    def temporal_amplitude(self, n=None, scaling="None", rotated=True):
        p = self.pcs(n, scaling="None", rotated=rotated)
        amp = {k: np.sqrt(v * v.conjugate()).real for k, v in p.items()}
        if scaling == "max":
            amp = {k: a / np.nanmax(a, axis=0) for k, a in amp.items()}
        return amp

    def temporal_phase(self, n=None, phase_shift=0, rotated=True):
        p = self.pcs(n, phase_shift=phase_shift, rotated=rotated)
        return {k: np.arctan2(v.imag, v.real).real for k, v in p.items()}

    def truncate(self, n):
        """array.py:1602-1627."""
        if self._analysis["is_rotated"] and n < self._analysis["n_rot"]:
            raise ValueError("Cannot truncte rotated solution. Please ensure `n` > `n_rot`")
        if n < self._singular_values.size:
            self._singular_values = self._singular_values[:n]     # also caps every vector request
            self._Vhost = {k: v[:, :n] for k, v in self._Vhost.items()}
            self._analysis["is_truncated"] = True
            self._analysis["is_truncated_at"] = n

    def rule_north(self, n=None):
        """array.py:1773-1811."""
        sv = self._get_svals(n)
        err = sv * np.sqrt(2.0 / self._n_observations["left"])
        if self._analysis["is_complex"]:
            err = err * np.sqrt(2)
        return err

    def summary(self):
        """array.py:2014-2024."""
        print(yaml.dump({k: (v.item() if hasattr(v, "item") else v) for k, v in self._analysis.items()},
                        sort_keys=False))

    # --------------------------------------------------------------- rule N
    def rule_n(self, n_runs, n_modes=None, seed=None, group=None, surrogate_dtype=None):
        """Rule N (Overland & Preisendorfer 1982), semantics of array.py:1716-1771,
        with the surrogate loop sharded over the ranks of ``group``
        (``torch.distributed``; ``None`` = single GPU).  See ``xmca_b200.rule_n``."""
        from . import rule_n as RN
        return RN.rule_n(self, n_runs, n_modes=n_modes, seed=seed, group=group, surrogate_dtype=surrogate_dtype)

    # ------------------------------------------- callers downstream of the hot path
    def _scale_X(self, data_dict):
        return DS.scale_X(self, data_dict)

    def _scale_X_inverse(self, data_dict):
        return DS.scale_X_inverse(self, data_dict)

    def predict(self, left=None, right=None, n=None, scaling="None", phase_shift=0):
        return DS.predict(self, left, right, n, scaling, phase_shift)

    def _reconstructed_X(self, mode=None, original_scale=True):
        return DS.reconstructed_X(self, mode, original_scale)

    def reconstructed_fields(self, mode=None, original_scale=True):
        return DS.reconstructed_fields(self, mode, original_scale)

    def homogeneous_patterns(self, n=None, phase_shift=0):
        return DS.homogeneous_patterns(self, n, phase_shift)

    def heterogeneous_patterns(self, n=None, phase_shift=0):
        return DS.heterogeneous_patterns(self, n, phase_shift)

    def bootstrapping(self, n_runs, n_modes=20, axis=0, on_left=True, on_right=False, block_size=1,
                      replace=True, strategy="standard", disable_progress=False):
        return DS.bootstrapping(self, n_runs, n_modes, axis, on_left, on_right, block_size, replace, strategy,
                                disable_progress)

    # ---------------------------------------------------------- out of scope
    def _out_of_scope(self, name):
        raise NotImplementedError("`{}` is outside the solve/rotate/rule_n hot path this engine "
                                  "accelerates (SURVEY.md section 8f).".format(name))

    def plot(self, *a, **k):
        self._out_of_scope("plot")

    def save_plot(self, *a, **k):
        self._out_of_scope("save_plot")

    # ------------------------------------------------ checkpoint (info file + arrays)
    def _get_analysis_path(self, path=None):
        return ST.analysis_path(self, path)

    def _create_analysis_path(self, path):
        os.makedirs(self._get_analysis_path(path), exist_ok=True)

    def _get_file_names(self, format):
        return ST.file_names(self, format)

    def _create_info_file(self, path):
        ST.create_info_file(self, path)

    def _set_info_from_file(self, path):
        ST.set_info_from_file(self, path)

    def _save_data(self, data_array, path, *args, **kwargs):
        raise NotImplementedError("only works for `xarray`")           # array.py:1687-1688

    def load_analysis(self, path, fields=None, eofs=None, singular_values=None):
        ST.load_analysis(self, path, fields, eofs, singular_values)
