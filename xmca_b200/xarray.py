"""``xMCA`` -- drop-in for ``xmca.xarray.xMCA`` (xmca/xarray.py:23): the xarray facade over
``xmca_b200.MCA``.  It only strips ``.values`` from the DataArrays on the way in and attaches
``mode`` / ``time`` / ``lat`` / ``lon`` coordinates and the ``_analysis`` attributes on the way
out (xarray.py:85, :286-299, :455-469, :495-514, :1477-1488); every number comes from the
B200 engine through the parent class.

xarray is imported lazily (it is not part of the build image); ``set_backend`` lets the tests
inject a duck-typed stand-in exposing ``DataArray(data, dims=, coords=, name=, attrs=)``.
"""
from __future__ import annotations

import numpy as np

from .array import MCA

_backend = None


def set_backend(module):
    """Use `module.DataArray` instead of importing xarray (tests / minimal installs)."""
    global _backend
    _backend = module


def _xr():
    global _backend
    if _backend is None:
        try:
            import xarray
        except ImportError as err:            # pragma: no cover - depends on the environment
            raise ImportError("xmca_b200.xarray.xMCA needs the `xarray` package (or a backend "
                              "registered with xmca_b200.xarray.set_backend)") from err
        _backend = xarray
    return _backend


class xMCA(MCA):
    """Maximum Covariance Analysis / PCA of one or two ``xarray.DataArray`` fields with
    dimensions (time, lat, lon).  Same call surface as ``xmca.xarray.xMCA``."""

    def __init__(self, *fields):
        xr = _xr()
        if len(fields) > 2:
            raise ValueError("Too many fields. Pass 1 or 2 fields.")
        if not all(isinstance(f, xr.DataArray) for f in fields):
            raise TypeError("One or more fields are not `xarray.DataArray`. "
                            "Please provide `xarray.DataArray` only.")
        keys = ["left", "right"]
        self._field_dims = {keys[i]: f.dims for i, f in enumerate(fields)}
        self._field_coords = {keys[i]: f.coords for i, f in enumerate(fields)}
        super().__init__(*[np.asarray(f.values) for f in fields])

    # ------------------------------------------------------------ helpers
    def _attrs(self):
        return {k: str(v) for k, v in self._analysis.items()}

    def _modes(self, n, length):
        sl = self._get_slice(n)
        return list(range(sl.start + 1, sl.stop + 1))[:length]

    def _mode_array(self, values, n, name):
        return _xr().DataArray(values, dims=["mode"], coords={"mode": self._modes(n, len(values))},
                               name=name, attrs=self._attrs())

    # ------------------------------------------------------ pre-processing
    def apply_weights(self, **weights):
        """xarray.py:136-165: multiply the (full-grid) fields by broadcastable weights."""
        fields = super().fields()
        new = dict(self._fields)
        for k, weight in weights.items():
            if k not in fields:
                raise KeyError("Key `{:}` not found. Please use `left` or `right`".format(k))
            w = np.asarray(getattr(weight, "values", weight))
            try:
                dims = getattr(weight, "dims", None)
                if dims is not None and len(dims) == 1 and dims[0] in self._field_dims[k]:
                    shape = [1] * fields[k].ndim
                    shape[self._field_dims[k].index(dims[0])] = w.size
                    w = w.reshape(shape)
                new_field = (fields[k] * w).reshape(self._n_observations[k], self._n_variables[k])
                new[k] = new_field[:, self._no_nan_index[k]]
            except ValueError as err:
                raise ValueError("Error for {:} weights. Mismatch between dimensions of weights ({:}) "
                                 "and original field ({:}).".format(k, w.shape, fields[k].shape)) from err
        self._fields = new

    def apply_coslat(self):
        """Area weighting sqrt(cos(lat)) (xarray.py:167-181)."""
        weights = {}
        for k, coord in self._field_coords.items():
            lat = np.asarray(getattr(coord["lat"], "values", coord["lat"]), dtype=np.float64)
            w = np.sqrt(np.cos(np.deg2rad(lat)) + 1e-6)
            shape = [1] * len(self._field_dims[k])
            shape[list(self._field_dims[k]).index("lat")] = w.size
            weights[k] = w.reshape(shape)
        self.apply_weights(**weights)
        self._analysis["is_coslat_corrected"] = True

    # -------------------------------------------------------------- getters
    def fields(self, original_scale=False):
        xr = _xr()
        out = super().fields(original_scale)
        return {k: xr.DataArray(out[k], dims=self._field_dims[k], coords=self._field_coords[k],
                                name=self._field_names[k]) for k in self._keys}

    def singular_values(self, n=None):
        return self._mode_array(super().singular_values(n), n, "singular values")

    def norm(self, n=None, sorted=True):
        norms = super().norm(n=n, sorted=sorted)
        return {k: self._mode_array(v, n, " ".join([self._field_names[k], "norm"])) for k, v in norms.items()}

    def variance(self, n=None, sorted=True):
        return self._mode_array(super().variance(n=n, sorted=sorted), n, "variance")

    def explained_variance(self, n=None):
        return self._mode_array(super().explained_variance(n), n, "covariance fraction")

    def scf(self, n=None):
        return self._mode_array(super().scf(n), n, "squared covariance fraction")

    def _wrap_time(self, data, n, suffix):
        xr = _xr()
        return {k: xr.DataArray(v, dims=["time", "mode"],
                                coords={"time": self._field_coords[k]["time"], "mode": self._modes(n, v.shape[-1])},
                                name=" ".join([self._field_names[k], suffix]), attrs=self._attrs())
                for k, v in data.items()}

    def _wrap_space(self, data, n, suffix):
        xr = _xr()
        return {k: xr.DataArray(v, dims=["lat", "lon", "mode"],
                                coords={"lon": self._field_coords[k]["lon"], "lat": self._field_coords[k]["lat"],
                                        "mode": self._modes(n, v.shape[-1])},
                                name=" ".join([self._field_names[k], suffix]), attrs=self._attrs())
                for k, v in data.items()}

    def pcs(self, n=None, scaling="None", phase_shift=0, rotated=True):
        return self._wrap_time(super().pcs(n, scaling, phase_shift, rotated), n, "pcs")

    def eofs(self, n=None, scaling="None", phase_shift=0, rotated=True):
        return self._wrap_space(super().eofs(n, scaling, phase_shift, rotated), n, "eofs")

    def spatial_amplitude(self, n=None, scaling="None", rotated=True):
        return self._wrap_space(super().spatial_amplitude(n, scaling, rotated), n, "spatial amplitude")

    def spatial_phase(self, n=None, phase_shift=0, rotated=True):
        return self._wrap_space(super().spatial_phase(n, phase_shift, rotated), n, "spatial phase")

    def temporal_amplitude(self, n=None, scaling="None", rotated=True):
        return self._wrap_time(super().temporal_amplitude(n, scaling, rotated), n, "temporal amplitude")

    def temporal_phase(self, n=None, phase_shift=0, rotated=True):
        return self._wrap_time(super().temporal_phase(n, phase_shift, rotated), n, "temporal phase")

    # --------------------------------------------------------- significance
    def rule_north(self, n=None):
        return self._mode_array(super().rule_north(n), n, "error singular values")

    def rule_n(self, n_runs, n_modes=None, seed=None, group=None, surrogate_dtype=None):
        """xarray.py:1447-1488: (mode, run) DataArray of the surrogate spectra."""
        sv = super().rule_n(n_runs, n_modes, seed=seed, group=group, surrogate_dtype=surrogate_dtype)
        return _xr().DataArray(sv, dims=["mode", "run"],
                               coords={"mode": self._modes(n_modes, sv.shape[0]),
                                       "run": np.arange(1, sv.shape[1] + 1)},
                               name="singular values")

    # ----------------------------------------- callers downstream of the hot path
    def _wrap_patterns(self, pats, pvals, n, suffix):
        return (self._wrap_space(pats, n, suffix), self._wrap_space(pvals, n, "pvalues " + suffix))

    def homogeneous_patterns(self, n=None, phase_shift=0):
        """xarray.py:690-745."""
        r, p = super().homogeneous_patterns(n=n, phase_shift=phase_shift)
        return self._wrap_patterns(r, p, n, "homogeneous patterns")

    def heterogeneous_patterns(self, n=None, phase_shift=0):
        """xarray.py:747-802."""
        r, p = super().heterogeneous_patterns(n=n, phase_shift=phase_shift)
        return self._wrap_patterns(r, p, n, "heterogeneous patterns")

    def reconstructed_fields(self, mode=slice(1, None), original_scale=True):
        """xarray.py:804-833."""
        xr = _xr()
        rec = super().reconstructed_fields(mode=mode, original_scale=original_scale)
        return {k: xr.DataArray(rec[k], dims=self._field_dims[k], coords=self._field_coords[k],
                                name="reconstructed_{:}_field".format(k)) for k in self._keys}

    def predict(self, left=None, right=None, n=None, scaling="None", phase_shift=0):
        """xarray.py:835-892: DataArrays in, (time, mode) DataArrays out."""
        xr = _xr()
        data = dict(zip(self._keys, [left, right]))
        try:
            values = {k: (d if d is None else d.values) for k, d in data.items()}
        except AttributeError as err:
            raise ValueError("Please provide `xr.DataArray` to `left` and `right`") from err
        pcs_new = super().predict(values.get("left"), values.get("right") if self._analysis["is_bivariate"] else None,
                                  n, scaling, phase_shift)
        return {k: xr.DataArray(pc, dims=("time", "mode"),
                                coords={"time": data[k].coords["time"], "mode": list(range(1, pc.shape[1] + 1))})
                for k, pc in pcs_new.items()}

    def bootstrapping(self, n_runs, n_modes=20, axis=0, on_left=True, on_right=False, block_size=1, replace=True,
                      strategy="standard", disable_progress=False):
        """xarray.py:1357-1439 (the reference forwards axis=0 whatever is passed, :1419)."""
        sv = super().bootstrapping(n_runs=n_runs, n_modes=n_modes, axis=0, on_left=on_left, on_right=on_right,
                                   block_size=block_size, replace=replace, strategy=strategy,
                                   disable_progress=disable_progress)
        return _xr().DataArray(sv, dims=["mode", "run"],
                               coords={"mode": self._modes(n_modes, len(sv)), "run": list(range(1, sv.shape[1] + 1))},
                               name="singular values", attrs=self._attrs())

    # ------------------------------------------------------------ checkpoint (NetCDF)
    def _save_data(self, data, path, engine="h5netcdf", *args, **kwargs):
        """xarray.py:1239-1251 -- needs a DataArray backend with `to_netcdf` (real xarray)."""
        import os
        from .storage import secure_str
        if not hasattr(data, "to_netcdf"):
            raise NotImplementedError("saving needs xarray (+ h5netcdf); the active DataArray backend cannot write NetCDF")
        out = os.path.join(path, secure_str(".".join([data.name, "nc"])))
        data.to_netcdf(path=out, engine=engine, invalid_netcdf=(engine == "h5netcdf"), *args, **kwargs)

    def save_analysis(self, path=None, engine="h5netcdf"):
        """xarray.py:1253-1279: info.xmca + original-scale real fields, UNROTATED EOFs, singular values."""
        path = self._get_analysis_path(path)
        self._create_analysis_path(path)
        self._create_info_file(path)
        fields = self.fields(original_scale=True)
        eofs = self.eofs(rotated=False)
        self._save_data(self.singular_values(), path, engine)
        for k in self._keys:
            self._save_data(eofs[k], path, engine)
            f = fields[k]
            f.values = np.real(f.values)
            self._save_data(f, path, engine)

    def load_analysis(self, path, engine="h5netcdf"):
        """xarray.py:1281-1314."""
        import os
        xr = _xr()
        if not hasattr(xr, "open_dataarray"):
            raise NotImplementedError("loading needs xarray (+ h5netcdf)")
        self._set_info_from_file(path)
        folder, _ = os.path.split(path)
        names = self._get_file_names(format="nc")
        sv = xr.open_dataarray(os.path.join(folder, names["singular"]), engine=engine).data
        fields, eofs = {}, {}
        for k in self._field_names:
            eofs[k] = xr.open_dataarray(os.path.join(folder, names["eofs"][k]), engine=engine).data
            f = xr.open_dataarray(os.path.join(folder, names["fields"][k]), engine=engine)
            self._field_coords[k], self._field_dims[k] = f.coords, f.dims
            fields[k] = f.data
        MCA.load_analysis(self, path=path, fields=fields, eofs=eofs, singular_values=sv)
        if self._analysis["is_coslat_corrected"]:
            self.apply_coslat()
