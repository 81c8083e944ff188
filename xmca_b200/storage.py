"""Checkpoint format of the reference: the plain-text ``info.xmca`` file
(xmca/array.py:1629-1714) and ``load_analysis`` from fields + unrotated EOFs + singular values
(array.py:1954-2012).  The NetCDF side (xmca/xarray.py:1239-1314) needs xarray + h5netcdf and is
delegated to the DataArray backend when it offers ``to_netcdf`` / ``open_dataarray``.
Bound to ``MCA`` / ``xMCA`` in array.py / xarray.py.
"""
from __future__ import annotations

import os
import textwrap
from datetime import datetime

import numpy as np

from . import device as D


def secure_str(string):
    """tools/text.py:15-16."""
    return string.lower().replace(" ", "_")


def analysis_path(self, path=None):
    """array.py:248-256."""
    if path is None:
        path = os.path.join(os.getcwd(), "xmca", secure_str("_".join(self._field_names.values())))
    elif not os.path.isabs(path):
        path = os.path.abspath(path)
    return path


def file_names(self, fmt):
    """array.py:1661-1685."""
    fields, eofs = {}, {}
    for key, variable in self._field_names.items():
        variable = secure_str(variable)
        fields[key] = ".".join([variable, fmt])
        eofs[key] = ".".join(["_".join([variable, "eofs"]), fmt])
    return {"fields": fields, "eofs": eofs, "pcs": {}, "singular": ".".join(["singular_values", fmt]), "norm": {}}


def create_info_file(self, path):
    """Write ``info.xmca`` exactly in the reference's layout (array.py:1629-1659)."""
    sep_line = "\n#" + "-" * 79
    now = datetime.now().strftime("%Y-%m-%d %H:%M:%S")
    header = ("This file contains information neccessary to load stored analysis"
              "data from xmca module.")
    with open(os.path.join(path, "info.xmca"), "w+") as fh:
        fh.write(textwrap.indent(textwrap.fill(header, width=80), "# "))
        fh.write("\n# To load this analysis use:")
        fh.write("\n# from xmca.xarray import xMCA")
        fh.write("\n# mca = xMCA()")
        fh.write("\n# mca.load_analysis(PATH_TO_THIS_FILE)")
        fh.write("\n")
        fh.write(sep_line)
        fh.write(sep_line)
        fh.write("\n{:<20} : {:<57}".format("created", now))
        fh.write(sep_line)
        for key, name in self._field_names.items():
            fh.write("\n{:<20} : {:<57}".format(key, str(name)))
        fh.write(sep_line)
        for key, info in self._analysis.items():
            if key in ["is_bivariate", "is_complex", "is_rotated", "is_truncated"]:
                fh.write(sep_line)
            fh.write("\n{:<20} : {:<57}".format(key, str(info)))


def set_info_from_file(self, path):
    """Parse ``info.xmca`` (array.py:1690-1714): values are cast to the type of the default."""
    with open(path, "r") as fh:
        for line in fh.readlines():
            if line[0] == "#":
                continue
            key = line.split(":")[0].rstrip()
            if key in ["left", "right"]:
                self._field_names[key] = line.split(":")[1].strip()
            if key in self._analysis:
                value = line.split(":")[1].strip()
                kind = type(self._analysis[key])
                if isinstance(self._analysis[key], (bool, np.bool_)):
                    self._analysis[key] = (value == "True")
                elif isinstance(self._analysis[key], (np.floating, float)):
                    self._analysis[key] = float(value)
                elif isinstance(self._analysis[key], (np.integer, int)):
                    self._analysis[key] = int(value)
                else:
                    self._analysis[key] = kind(value)


class StoredVectors:
    """Vector provider of a LOADED model: the unrotated singular vectors come from the checkpoint
    (host arrays, S' x modes) instead of a solve; uploaded on demand."""

    route = "loaded"
    sweeps = []

    def __init__(self, V):
        self.Vh = V
        self._dev = {}

    def vectors(self, m):
        out = {}
        for k, v in self.Vh.items():
            if k not in self._dev:
                if np.iscomplexobj(v):
                    self._dev[k] = (D.to_device(np.ascontiguousarray(v.real)), D.to_device(np.ascontiguousarray(v.imag)))
                else:
                    self._dev[k] = D.to_device(np.ascontiguousarray(v))
            d = self._dev[k]
            out[k] = (d[0][:, :m], d[1][:, :m]) if isinstance(d, tuple) else d[:, :m]
        return out


def load_analysis(self, path, fields=None, eofs=None, singular_values=None):
    """array.py:1954-2012: rebuild a model from `info.xmca`, the ORIGINAL-scale real fields, the unrotated
    EOFs and the singular values; re-centres, re-normalises, re-rotates like the reference."""
    set_info_from_file(self, path)
    self._keys = ["left", "right"] if self._analysis["is_bivariate"] else ["left"]
    is_complex, is_norm = self._analysis["is_complex"], self._analysis["is_normalized"]
    self._analysis["is_complex"] = False                 # fields are ingested as real data first
    self._host, self._dev, self._devY = {}, {}, {}
    from . import _lib as L
    for k in self._keys:
        f = np.asarray(fields[k])
        self._shape[k] = f.shape
        self._n_observations[k] = f.shape[0]
        self._fields_spatial_shape[k] = f.shape[1:]
        self._n_variables[k] = int(np.prod(f.shape[1:]))
        self._field_names.setdefault(k, k)
        flat = f.reshape(f.shape[0], self._n_variables[k])
        if not np.issubdtype(flat.dtype, np.floating) or flat.dtype.itemsize < 4:
            flat = flat.astype(np.float64)
        if L.cuda_available():
            self._ingest_device(k, flat)
        else:
            self._ingest_host(k, flat)
    if is_norm:
        self._analysis["is_normalized"] = False
        self.normalize()
    self._analysis["is_complex"] = is_complex            # `_fields` / PCs build the analytic signal lazily
    sv = np.asarray(singular_values)
    self._singular_values = sv
    self._variance = sv
    self._var_idx = np.argsort(sv)[::-1]
    self._norm = {k: np.sqrt(sv) for k in self._keys}
    V = {}
    for k in self._keys:
        e = np.asarray(eofs[k])
        e = e.reshape(self._n_variables[k], e.shape[-1])
        V[k] = e[self._no_nan_index[k], :]               # remove_nan_cols(eofs.T).T
    self._dV = ("complex" if is_complex else "real", StoredVectors(V))
    self._Vhost = {}
    self._solve_info = {"route": "loaded", "sweeps": []}
    self._rot_R = None
    self._rot_Phi = None
    if self._analysis["is_rotated"]:
        self.rotate(self._analysis["n_rot"], self._analysis["power"])
