"""Import the UNMODIFIED reference (``xmca`` package of nicrie/xmca) so the oracle can be pinned against it,
golden vectors generated, and -- in ``bench.py --impl reference`` / ``cpu_baseline`` -- the reference's own CPU
path timed on the GPU box's host cores.

TEST / BASELINE INFRASTRUCTURE ONLY; the product never imports it.  Two locations are tried:
``/root/reference`` (the build container) and ``oracle/_ref`` -- a verbatim, git-ignored copy of the reference's
``xmca/`` package directory that ``__graft_entry__.build()`` makes when ``/root/reference`` exists, so that it
travels to the GPU box with the snapshot like the built ``.so`` (it is never committed).

The three shims are the ones recorded in SURVEY.md Appendix A; they live here,
never in the reference tree.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = [os.environ.get("XMCA_REFERENCE_ROOT", "/root/reference"), os.path.join(_HERE, "_ref")]
REFERENCE_ROOT = next((c for c in _CANDIDATES if os.path.isdir(os.path.join(c, "xmca"))), _CANDIDATES[0])
FIXTURES = os.path.join(REFERENCE_ROOT, "tests", "integration", "fixtures")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "xmca"))


def reference_location() -> str:
    return REFERENCE_ROOT


def import_reference_mca():
    """Return the reference ``xmca.array.MCA`` class (numpy path)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if not hasattr(np, "product"):                      # numpy>=2 (array.py:196,235)
        np.product = np.prod
    if "matplotlib" not in sys.modules:                 # array.py:13, tools/text.py:8
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        plt.rcParams = {"text.usetex": False}
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    for name in ("statsmodels", "statsmodels.tsa", "statsmodels.tsa.forecasting",
                 "statsmodels.tsa.forecasting.theta"):  # array.py:17 (extend='theta' only)
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["statsmodels.tsa.forecasting.theta"].ThetaModel = None
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from xmca.array import MCA                          # noqa: E402
    return MCA


def import_reference_rotation():
    import_reference_mca()
    from xmca.tools.rotation import promax, varimax     # noqa: E402
    return varimax, promax


def _read(path, offset, shape, dtype):
    with open(os.path.join(FIXTURES, path), "rb") as fh:
        buf = fh.read()
    return np.frombuffer(buf, dtype, int(np.prod(shape)), offset).reshape(shape).copy()


def read_reference_fixtures() -> dict:
    """The reference's own golden vectors (HDF5 superblock v0, contiguous,
    little-endian; byte offsets from SURVEY.md section 8c)."""
    out = {
        "sst": _read("sst.nc", 12236, (492, 9, 18), "<f4"),
        "prcp": _read("prcp.nc", 12236, (492, 9, 18), "<f4"),
        "sv_std": _read("std/singular_values.nc", 9432, (155,), "<f4"),
        "sv_cplx": _read("cplx/singular_values.nc", 9432, (155,), "<f8"),
        "eofs_std_sst": _read("std/sst_eofs.nc", 10240, (9, 18, 155), "<f4"),
        "eofs_std_prcp": _read("std/prcp_eofs.nc", 10240, (9, 18, 155), "<f4"),
    }
    for name in ("sst", "prcp"):
        raw = _read("cplx/%s_eofs.nc" % name, 10240, (9, 18, 155, 2), "<f8")
        out["eofs_cplx_%s" % name] = raw[..., 0] + 1j * raw[..., 1]
    return out
