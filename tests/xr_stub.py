"""Duck-typed stand-in for the parts of xarray that xMCA touches (xarray is not installed in
the build image).  Registered with ``xmca_b200.xarray.set_backend``."""
import numpy as np


class DataArray:
    def __init__(self, data, dims=None, coords=None, name=None, attrs=None):
        self.values = np.asarray(data)
        self.dims = tuple(dims) if dims is not None else tuple("dim_%d" % i for i in range(self.values.ndim))
        self.coords = {k: (v if isinstance(v, DataArray) else np.asarray(v)) for k, v in dict(coords or {}).items()}
        self.name = name
        self.attrs = dict(attrs or {})

    @property
    def shape(self):
        return self.values.shape
