"""xmca_b200 -- B200-native (sm_100a) engine for the solve / rotate / rule_n hot
path of Maximum Covariance Analysis, behind the ``xmca.array.MCA`` /
``xmca.xarray.xMCA`` class surface.  See DESIGN.md."""
__version__ = "0.1.0"

from .array import MCA  # noqa: E402,F401
from .xarray import xMCA  # noqa: E402,F401  (xarray itself is imported lazily)

__all__ = ["MCA", "xMCA", "__version__"]
