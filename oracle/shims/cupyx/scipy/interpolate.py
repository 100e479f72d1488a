'ORACLE SHIM — cupyx.scipy.interpolate stand-in (forwards to SciPy).'
from scipy.interpolate import RegularGridInterpolator  # noqa: F401
