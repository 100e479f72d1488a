'''
ORACLE SHIM — test infrastructure only.  A NumPy pass-through standing in for CuPy so that the
UNMODIFIED reference Python (/root/reference) imports and runs on a CPU-only box.  Needed because
olfactory_navigation/simulation.py:139,146 dereference `cp` unconditionally and
olfactory_navigation/agents/infotaxis_agent.py:254 ends up with `xp = cp`.
Every attribute forwards to numpy.
'''
import numpy as _np


class ndarray:  # isinstance(x, cp.ndarray) must be False for NumPy arrays
    pass


def get_array_module(*args):
    return _np


def asnumpy(a, *args, **kwargs):
    return _np.asarray(a)


def __getattr__(name):
    return getattr(_np, name)
