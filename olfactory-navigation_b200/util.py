'RNG helper with the reference semantics (reference olfactory_navigation/util.py:11-30).'
import numpy as np


def get_rng(rng=None) -> np.random.Generator:
    if rng is None:
        return np.random.default_rng()
    if isinstance(rng, (int, np.integer)):
        return np.random.default_rng(seed=int(rng))
    return rng
