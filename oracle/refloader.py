'''
ORACLE — TEST INFRASTRUCTURE ONLY.

Makes the UNMODIFIED reference Python importable in THIS container:
  * /root/reference on sys.path (package `olfactory_navigation`),
  * oracle/shims (cupy pass-through, h5py / matplotlib stubs),
  * oracle/ on sys.path so that `import pomdp_toolkit` resolves to the restated oracle.

/root/reference does not exist on the GPU box: callers must check `available()` and skip.
Used by oracle/make_golden.py and the in-container cross-checks in tests/ (not `-m gpu`).
'''
import os
import sys

REFERENCE_ROOT = '/root/reference'
_HERE = os.path.dirname(os.path.abspath(__file__))


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'olfactory_navigation'))


def install_oracle_toolkit() -> None:
    'Make `import pomdp_toolkit` resolve to oracle/pomdp_toolkit (no reference needed).'
    if _HERE not in sys.path:
        sys.path.insert(0, _HERE)


def load():
    'Returns the imported reference package `olfactory_navigation` (with .agents loaded).'
    if not available():
        raise RuntimeError('/root/reference is not present on this machine')
    install_oracle_toolkit()
    shims = os.path.join(_HERE, 'shims')
    for p in (shims, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(1, p)
    import olfactory_navigation
    import olfactory_navigation.agents  # noqa: F401
    return olfactory_navigation
