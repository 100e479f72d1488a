'ORACLE SHIM — matplotlib sub-module stand-in: names used in annotations only.'


class _Placeholder:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError('matplotlib is not available (oracle shim)')


Axes = Rectangle = Circle = LineCollection = Normalize = ListedColormap = _Placeholder


def __getattr__(name):
    return _Placeholder
