'ORACLE SHIM — empty stand-in for h5py (absent here); in-memory ndarray data never touches it.'


def File(*args, **kwargs):
    raise NotImplementedError('h5py is not available in this container (oracle shim)')
