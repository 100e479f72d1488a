'ORACLE SHIM — empty stand-in for cupyx (see ../cupy/__init__.py).'
