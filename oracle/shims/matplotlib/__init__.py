'ORACLE SHIM — stand-in for matplotlib (absent here); plotting is out of scope.'
