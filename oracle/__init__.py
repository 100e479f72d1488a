'''
ORACLE — TEST INFRASTRUCTURE ONLY (parity unpinned; see oracle/pomdp_toolkit/__init__.py).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import anything below this directory.  The product package never does.
'''
