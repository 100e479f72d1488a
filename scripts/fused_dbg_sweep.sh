# ablation sweep of the fused update+evaluate kernel (OLFNAV_FUSED_DBG bits, see fused_step.cu)
for d in ${DBG_LIST:-0 1 8 7 9 11 13 29 45 77 125 15}; do echo "dbg=$d"; OLFNAV_FUSED_DBG=$d timeout 100 python scripts/fused_bench.py C2 100000 75 2>&1 | grep -E '"random"|"all-N"' | sed 's/.*"pattern": \("[^"]*"\).*"fused_ms": \([0-9.]*\).*/  \1 fused_ms=\2/'; done
