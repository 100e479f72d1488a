#!/bin/bash
# period of the full-completion wait of the TMA stores (0 = never): speed of the one-pass kernel and fault count in the ablation mode that faulted 6 times in 6 runs
for g in "$@"; do
  export OLFNAV_F2_STORE_SYNC=$g
  ms=$(python scripts/fused2_profile.py 100000 75 5 2>&1 | grep '"random"' | sed 's/.*"ms": \([0-9.]*\).*/\1/')
  f=0
  for k in 1 2 3; do OLFNAV_F2_DBG=46 timeout 100 python scripts/fused2_ablation_run.py 2>&1 | grep -q "exception" && f=$((f + 1)); done
  echo "store sync period $g: $ms ms (random actions), $f faults in 3 ablation runs"
done
