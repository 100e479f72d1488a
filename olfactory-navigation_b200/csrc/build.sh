#!/usr/bin/env bash
# Builds libolfnav_b200.so (sm_100a only) next to the Python package.  Usage: build.sh [extra nvcc flags]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libolfnav_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a \
    -Xcompiler -fPIC -shared \
    "$@" \
    "$HERE/model.cu" "$HERE/belief.cu" "$HERE/evaluate.cu" "$HERE/solver.cu" "$HERE/infotaxis.cu" "$HERE/umma_gemm.cu" "$HERE/sim.cu" "$HERE/fused_step.cu" "$HERE/dense.cu" \
    -o "$OUT"
echo "built $OUT"
