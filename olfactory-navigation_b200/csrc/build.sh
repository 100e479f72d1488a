#!/usr/bin/env bash
# Builds libolfnav_b200.so (sm_100a only) next to the Python package.  Usage: build.sh [extra nvcc flags]
# Every translation unit is compiled on its own (in parallel, objects cached under csrc/build/ by source time stamp), then linked.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libolfnav_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
OBJ="$HERE/build"
mkdir -p "$OBJ"
SRCS=(model belief evaluate solver infotaxis umma_gemm sim fused_step2 dense)
FLAGS=(-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC "$@")
STAMP="$OBJ/.flags"
if [[ ! -f "$STAMP" || "$(cat "$STAMP")" != "${FLAGS[*]}" ]]; then rm -f "$OBJ"/*.o; echo "${FLAGS[*]}" > "$STAMP"; fi
pids=()
for s in "${SRCS[@]}"; do
    src="$HERE/$s.cu"; obj="$OBJ/$s.o"
    newest="$src"
    for h in "$HERE"/*.cuh "$HERE/../../include/olfnav_b200.h"; do [[ "$h" -nt "$newest" ]] && newest="$h"; done
    if [[ ! -f "$obj" || "$newest" -nt "$obj" ]]; then
        "$NVCC" "${FLAGS[@]}" -c "$src" -o "$obj" &
        pids+=($!)
    fi
done
for p in "${pids[@]:-}"; do [[ -n "$p" ]] && wait "$p"; done
OBJS=()
for s in "${SRCS[@]}"; do OBJS+=("$OBJ/$s.o"); done
"$NVCC" -shared -gencode arch=compute_100a,code=sm_100a "${OBJS[@]}" -o "$OUT"
echo "built $OUT"
