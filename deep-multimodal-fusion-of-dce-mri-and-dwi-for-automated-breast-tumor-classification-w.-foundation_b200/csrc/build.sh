#!/usr/bin/env bash
# Builds libb200fusion.so (sm_100a only) next to the Python modules.  nvcc cross-compiles
# without a GPU; the .so is git-ignored but travels with the tree to the GPU box.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
OUT="$HERE/../libb200fusion.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I"$ROOT/include" -I"$HERE")
mkdir -p "$HERE/build"
objs=()
for f in "$HERE"/*.cu; do
  o="$HERE/build/$(basename "${f%.cu}").o"
  if [[ ! -f "$o" || "$f" -nt "$o" || "$ROOT/include/b200_fusion.h" -nt "$o" || "$HERE/ptx.cuh" -nt "$o" || "$HERE/common.cuh" -nt "$o" ]]; then
    "$NVCC" "${FLAGS[@]}" ${PTXAS_V:+-Xptxas -v} -c "$f" -o "$o" &
  fi
  objs+=("$o")
done
wait
"$NVCC" -shared -o "$OUT" "${objs[@]}" -lcudart
echo "built $OUT"
