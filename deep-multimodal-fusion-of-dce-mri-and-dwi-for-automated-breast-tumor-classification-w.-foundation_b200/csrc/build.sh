#!/usr/bin/env bash
# Builds libb200fusion.so (sm_100a only).  nvcc cross-compiles without a GPU; the .so is git-ignored but
# travels with the tree to the GPU box.  Output: <repo>/lib/libb200fusion.so (a short, plain path: that is the
# path handed to dlopen).
#   build.sh            incremental: recompiles the translation units whose sources / headers changed
#   build.sh --clean    drops every object first (what __graft_entry__.build() runs)
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
OUT="$ROOT/lib/libb200fusion.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I"$ROOT/include" -I"$HERE" ${B200_EXTRA_NVCC_FLAGS:-})
if [[ "${1:-}" == "--clean" ]]; then rm -rf "$HERE/build"; fi
mkdir -p "$HERE/build" "$ROOT/lib"
objs=()
pids=()
names=()
for f in "$HERE"/*.cu; do
  o="$HERE/build/$(basename "${f%.cu}").o"
  stale=0
  if [[ ! -f "$o" || "$f" -nt "$o" || "$ROOT/include/b200_fusion.h" -nt "$o" ]]; then stale=1; fi
  for h in "$HERE"/*.cuh; do [[ "$h" -nt "$o" ]] && stale=1; done
  if [[ $stale == 1 ]]; then
    rm -f "$o"   # a failed compile must break the link, never fall back to an older object
    "$NVCC" "${FLAGS[@]}" ${PTXAS_V:+-Xptxas -v} -c "$f" -o "$o" &
    pids+=($!)
    names+=("$(basename "$f")")
  fi
  objs+=("$o")
done
fail=0
for i in "${!pids[@]}"; do
  if ! wait "${pids[$i]}"; then echo "build.sh: nvcc failed on ${names[$i]}" >&2; fail=1; fi
done
[[ $fail == 0 ]] || exit 1
echo "compiled: ${names[*]:-nothing (all objects up to date)}"
"$NVCC" -shared -o "$OUT" "${objs[@]}" -lcudart
echo "built $OUT"
