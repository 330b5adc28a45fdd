#!/usr/bin/env bash
# Build libgeomae_b200.so in-tree for sm_100a (B200).  nvcc cross-compiles without a GPU.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="$here/../libgeomae_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
flags=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr
       -Xcompiler -fPIC,-O3,-Wall,-Wno-unused-function -shared -cudart static ${GEOMAE_NVCC_EXTRA:-})
objs=()
mkdir -p "$here/../../build/obj"
pids=()
for src in "$here"/*.cu; do
  obj="$here/../../build/obj/$(basename "${src%.cu}").o"
  objs+=("$obj")
  if [[ ! -f "$obj" || "$src" -nt "$obj" || -n "$(find "$here" "$here/../../include" \( -name '*.cuh' -o -name '*.h' \) -newer "$obj" -print -quit)" ]]; then
    "$NVCC" "${flags[@]/-shared/-c}" -o "$obj" "$src" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [[ -n "$p" ]] && wait "$p"; done
"$NVCC" "${flags[@]}" -o "$out" "${objs[@]}"
echo "built $out"
