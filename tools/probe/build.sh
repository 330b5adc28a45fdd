#!/usr/bin/env bash
# Builds the scatter microbenchmark probe (tools only; not loaded by the product or the tests).
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr \
  -Xcompiler -fPIC,-O3 -shared -cudart static -o "$here/scatter_probe.so" "$here/scatter_probe.cu"
echo "built $here/scatter_probe.so"
