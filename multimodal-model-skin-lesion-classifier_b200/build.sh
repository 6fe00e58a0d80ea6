#!/usr/bin/env bash
# Builds lib/libfb200.so for sm_100a (cross-compiles without a GPU).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
mkdir -p "$HERE/lib" "$HERE/build"
FLAGS=(${FB200_EXTRA_DEFS:-} -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall --expt-relaxed-constexpr)
for f in plan exec tabt; do
  "$NVCC" "${FLAGS[@]}" ${FB200_PTXAS_V:+-Xptxas -v} -c "$HERE/csrc/$f.cu" -o "$HERE/build/$f.o" &
done
wait
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$HERE/lib/libfb200.so" "$HERE/build/plan.o" "$HERE/build/exec.o" "$HERE/build/tabt.o" -lcudart_static -ldl -lrt -lpthread
echo "built $HERE/lib/libfb200.so"
