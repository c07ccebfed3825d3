#!/usr/bin/env bash
# Builds libmppi_b200.so (sm_100a only) in-tree. Usage: ./build.sh [extra nvcc flags]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
CUDA_HOME="${CUDA_HOME:-/usr/local/cuda}"
OUT="$HERE/libmppi_b200.so"
"$CUDA_HOME/bin/nvcc" -std=c++17 -O3 -lineinfo \
  -gencode arch=compute_100a,code=sm_100a \
  -Xcompiler -fPIC,-Wall,-Wno-unused-function -shared \
  -Xptxas -v \
  "$@" \
  -o "$OUT" "$HERE/csrc/engine.cu" "$HERE/csrc/host_twins.cpp" "$HERE/csrc/npz_reader.cpp" \
  -I"$HERE/../include" \
  -L"$CUDA_HOME/lib64" -Xlinker -rpath -Xlinker "$CUDA_HOME/lib64" -lcurand -lcufft -ldl -lz
echo "built $OUT"
