#!/usr/bin/env bash
# Builds the example out-of-tree pair against the engine's internal header (same source revision as libmppi_b200.so).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
CUDA_HOME="${CUDA_HOME:-/usr/local/cuda}"
"$CUDA_HOME/bin/nvcc" -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -shared \
  -o "$HERE/libmppi_plugin_pendulum.so" "$HERE/pendulum_pair.cu" -I"$HERE/../include" \
  -L"$HERE/../mppi-generic_b200" -Xlinker -rpath -Xlinker "$HERE/../mppi-generic_b200" -l:libmppi_b200.so \
  -L"$CUDA_HOME/lib64" -lcurand -lcufft
echo "built $HERE/libmppi_plugin_pendulum.so"
