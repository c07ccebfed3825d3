#!/usr/bin/env bash
# oracle/ref_build/build.sh — compiles the UNMODIFIED reference (ACDSLab/MPPI-Generic) GPU path from the sources where they
# lie under /root/reference into oracle/_ref/libmppi_ref_gpu.so: VanillaMPPIController + the reference's own kernels for the
# Cartpole and Autorally pairs, behind the C harness oracle/ref_build/ref_gpu.cu. "Reference kernels, shimmed host":
#   * Eigen (a system dependency of the reference that this image does not have) is replaced by oracle/ref_build/shim/Eigen,
#   * DDP feedback by a no-op feedback controller (the hot path never calls it),
#   * shim/ref_prefix.h restores M_PIf32 as a float literal (glibc + GCC 13 spell it as a _Float32 literal nvcc rejects),
#   * cnpy (the reference's vendored submodule) is compiled from /root/reference/submodules/cnpy/cnpy.cpp.
# The reference's own build system (cmake, googletest, yaml-cpp downloads) is not run. Nothing is copied out of
# /root/reference; outputs go to oracle/_ref/ only (git-ignored, shipped to the GPU box by gpurun).
#   REF=/root/reference bash oracle/ref_build/build.sh
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REF:-/root/reference}"
OUT="$HERE/../_ref"
CUDA_HOME="${CUDA_HOME:-/usr/local/cuda}"
[ -d "$REF/include/mppi" ] || { echo "reference tree not found at $REF (the GPU box uses the prebuilt oracle/_ref)"; exit 0; }
mkdir -p "$OUT"
# compute_100 (not 100a): the reference is architecture-generic CUDA; this is what its own CMake would emit for a B200
"$CUDA_HOME/bin/nvcc" -std=c++17 -O3 -g -lineinfo -gencode arch=compute_100,code=sm_100 -Xcompiler -fPIC -shared \
  -include "$HERE/shim/ref_prefix.h" -I"$HERE/shim" -I"$REF/include" -I"$REF/submodules/cnpy" \
  "$HERE/ref_gpu.cu" "$REF/submodules/cnpy/cnpy.cpp" -o "$OUT/libmppi_ref_gpu.so" \
  -L"$CUDA_HOME/lib64" -Xlinker -rpath -Xlinker "$CUDA_HOME/lib64" -lcurand -lcufft -lz
echo "built $OUT/libmppi_ref_gpu.so"
