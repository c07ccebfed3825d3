#!/bin/bash
# Builds an experimental variant of the library next to the shipped one: tools/libexp_<NAME>.so with -DMPPIB_EXP_<NAME>
# (plugins/dynamics.cuh, plugins/costs.cuh). Select it at run time with MPPIB_LIB=/root/repo/tools/libexp_<NAME>.so.
#   bash tools/build_exp.sh NO_COST [MORE_MACROS...]   (macros left in the tree: NO_COST, NO_TEX — ablations of the Autorally cost)
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
NAME=$1
DEFS=""
for m in "$@"; do DEFS="$DEFS -DMPPIB_EXP_$m"; done
cd "$HERE/mppi-generic_b200"
/usr/local/cuda/bin/nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wno-unused-function \
  -shared $DEFS -o "$HERE/tools/libexp_$NAME.so" csrc/engine.cu csrc/host_twins.cpp csrc/npz_reader.cpp -I../include \
  -L/usr/local/cuda/lib64 -Xlinker -rpath -Xlinker /usr/local/cuda/lib64 -lcurand -lcufft -ldl -lz
echo "built $HERE/tools/libexp_$NAME.so"
