#!/bin/bash
# Builds libmppi_b200_controllers.so: the explicit controller instantiations of src/controllers/*/ (the reference's
# src/controllers/*/CMakeLists.txt libraries cartpole_mppi, double_integrator_mppi, quadrotor_mppi, autorally_mppi in one file).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")/../.." && pwd)"
OUT="$HERE/mppi-generic_b200/libmppi_b200_controllers.so"
g++ -std=c++17 -O2 -fPIC -shared -I"$HERE/include" -I/usr/local/cuda/include \
  "$HERE"/src/controllers/*/*_mppi.cpp -o "$OUT" \
  -L"$HERE/mppi-generic_b200" -l:libmppi_b200.so -Wl,-rpath,"$HERE/mppi-generic_b200" -Wl,-rpath,'$ORIGIN'
echo "built $OUT"
