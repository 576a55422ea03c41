#!/bin/bash
# Builds libi3d_host.so: the reference-shaped C++ host API (nv::Optimizer, nv::NLSSolver, cost-term create(), nv::LightingSVSH, nv::SDFColorization::add/compute, nv::SDFAlgorithms, nv::Intrinsic3D orchestrator, file formats) on top of
# the C-ABI of libi3d_b200.so.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$HERE/../.."
/usr/bin/g++ -O2 -std=c++17 -fPIC -shared -Wall -Wextra -I"$ROOT/include" \
    -o "$HERE/../libi3d_host.so" "$HERE/optimizer.cpp" "$HERE/nls_solver.cpp" "$HERE/cost_terms.cpp" "$HERE/lighting_svsh.cpp" "$HERE/colorization.cpp" "$HERE/algorithms.cpp" "$HERE/io.cpp" "$HERE/intrinsic3d.cpp" "$HERE/c_wrapper.cpp" \
    -L"$HERE/.." -li3d_b200 -Wl,-rpath,'$ORIGIN'
echo "built $HERE/../libi3d_host.so"
