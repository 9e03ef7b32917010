#!/usr/bin/env bash
# TEST INFRASTRUCTURE. Runs the reference's OWN test files (unmodified, from /root/reference)
# under the NumPy stand-in for jax (oracle/jax_shim). Only works in the build container
# (/root/reference does not exist on the GPU box). This is the gate that pins the stand-in,
# which in turn pins oracle/cfd_oracle.py through tests/golden (see oracle/gen_golden.py).
set -u
here="$(cd "$(dirname "$0")" && pwd)"
export PYTHONPATH="$here/jax_shim:/root/reference"
cd /tmp
run() { echo "== $*"; python -m pytest -q -x --no-header -p no:cacheprovider "$@" 2>&1 | tail -3; }
B=/root/reference/jax_cfd/base
run $B/boundaries_test.py
run $B/grids_test.py
run $B/finite_differences_test.py
run $B/fast_diagonalization_test.py
run $B/forcings_test.py
run $B/funcutils_test.py
run $B/interpolation_test.py -k "not point_interpolation and not PointInterpolation"
run $B/pressure_test.py -k "fast_diag or poisson or Poisson"
run $B/advection_test.py -k "(using_limiters or equivalence_1d) and not gradients"
run $B/subgrid_models_test.py -k "smagorinsky_viscosity or evm_model"
run $B/equations_test.py -k "fast_diag"
