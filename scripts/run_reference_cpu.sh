#!/bin/bash
# Off-box reference harness: build the UNMODIFIED Fortran reference with
# gfortran + MPI on a machine that has them (this repository's image has neither), run its own executables on a
# case, and compare with libmagudi_gpu.  It is the only route from "parity unpinned" to parity pinned against the
# compiled reference.  Nothing here runs in the build image or on the GPU box.
#
#   scripts/run_reference_cpu.sh <path-to-magudi-checkout> <case-dir> [nprocs]
#
# <case-dir> holds magudi.inp, bc.dat and the PLOT3D grid / initial condition (e.g. a copy of
# examples/AcousticMonopole after `python config.py`).  Steps:
#   1. cmake Release build of the reference (its own CMakeLists; MPI Fortran compiler required)
#   2. mpirun -np N ./forward --output J0.txt          -> forward QoI J              (bin/Forward.f90)
#      mpirun -np N ./adjoint                          -> cost sensitivity, gradient (bin/Adjoint.f90)
#      mpirun -np N ./rhs <prefix>.ic.q                -> <prefix>.rhs.f             (utils/rhs.f90:129-201)
#   3. python scripts/compare_with_reference.py <case-dir>   (reads the PLOT3D files with magudi_b200.plot3d, runs
#      the same case through libmagudi_gpu and prints max relative differences of the RHS field, J and the gradient;
#      tolerances of BASELINE.json: 1e-12 on RHS fields, 1e-10 on J and gradient)
set -euo pipefail
REF=${1:?path to the magudi checkout}
CASE=${2:?case directory}
NP=${3:-$(nproc)}
command -v mpif90 >/dev/null || { echo "mpif90 not found: this harness needs gfortran + MPI" >&2; exit 2; }
command -v cmake >/dev/null || { echo "cmake not found" >&2; exit 2; }
BUILD=$REF/build-reference
mkdir -p "$BUILD"
( cd "$BUILD" && cmake -DCMAKE_BUILD_TYPE=Release -DCMAKE_Fortran_COMPILER=mpif90 .. && make -j"$NP" forward adjoint rhs )
cd "$CASE"
PREFIX=$(awk -F= '/output_prefix/ {gsub(/[ "\047]/, "", $2); print $2}' magudi.inp | head -1)
echo "cores: $NP   prefix: $PREFIX"
/usr/bin/time -v mpirun -np "$NP" "$BUILD/bin/forward" --output J0.txt 2> forward.time
/usr/bin/time -v mpirun -np "$NP" "$BUILD/bin/adjoint" 2> adjoint.time
mpirun -np "$NP" "$BUILD/bin/rhs" "$PREFIX.ic.q"
grep -E "Elapsed|Maximum resident" forward.time adjoint.time
echo "J (reference) = $(cat J0.txt)"
python "$(dirname "$0")/compare_with_reference.py" "$CASE"
