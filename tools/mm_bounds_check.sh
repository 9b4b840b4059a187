#!/bin/bash
# compute-sanitizer is closed on this pool: rebuild mm.cu with device-side bounds checks (every computed index of the
# Machado-Mata kernels is asserted; a violation traps with file:line and the CUDA call after it fails), run the
# Machado-Mata GPU tests on that build, then restore the production build.  Usage (on the GPU box): bash tools/mm_bounds_check.sh
set -u
cd "$(dirname "$0")/.."
touch oaxaca_blinder_rs_b200/csrc/mm.cu
make -C oaxaca_blinder_rs_b200/csrc -s EXTRA=-DOB_MM_BOUNDS_CHECK || exit 1
echo "== built with -DOB_MM_BOUNDS_CHECK"
python -m pytest tests/test_gpu_mm.py -q 2>&1 | tail -5
rc=${PIPESTATUS[0]}
touch oaxaca_blinder_rs_b200/csrc/mm.cu
make -C oaxaca_blinder_rs_b200/csrc -s || exit 1
echo "== production build restored"
exit $rc
