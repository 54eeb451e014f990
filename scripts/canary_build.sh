#!/bin/bash
# Shared-memory canary build of the library (compute-sanitizer is closed on this pool): every region of the active-set
# kernels' shared-memory carves is followed by guard words that the kernel fills when it starts and checks when it ends; a
# write past the end of a region prints the region and traps.  Run the GPU suite under it:
#   scripts/canary_build.sh && CMPC_LIB=$PWD/build/libcmpc_canary.so python -m pytest tests -m gpu -x -q
# -DCMPC_CANARY_SELFTEST re-introduces the round-1 first-tier defect (a 32-wide zeroing of a narrower row) to prove that the
# canaries see it:  scripts/canary_build.sh selftest  ->  build/libcmpc_canary_selftest.so
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
mkdir -p "$ROOT/build"
EXTRA=""; OUT="libcmpc_canary.so"
if [ "$1" = "selftest" ]; then EXTRA="-DCMPC_CANARY_SELFTEST"; OUT="libcmpc_canary_selftest.so"; fi
cd "$ROOT/quad-periodic-mpc_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -DCMPC_CANARY $EXTRA -o "$ROOT/build/$OUT" *.cu
echo "built build/$OUT"
