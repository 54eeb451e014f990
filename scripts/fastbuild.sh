#!/bin/bash
# Incremental build for development: one object per .cu (in parallel, only the stale ones), then the same
# shared library __graft_entry__.build() produces (which remains the reference recipe).
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
CSRC="$ROOT/quad-periodic-mpc_b200/csrc"
OBJ="$ROOT/build/obj"
mkdir -p "$OBJ"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
pids=()
for f in "$CSRC"/*.cu; do
  o="$OBJ/$(basename "${f%.cu}").o"
  stale=0
  [ -f "$o" ] || stale=1
  if [ $stale = 0 ]; then
    for d in "$f" "$CSRC"/*.cuh "$CSRC"/*.h "$ROOT/include/cmpc_b200.h"; do [ "$d" -nt "$o" ] && stale=1; done
  fi
  if [ $stale = 1 ]; then (cd "$CSRC" && nvcc $FLAGS -c "$f" -o "$o") & pids+=($!); fi
done
for p in "${pids[@]}"; do wait $p; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$ROOT/quad-periodic-mpc_b200/libcmpc_b200.so" "$OBJ"/*.o
echo built
