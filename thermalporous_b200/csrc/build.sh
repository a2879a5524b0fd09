#!/bin/bash
# Build libtpb200.so for sm_100a (B200). Usage: build.sh [extra nvcc flags]
# Each translation unit is compiled in parallel, then linked into thermalporous_b200/libtpb200.so.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
OUT="$HERE/../libtpb200.so"
OBJ="$HERE/../_obj"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
mkdir -p "$OBJ"
pids=()
for src in "$HERE"/*.cu; do
  o="$OBJ/$(basename "${src%.cu}").o"
  $NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
    -Xcompiler -fPIC -I"$ROOT/include" -I"$HERE" "$@" -c "$src" -o "$o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o "$OUT" "$OBJ"/*.o -ldl
echo "built $OUT"
