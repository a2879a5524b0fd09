#!/bin/bash
# Build libtpb200.so for sm_100a (B200). Usage: build.sh [extra nvcc flags]
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
OUT="$HERE/../libtpb200.so"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
SRCS=$(ls "$HERE"/*.cu)
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
  -Xcompiler -fPIC -shared -I"$ROOT/include" -I"$HERE" "$@" \
  -o "$OUT" $SRCS -ldl
echo "built $OUT"
