#!/bin/bash
# Build libchol_b200.so for sm_100a (cross-compiles without a GPU).
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=../libchol_b200.so
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
      -Xptxas -v -Xcompiler -fPIC -shared -t 4 -o "$OUT" chol_abi.cu peer_abi.cu 2> build.log || { cat build.log; exit 1; }
grep -E "error|warning" build.log || true
echo "built $(realpath $OUT)"
