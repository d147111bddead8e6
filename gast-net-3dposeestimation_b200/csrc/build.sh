#!/bin/bash
# Builds libgast_b200.so in-tree for sm_100a (B200).  nvcc cross-compiles without a GPU.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
  -Xcompiler -fPIC,-O2,-Wall -shared ${GAST_NVCC_EXTRA} \
  -o libgast_b200.so.tmp gast_api.cu -lcudart
mv -f libgast_b200.so.tmp libgast_b200.so    # atomic: a snapshot of the tree never sees a half-written library
echo "built $(pwd)/libgast_b200.so"
