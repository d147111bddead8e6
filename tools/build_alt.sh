#!/bin/bash
# builds an experiment variant of the library: tools/build_alt.sh <name> <nvcc -D flags...>  ->  csrc/alt/libgast_b200_<name>.so
# (loaded through GAST_B200_LIB; *.so files are git-ignored but travel to the GPU box)
set -e
cd "$(dirname "$0")/../gast-net-3dposeestimation_b200/csrc"
name=$1; shift
mkdir -p alt
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O2 -shared "$@" \
  -o alt/libgast_b200_$name.so gast_api.cu -lcudart 2>/dev/null
echo "built alt/libgast_b200_$name.so"
