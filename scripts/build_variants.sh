#!/bin/bash
# Builds launch-configuration variants of libtspcuda.so into build/variants/ (git-ignored; they
# travel to the GPU box).  usage: scripts/build_variants.sh "R BLOCK MINB UNROLL" ...
cd "$(dirname "$0")/../travellingsalesmanoptimization_b200/csrc" || exit 1
mkdir -p ../../build/variants
for v in "$@"; do
  set -- $v
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared \
    -DTSP_R=$1 -DTSP_BLOCK_BIG=$2 -DTSP_MINB_BIG=$3 -DTSP_UNROLL=$4 -o ../../build/variants/lib_r$1_b$2_m$3_u$4.so tspcuda.cu &
done
wait
