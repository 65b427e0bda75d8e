#!/bin/bash
# Builds variants of libtspcuda.so into build/variants/ (git-ignored; they travel to the GPU box).
# usage: scripts/build_variants.sh "name -DFLAG=.. -DFLAG=.." ...
cd "$(dirname "$0")/../travellingsalesmanoptimization_b200/csrc" || exit 1
mkdir -p ../../build/variants
for v in "$@"; do
  set -- $v
  name=$1; shift
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared "$@" \
    -o ../../build/variants/lib_$name.so tspcuda.cu &
done
wait
