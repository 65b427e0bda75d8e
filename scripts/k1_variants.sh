#!/bin/bash
# builds the K1 stand-alone bench in several shapes (here, no GPU needed) -> build/k1v/;  run with scripts/gpu_r02_k.sh
# variant = "rows block minb rot st"
set -e
cd "$(dirname "$0")"
rm -rf ../build/k1v; mkdir -p ../build/k1v
for v in "8 128 1 0 0" "8 128 1 1 0" "8 256 2 0 0" "8 64 1 0 0" "4 128 1 0 0" "16 128 1 0 0" "8 192 1 0 0" "8 96 1 0 0"; do
  set -- $v
  nvcc -DTSP_K1_ROWS=$1 -DTSP_K1_BLOCK=$2 -DTSP_K1_MINB=$3 -DTSP_K1_ROT=$4 -DTSP_K1_ST=$5 -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xptxas -v -o ../build/k1v/k1_r$1_b$2_m$3_rot$4_st$5 k1_bench.cu 2>&1 | grep -A2 "Function properties.*cost_matrix_kernelINS_6EucInt" | grep -o "Used [0-9]* registers\|[0-9]* bytes spill stores" | tr '\n' ' '
  echo " <- rows=$1 block=$2 minb=$3 rot=$4 st=$5"
done
