#!/bin/bash
mkdir -p gpurun_out
for v in k1_base k1_m3 k1_m4 k1_r4m4 k1_r4m6; do
  echo "== $v" >> gpurun_out/r02h_k1_variants.log
  TSPCUDA_LIB=$PWD/build/variants/lib_$v.so timeout 300 python scripts/matrix_time.py usa13509 rand20000 pla85900 >> gpurun_out/r02h_k1_variants.log 2>&1
done
cat gpurun_out/r02h_k1_variants.log
