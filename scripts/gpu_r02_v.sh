#!/bin/bash
# apply_move batch size A/B on one box: 4 (main) / 2 / 1 pairs per thread and trip, K7 and K8
mkdir -p gpurun_out
for i in 1 2; do
for v in main am1 am2; do
  if [ $v = main ]; then unset TSPCUDA_LIB; else export TSPCUDA_LIB=$PWD/build/variants/lib_$v.so; fi
  echo "== $v run $i"; timeout 300 python scripts/r02_time.py k7 > gpurun_out/r02v_k7_${v}_$i.log 2>&1; cat gpurun_out/r02v_k7_${v}_$i.log
done; done
for v in main am1 am2; do
  if [ $v = main ]; then unset TSPCUDA_LIB; else export TSPCUDA_LIB=$PWD/build/variants/lib_$v.so; fi
  echo "== $v K8"; timeout 300 python scripts/r02_time.py k8fast > gpurun_out/r02v_k8_$v.log 2>&1; cat gpurun_out/r02v_k8_$v.log
done
