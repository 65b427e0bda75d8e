#!/bin/bash
# heap records in flight per lane (NB_HU): 2 / 4 (main) / 6 / 8, K8 and K7 on one box
mkdir -p gpurun_out
for v in main hu2 hu6 hu8; do
  if [ $v = main ]; then unset TSPCUDA_LIB; else export TSPCUDA_LIB=$PWD/build/variants/lib_$v.so; fi
  echo "== $v"; timeout 300 python scripts/r02_time.py k8fast k7 > gpurun_out/r02z_hu_$v.log 2>&1; cat gpurun_out/r02z_hu_$v.log
done
