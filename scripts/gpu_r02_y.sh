#!/bin/bash
# K8 scan sections (warp 1 of CTA 0): heap stream / lines / tail
mkdir -p gpurun_out
for a in "pla85900 148" "usa13509 148" "d15112 148"; do
TSPCUDA_LIB=$PWD/build/variants/lib_phase.so timeout 300 python scripts/k8_one.py $a > gpurun_out/r02y_k8_sections_${a%% *}.log 2>&1; echo "== $a"; tail -3 gpurun_out/r02y_k8_sections_${a%% *}.log
done
