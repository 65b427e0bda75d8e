#!/bin/bash
# K8: heap stream through TMA bulk copies + mbarrier stages (4 per warp) against direct loads, and 3 / 5 stages
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_northstar.py tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -p no:cacheprovider -k "northstar or cluster or multistart" > gpurun_out/r02q_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02q_tests.log; tail -5 gpurun_out/r02q_tests.log
timeout 300 python scripts/r02_time.py k8fast cl > gpurun_out/r02q_k8_st4.log 2>&1; cat gpurun_out/r02q_k8_st4.log
for v in nostage st3 st5; do
TSPCUDA_LIB=$PWD/build/variants/lib_$v.so timeout 300 python scripts/r02_time.py k8fast > gpurun_out/r02q_k8_$v.log 2>&1; echo "== $v"; cat gpurun_out/r02q_k8_$v.log
done
