#!/bin/bash
# K8: 128-node line blocks dealt round-robin against contiguous ranges; K7 move-phase instrumentation
mkdir -p gpurun_out
timeout 600 python scripts/r02_time.py k8fast > gpurun_out/r02p_k8_rr.log 2>&1; cat gpurun_out/r02p_k8_rr.log
TSPCUDA_LIB=$PWD/build/variants/lib_ranges.so timeout 600 python scripts/r02_time.py k8fast > gpurun_out/r02p_k8_ranges.log 2>&1; cat gpurun_out/r02p_k8_ranges.log
TSPCUDA_VERBOSE=1 TSPCUDA_LIB=$PWD/build/variants/lib_phase.so timeout 300 python scripts/r02_time.py k7 > gpurun_out/r02p_k7_phase.log 2>&1; grep "phases\|^K7" gpurun_out/r02p_k7_phase.log | awk 'NR%3!=1'
timeout -k 10 900 python -m pytest tests/test_gpu_northstar.py tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "northstar or cluster or multistart" > gpurun_out/r02p_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02p_tests.log; tail -5 gpurun_out/r02p_tests.log
