#!/bin/bash
# K7 with the four searching warps taken out of the stream (split) against the same build without it, phase clock, K8 check
mkdir -p gpurun_out
for i in 1 2; do
timeout 300 python scripts/r02_time.py k7 > gpurun_out/r02m_k7_split_$i.log 2>&1; cat gpurun_out/r02m_k7_split_$i.log
TSPCUDA_LIB=$PWD/build/variants/lib_nosplit.so timeout 300 python scripts/r02_time.py k7 > gpurun_out/r02m_k7_nosplit_$i.log 2>&1; cat gpurun_out/r02m_k7_nosplit_$i.log
done
TSPCUDA_VERBOSE=1 TSPCUDA_LIB=$PWD/build/variants/lib_phase.so timeout 300 python scripts/r02_time.py k7 > gpurun_out/r02m_k7_phase.log 2>&1; grep -v "^$" gpurun_out/r02m_k7_phase.log | tail -40
timeout 600 python scripts/r02_time.py k8fast cl > gpurun_out/r02m_k8.log 2>&1; cat gpurun_out/r02m_k8.log
timeout 300 python scripts/vns_time.py 200 5 > gpurun_out/r02m_vns.log 2>&1; cat gpurun_out/r02m_vns.log
timeout -k 10 900 python -m pytest tests/test_gpu_northstar.py tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "northstar or two_opt or descent or cluster or vns" > gpurun_out/r02m_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02m_tests.log; tail -5 gpurun_out/r02m_tests.log
