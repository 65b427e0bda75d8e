#!/bin/bash
# K8: batched move (4 pairs per thread and trip), searching warps out of the stream for one-CTA teams (split32), phase clock
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_northstar.py tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -p no:cacheprovider -k "northstar or cluster or multistart or two_opt or vns or descent or tabu" > gpurun_out/r02u_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02u_tests.log; tail -4 gpurun_out/r02u_tests.log
echo "== main"; timeout 300 python scripts/r02_time.py k8fast k7 cl > gpurun_out/r02u_main.log 2>&1; cat gpurun_out/r02u_main.log
echo "== split32"; TSPCUDA_LIB=$PWD/build/variants/lib_split32.so timeout 300 python scripts/r02_time.py k8fast > gpurun_out/r02u_split32.log 2>&1; cat gpurun_out/r02u_split32.log
for v in phase split32phase; do
for a in "pla85900 148" "usa13509 148" "d15112 148"; do
TSPCUDA_LIB=$PWD/build/variants/lib_$v.so timeout 300 python scripts/k8_one.py $a > gpurun_out/r02u_k8_${v}_${a%% *}.log 2>&1; echo "== $v $a"; tail -2 gpurun_out/r02u_k8_${v}_${a%% *}.log
done; done
