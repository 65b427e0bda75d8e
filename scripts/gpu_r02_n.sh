#!/bin/bash
# K7: flat search of the usual small box + the single-thread jobs of a move spread over three threads; A/B against the row-by-row walk
mkdir -p gpurun_out
for i in 1 2; do
timeout 300 python scripts/r02_time.py k7 > gpurun_out/r02n_k7_flat_$i.log 2>&1; cat gpurun_out/r02n_k7_flat_$i.log
TSPCUDA_LIB=$PWD/build/variants/lib_noflat.so timeout 300 python scripts/r02_time.py k7 > gpurun_out/r02n_k7_noflat_$i.log 2>&1; cat gpurun_out/r02n_k7_noflat_$i.log
done
TSPCUDA_VERBOSE=1 TSPCUDA_LIB=$PWD/build/variants/lib_phase.so timeout 300 python scripts/r02_time.py k7 > gpurun_out/r02n_k7_phase.log 2>&1; grep "phases\|^K7" gpurun_out/r02n_k7_phase.log | awk 'NR%3!=1' 
timeout 600 python scripts/r02_time.py k8fast > gpurun_out/r02n_k8.log 2>&1; cat gpurun_out/r02n_k8.log
timeout 300 python scripts/vns_time.py 200 5 > gpurun_out/r02n_vns.log 2>&1; cat gpurun_out/r02n_vns.log
timeout -k 10 900 python -m pytest tests/test_gpu_northstar.py tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "northstar or two_opt or descent or cluster or vns or multistart" > gpurun_out/r02n_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02n_tests.log; tail -5 gpurun_out/r02n_tests.log
