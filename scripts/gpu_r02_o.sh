#!/bin/bash
# K7 / K5c: the words the whole grid reads after a barrier go through one reader per CTA + shared memory; whole suite
mkdir -p gpurun_out
for i in 1 2; do
timeout 300 python scripts/r02_time.py k7 > gpurun_out/r02o_k7_$i.log 2>&1; cat gpurun_out/r02o_k7_$i.log
done
TSPCUDA_LIB=$PWD/build/variants/lib_noflat.so timeout 300 python scripts/r02_time.py k7 > gpurun_out/r02o_k7_before.log 2>&1; cat gpurun_out/r02o_k7_before.log
TSPCUDA_VERBOSE=1 TSPCUDA_LIB=$PWD/build/variants/lib_phase.so timeout 300 python scripts/r02_time.py k7 > gpurun_out/r02o_k7_phase.log 2>&1; grep "phases\|^K7" gpurun_out/r02o_k7_phase.log | awk 'NR%3!=1' 
timeout 300 python scripts/r02_time.py c5 > gpurun_out/r02o_c5.log 2>&1; cat gpurun_out/r02o_c5.log
TSPCUDA_LIB=$PWD/build/variants/lib_noflat.so timeout 300 python scripts/r02_time.py c5 > gpurun_out/r02o_c5_before.log 2>&1; cat gpurun_out/r02o_c5_before.log
timeout -k 10 1200 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/r02o_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02o_tests.log; tail -5 gpurun_out/r02o_tests.log
