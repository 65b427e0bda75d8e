#!/bin/bash
# K8, one CTA per descent: orientation / epoch bytes in shared memory (main) against global memory
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_northstar.py tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -p no:cacheprovider -k "northstar or cluster or multistart or maximum_size" > gpurun_out/r02o8_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02o8_tests.log; tail -4 gpurun_out/r02o8_tests.log
echo "== o8 in shared memory"; timeout 300 python scripts/r02_time.py k8fast > gpurun_out/r02o8_smem.log 2>&1; cat gpurun_out/r02o8_smem.log
echo "== o8 in global memory"; TSPCUDA_LIB=$PWD/build/variants/lib_o8global.so timeout 300 python scripts/r02_time.py k8fast > gpurun_out/r02o8_global.log 2>&1; cat gpurun_out/r02o8_global.log
