#!/bin/bash
# delta-cache bring-up: the parity tests of the pruned kernels first, then timings
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_northstar.py -m gpu -q -x --timeout 420 -p no:cacheprovider > gpurun_out/r02c_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02c_tests.log
tail -30 gpurun_out/r02c_tests.log
timeout -k 10 900 python scripts/r02_time.py k7 k8 cl c5 > gpurun_out/r02c_time.log 2>&1
echo "time rc=$?"; cat gpurun_out/r02c_time.log
