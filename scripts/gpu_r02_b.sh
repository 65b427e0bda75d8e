#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python scripts/r02_time.py k7 k8 cl c5 > gpurun_out/r02b_time.log 2>&1
echo "time rc=$?"; cat gpurun_out/r02b_time.log
timeout -k 10 1500 python -m pytest tests -m gpu -q --maxfail=25 --timeout 420 -p no:cacheprovider > gpurun_out/r02b_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02b_tests.log
tail -5 gpurun_out/r02b_tests.log
