#!/bin/bash
mkdir -p gpurun_out
export TSPCUDA_LIB=$PWD/build/variants/lib_dev.so
timeout -k 10 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "mileage or extra" --timeout 600 -p no:cacheprovider > gpurun_out/r02j_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02j_tests.log; tail -5 gpurun_out/r02j_tests.log
timeout 600 python scripts/em_time.py > gpurun_out/r02j_em.log 2>&1; cat gpurun_out/r02j_em.log
