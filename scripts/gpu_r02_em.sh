#!/bin/bash
# Extra Mileage: rescans spread over the whole grid, last CTA finalises
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py tests/test_driver_e2e.py -m gpu -q -x --timeout 300 -p no:cacheprovider -k "mileage or extra" > gpurun_out/r02em_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02em_tests.log; tail -4 gpurun_out/r02em_tests.log
timeout 300 python scripts/em_time.py > gpurun_out/r02em_time.log 2>&1; cat gpurun_out/r02em_time.log
