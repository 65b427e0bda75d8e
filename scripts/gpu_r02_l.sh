#!/bin/bash
# new K1 (one unit per CTA, branch-free trips) through the library; giant-box deferral in the VNS loop; whole GPU suite
mkdir -p gpurun_out
timeout 600 python scripts/matrix_time.py > gpurun_out/r02l_matrix.log 2>&1; cat gpurun_out/r02l_matrix.log
timeout 600 python scripts/vns_time.py > gpurun_out/r02l_vns.log 2>&1; cat gpurun_out/r02l_vns.log
timeout 600 python scripts/r02_time.py k7 > gpurun_out/r02l_k7.log 2>&1; cat gpurun_out/r02l_k7.log
timeout -k 10 1200 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/r02l_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02l_tests.log; tail -5 gpurun_out/r02l_tests.log
