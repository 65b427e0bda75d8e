#!/bin/bash
# K8/K7: giant runs kept in place when the new box is no larger; heap restart policy sweep (TSPCUDA_NB_RESTART_PCT)
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_northstar.py tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -p no:cacheprovider -k "northstar or cluster or multistart or two_opt or vns or descent" > gpurun_out/r02s_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02s_tests.log; tail -4 gpurun_out/r02s_tests.log
echo "== keep (main)"; timeout 300 python scripts/r02_time.py k8fast k7 > gpurun_out/r02s_keep.log 2>&1; cat gpurun_out/r02s_keep.log
echo "== nokeep"; TSPCUDA_LIB=$PWD/build/variants/lib_nokeep.so timeout 300 python scripts/r02_time.py k8fast k7 > gpurun_out/r02s_nokeep.log 2>&1; cat gpurun_out/r02s_nokeep.log
for pct in 25 50 200 400; do
echo "== keep, restart pct $pct"; TSPCUDA_NB_RESTART_PCT=$pct timeout 300 python scripts/r02_time.py k8fast k7 > gpurun_out/r02s_keep_pct$pct.log 2>&1; grep "pla85900\|usa13509  1184\|pr2392" gpurun_out/r02s_keep_pct$pct.log
done
echo "== stats"; TSPCUDA_VERBOSE=1 TSPCUDA_LIB=$PWD/build/variants/lib_stats.so timeout 600 python scripts/r02_time.py k8fast > gpurun_out/r02s_stats.log 2>&1; grep "delta cache\|^K8" gpurun_out/r02s_stats.log | tail -14
