#!/bin/bash
# final state of the round (after the last kernel changes): whole GPU suite, the bench line, ncu --set full of K1 / K7 / K8 + the launch list of the short bench
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
R=/tmp/ncurep; mkdir -p $R
timeout -k 10 1200 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/r02w_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02w_tests.log; tail -4 gpurun_out/r02w_tests.log
timeout -k 10 900 python bench.py > gpurun_out/r02w_bench.json 2> gpurun_out/r02w_bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r02w_bench.json
timeout -k 10 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02w_bench_ref.json 2> gpurun_out/r02w_bench_ref.err; echo "ref rc=$?"; tail -c 600 gpurun_out/r02w_bench_ref.json
python scripts/matrix_time.py pla85900 > gpurun_out/r02w_k1_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:cost_matrix_kernel -s 1 -c 1 -o $R/k1 python scripts/matrix_time.py pla85900 > gpurun_out/r02w_ncu_k1.log 2>&1
python scripts/matrix_time.py rand20000 > gpurun_out/r02w_k1b_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:cost_matrix_kernel -s 1 -c 1 -o $R/k1b python scripts/matrix_time.py rand20000 > gpurun_out/r02w_ncu_k1b.log 2>&1
python scripts/pruned_time.py pla85900 > gpurun_out/r02w_k7_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:descend_pruned_kernel -s 2 -c 1 -o $R/k7 python scripts/pruned_time.py pla85900 > gpurun_out/r02w_ncu_k7.log 2>&1
python scripts/k8_one.py usa13509 148 > gpurun_out/r02w_k8_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:multi_pruned_kernel -c 1 -o $R/k8 python scripts/k8_one.py usa13509 148 > gpurun_out/r02w_ncu_k8.log 2>&1
SHORT="python bench.py --steps 2 --warmup 3 --full-descent 0 --cpu-seconds 0.5 --fixed 0"
$SHORT > gpurun_out/r02w_short_plain.json 2> gpurun_out/r02w_short_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02w_launches.csv $SHORT > gpurun_out/r02w_ncu_list.log 2>&1
for k in k1 k1b k7 k8; do
  [ -f $R/$k.ncu-rep ] && ncu -i $R/$k.ncu-rep --page raw --csv > gpurun_out/r02w_$k.raw.csv 2>/dev/null
  [ -f $R/$k.ncu-rep ] && python scripts/ncu_summary.py $R/$k.ncu-rep > gpurun_out/r02w_$k.summary.md 2>&1
done
[ -f $R/k1.ncu-rep ] && ncu -i $R/k1.ncu-rep --page source --csv 2>/dev/null | head -c 3000000 > gpurun_out/r02w_k1.source.csv
ls -la gpurun_out | grep r02r; du -sh gpurun_out
timeout 300 python scripts/r02_time.py k7 k8fast cl c5 > gpurun_out/r02w_time.log 2>&1; cat gpurun_out/r02w_time.log
