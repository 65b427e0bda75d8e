#!/bin/bash
# ncu --set full of K8 / K1 / K2 / K7; the reports stay on the box, their raw pages come back as csv
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
R=/tmp/ncurep; mkdir -p $R
python scripts/k8_one.py usa13509 148 > gpurun_out/r02g_k8_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:multi_pruned_kernel -c 1 -o $R/k8 python scripts/k8_one.py usa13509 148 > gpurun_out/r02g_ncu_k8.log 2>&1
python scripts/matrix_time.py pla85900 > gpurun_out/r02g_k1_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:cost_matrix_kernel -s 1 -c 1 -o $R/k1 python scripts/matrix_time.py pla85900 > gpurun_out/r02g_ncu_k1.log 2>&1
python scripts/matrix_time.py rand20000 > gpurun_out/r02g_k1b_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:cost_matrix_kernel -s 1 -c 1 -o $R/k1b python scripts/matrix_time.py rand20000 > gpurun_out/r02g_ncu_k1b.log 2>&1
SHORT="python bench.py --steps 2 --warmup 3 --full-descent 0 --cpu-seconds 0.5 --fixed 0 --multistart 0"
$SHORT > gpurun_out/r02g_short_plain.json 2> gpurun_out/r02g_short_plain.err &&
ncu --set full --clock-control none -k regex:descend_kernel -s 3 -c 1 -o $R/k2 $SHORT > gpurun_out/r02g_ncu_k2.log 2>&1
python scripts/pruned_time.py pla85900 > gpurun_out/r02g_k7_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:descend_pruned_kernel -s 1 -c 1 -o $R/k7 python scripts/pruned_time.py pla85900 > gpurun_out/r02g_ncu_k7.log 2>&1
for k in k8 k1 k1b k2 k7; do
  [ -f $R/$k.ncu-rep ] && ncu -i $R/$k.ncu-rep --page raw --csv > gpurun_out/r02g_$k.raw.csv 2>/dev/null
  [ -f $R/$k.ncu-rep ] && python scripts/ncu_summary.py $R/$k.ncu-rep > gpurun_out/r02g_$k.summary.md 2>&1
done
[ -f $R/k8.ncu-rep ] && ncu -i $R/k8.ncu-rep --page source --csv 2>/dev/null | head -c 6000000 > gpurun_out/r02g_k8.source.csv
timeout -k 10 600 python bench.py > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; echo "bench rc=$?"
ls -la gpurun_out | grep r02g; du -sh gpurun_out
