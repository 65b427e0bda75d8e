#!/bin/bash
# bench line, K1 unit-size sweep, ncu launch list of the short bench, ncu --set full of K8 / K1 / K2
mkdir -p gpurun_out
timeout -k 10 600 python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench rc=$?"
for c in 0 1 2 4 8 16 64; do
  echo "TSPCUDA_K1_CHUNK=$c" >> gpurun_out/r02f_k1_sweep.log
  TSPCUDA_K1_CHUNK=$c timeout 300 python scripts/matrix_time.py rand20000 pla85900 >> gpurun_out/r02f_k1_sweep.log 2>&1
done
cat gpurun_out/r02f_k1_sweep.log
SHORT="python bench.py --steps 2 --warmup 3 --full-descent 0 --cpu-seconds 0.5 --fixed 0"
$SHORT > gpurun_out/r02f_short_plain.json 2> gpurun_out/r02f_short_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_launches.csv $SHORT > gpurun_out/r02f_ncu_list.log 2>&1
python scripts/k8_one.py usa13509 148 > gpurun_out/r02f_k8_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:multi_pruned_kernel -c 1 -o gpurun_out/r02f_prof_k8 python scripts/k8_one.py usa13509 148 > gpurun_out/r02f_ncu_k8.log 2>&1
tail -3 gpurun_out/r02f_ncu_k8.log
python scripts/matrix_time.py pla85900 > gpurun_out/r02f_k1_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:cost_matrix_kernel -s 1 -c 1 -o gpurun_out/r02f_prof_k1 python scripts/matrix_time.py pla85900 > gpurun_out/r02f_ncu_k1.log 2>&1
tail -3 gpurun_out/r02f_ncu_k1.log
$SHORT > gpurun_out/r02f_short_plain2.json 2> gpurun_out/r02f_short_plain2.err &&
ncu --set full --clock-control none --import-source on -k regex:descend_kernel -s 3 -c 1 -o gpurun_out/r02f_prof_k2 $SHORT > gpurun_out/r02f_ncu_k2.log 2>&1
tail -3 gpurun_out/r02f_ncu_k2.log
ls -la gpurun_out | grep r02f
