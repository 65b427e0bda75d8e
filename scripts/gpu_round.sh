#!/bin/bash
# One GPU-box visit: smoke, bench, then the ncu launch list and one --set full capture of the
# dominant kernel (each ncu command only after the same command exited 0 without ncu).
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1 || { tail -20 gpurun_out/smoke.log; exit 1; }
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err || { tail -20 gpurun_out/bench.err; exit 1; }
cat gpurun_out/bench.json
SHORT="python bench.py --steps 2 --warmup 3 --full-descent 0 --cpu-seconds 0.5"
$SHORT > gpurun_out/short_plain.json 2> gpurun_out/short_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $SHORT > gpurun_out/ncu_list.log 2>&1
$SHORT > gpurun_out/short_plain2.json 2> gpurun_out/short_plain2.err &&
ncu --set full --clock-control none --import-source on -k regex:descend_kernel -s 3 -c 2 -o gpurun_out/prof_descend $SHORT > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/ncu_full.log
ls -la gpurun_out
# K7 (pruned descent): one --set full capture of a mid-descent launch
python scripts/pruned_time.py pla85900 > gpurun_out/pruned_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:descend_pruned_kernel -s 2 -c 1 -o gpurun_out/prof_pruned python scripts/pruned_time.py pla85900 > gpurun_out/ncu_pruned.log 2>&1
tail -3 gpurun_out/ncu_pruned.log
