set -x
mkdir -p gpurun_out
for cl in 4 1; do
  timeout 60 python scripts/k8_one.py usa13509 9 $cl || exit 1
  timeout 300 ncu --set full --clock-control none -k regex:multi_pruned_kernel -c 1 -o /tmp/prof_k8_cl$cl -f python scripts/k8_one.py usa13509 9 $cl > gpurun_out/ncu_k8_cl$cl.log 2>&1
  tail -2 gpurun_out/ncu_k8_cl$cl.log
  python scripts/ncu_summary.py /tmp/prof_k8_cl$cl.ncu-rep > gpurun_out/k8_cl$cl.summary.md 2>&1
  ncu -i /tmp/prof_k8_cl$cl.ncu-rep --page raw --csv > gpurun_out/k8_cl$cl.raw.csv 2>/dev/null
done
ls -la gpurun_out
