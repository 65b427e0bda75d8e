set -x
mkdir -p gpurun_out
timeout 300 python scripts/cluster_outliers.py > gpurun_out/cluster_outliers3.log 2>&1; grep -v "^tspcuda: multistart" gpurun_out/cluster_outliers3.log; grep -c "release 0.00" gpurun_out/cluster_outliers3.log; grep "multistart" gpurun_out/cluster_outliers3.log | sort -t'+' -k5 | tail -3
timeout 300 python scripts/cluster_sweep.py pla85900 2>&1 | tee gpurun_out/cluster_sweep_pla.log
timeout 800 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
