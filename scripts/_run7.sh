set -x
mkdir -p gpurun_out
timeout 300 python scripts/cluster_outliers.py > gpurun_out/cluster_outliers2.log 2>&1; cat gpurun_out/cluster_outliers2.log
timeout 200 python -m pytest tests -x -q -m gpu -k "multistart" 2>&1 | tail -4
