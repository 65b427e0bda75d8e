set -x
timeout 300 python scripts/cluster_sweep.py usa13509 2>&1 | tee gpurun_out/cluster_sweep.log
