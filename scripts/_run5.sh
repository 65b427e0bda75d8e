set -x
timeout 450 python scripts/cluster_sweep.py 2>&1 | tee gpurun_out/cluster_sweep.log
