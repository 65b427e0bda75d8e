set -x
python -m pytest tests -x -q -m gpu -k "vns or VNS or berlin52_cli or random_instances or patching" 2>&1 | tail -15
python scripts/config5_time.py 200 2>&1 | tee gpurun_out/config5_time.log
python scripts/patching_time.py 2>&1 | tee gpurun_out/patching_time.log
