set -x
python -m pytest tests -x -q -m gpu -k "patching" 2>&1 | tail -5
python scripts/patching_time.py 2>&1 | tee gpurun_out/patching_time.log
ncu --set full --clock-control none --import-source on -k regex:patch_scan_kernel -s 4 -c 1 -o gpurun_out/prof_patch python scripts/patching_time.py > gpurun_out/ncu_patch.log 2>&1
tail -3 gpurun_out/ncu_patch.log
