set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu -k "cluster_sizes" 2>&1 | tail -5 | tee gpurun_out/cluster_test.log
timeout 800 python -m pytest tests -x -q -m gpu --durations=15 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 450 python scripts/cluster_sweep.py 2>&1 | tee gpurun_out/cluster_sweep.log
