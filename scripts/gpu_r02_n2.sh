#!/bin/bash
# two GPUs of one box: the multi-GPU tests, then the bench line the driver would produce at N = 2
mkdir -p gpurun_out
N=${1:-2}
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py tests/test_driver_e2e.py tests/test_gpu_northstar.py -m gpu -q -k "mgpu or multi_gpu or restarts" --timeout 420 -p no:cacheprovider > gpurun_out/r02n${N}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02n${N}_tests.log; tail -4 gpurun_out/r02n${N}_tests.log
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02n${N}_bench.json 2> gpurun_out/r02n${N}_bench.err
echo "bench rc=$?"; tail -c 3000 gpurun_out/r02n${N}_bench.json; tail -5 gpurun_out/r02n${N}_bench.err
