#!/bin/bash
# first GPU pass of round 2: the whole -m gpu suite, then a short bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_smi.txt 2>&1
timeout -k 10 1500 python -m pytest tests -m gpu -q --maxfail=25 --timeout 420 -p no:cacheprovider > gpurun_out/r02a_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02a_tests.log
tail -5 gpurun_out/r02a_tests.log
timeout -k 10 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
echo "bench rc=$?"
tail -c 1500 gpurun_out/r02a_bench.json
