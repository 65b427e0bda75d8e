#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python scripts/r02_time.py k7 k8fast cl c5 > gpurun_out/r02e_time.log 2>&1
echo "time rc=$?"; cat gpurun_out/r02e_time.log
