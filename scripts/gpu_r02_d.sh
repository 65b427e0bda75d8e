#!/bin/bash
mkdir -p gpurun_out
TSPCUDA_LIB=$PWD/build/variants/lib_nbstats.so timeout -k 10 600 python scripts/r02_time.py k7 > gpurun_out/r02d_nbstats.log 2>&1
cat gpurun_out/r02d_nbstats.log
