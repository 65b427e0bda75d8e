#!/bin/bash
# K8 phase clock (team 0): where the 309 us of a pla85900 scan + move go
mkdir -p gpurun_out
TSPCUDA_LIB=$PWD/build/variants/lib_phase.so timeout 300 python scripts/k8_one.py pla85900 148 > gpurun_out/r02t_k8_phase_pla.log 2>&1; cat gpurun_out/r02t_k8_phase_pla.log | tail -5
TSPCUDA_LIB=$PWD/build/variants/lib_phase.so timeout 300 python scripts/k8_one.py usa13509 148 > gpurun_out/r02t_k8_phase_usa.log 2>&1; cat gpurun_out/r02t_k8_phase_usa.log | tail -4
TSPCUDA_LIB=$PWD/build/variants/lib_phase.so timeout 300 python scripts/k8_one.py d15112 148 > gpurun_out/r02t_k8_phase_d15.log 2>&1; cat gpurun_out/r02t_k8_phase_d15.log | tail -4
TSPCUDA_LIB=$PWD/build/variants/lib_phase.so timeout 300 python scripts/k8_one.py pla85900 18 4 > gpurun_out/r02t_k8_phase_pla_cl4.log 2>&1; cat gpurun_out/r02t_k8_phase_pla_cl4.log | tail -4
