#!/bin/bash
# 2 GPUs: the CLI multi-GPU test; pipe rates incl. the FP32-only cost sequence; ncu launch list of the short bench
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_driver_e2e.py -m gpu -q -k "multi_gpu" --timeout 420 -p no:cacheprovider > gpurun_out/r02i_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02i_tests.log; tail -3 gpurun_out/r02i_tests.log
timeout 300 python - > gpurun_out/r02i_pipe.log 2>&1 <<'P'
import sys
sys.path.insert(0, '.')
import travellingsalesmanoptimization_b200 as T
xy, _ = T.load_instance("pr2392")
names = {0: "DADD", 2: "DFMA", 4: "MUFU.RSQ", 5: "FFMA", 8: "F2F.F32.F64", 7: "EucInt::K (scalar sequence)", 9: "EucInt::K2 (shipped: 4 FP64 + F2F + MUFU + 3 FP32x2 per cost)", 10: "EucIntF32::K2 (FP32 only: MUFU + 10.5 FP32x2 + 2 LOP3 + FMNMX per cost)"}
with T.Instance(xy) as I:
    for k, nm in names.items():
        print("kind %3d  %-75s %.4e /s" % (k, nm, I.pipe_rate(k)), flush=True)
    print("kind 101  mismatches of kind 10 against kind 9 over 1.5e9 triples: %d" % int(I.pipe_rate(101)))
P
cat gpurun_out/r02i_pipe.log
SHORT="python bench.py --steps 2 --warmup 3 --full-descent 0 --cpu-seconds 0.5 --fixed 0"
$SHORT > gpurun_out/r02i_short_plain.json 2> gpurun_out/r02i_short_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02i_launches.csv $SHORT > gpurun_out/r02i_ncu_list.log 2>&1
wc -l gpurun_out/r02i_launches.csv
