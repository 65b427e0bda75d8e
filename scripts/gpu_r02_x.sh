#!/bin/bash
# the bench line with the scratch-sizing warm call (twice: is multistart.seconds stable now?)
mkdir -p gpurun_out
for i in 1 2; do
timeout -k 10 900 python bench.py > gpurun_out/r02x_bench_$i.json 2> gpurun_out/r02x_bench_$i.err; echo "bench rc=$?"
python - <<P
import json
d=json.loads(open('gpurun_out/r02x_bench_$i.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['multistart']['seconds'], d['multistart_fixed']['pla85900']['seconds'], d['multistart_fixed']['usa13509']['seconds'], d['time_to_local_optimum_auto']['seconds'], d['matrix_build']['frac'], d['restarts']['tabu']['seconds'], d['restarts']['vns']['seconds'])
P
done
