#!/bin/bash
# last call of the round: smoke(), the whole GPU suite, the bench line, timings
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02f_smoke.log 2>&1; tail -2 gpurun_out/r02f_smoke.log
timeout -k 10 1200 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/r02f_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02f_tests.log; tail -4 gpurun_out/r02f_tests.log
timeout -k 10 900 python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench rc=$?"
python - <<P
import json
d=json.loads(open('gpurun_out/r02f_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['multistart']['seconds'], d['multistart_fixed']['pla85900']['seconds'], d['multistart_fixed']['usa13509']['seconds'], d['time_to_local_optimum_auto']['seconds'], d['matrix_build']['frac'], d['restarts']['tabu']['seconds'], d['restarts']['vns']['seconds'], d['clocks'])
P
timeout 300 python scripts/em_time.py > gpurun_out/r02f_em.log 2>&1; cat gpurun_out/r02f_em.log
timeout 300 python scripts/matrix_vs_recompute.py > gpurun_out/r02f_mx.log 2>&1; tail -4 gpurun_out/r02f_mx.log
