#!/bin/bash
mkdir -p gpurun_out
export TSPCUDA_LIB=$PWD/build/variants/lib_nbstats.so
timeout -k 10 600 python scripts/r02_time.py k7 > gpurun_out/r02d_nbstats.log 2>&1
timeout -k 10 600 python - >> gpurun_out/r02d_nbstats.log 2>&1 <<'P'
import sys, time
sys.path.insert(0, '.')
import travellingsalesmanoptimization_b200 as T
for name, count in (("pr2392", 592), ("usa13509", 148), ("pla85900", 148)):
    xy, _ = T.load_instance(name)
    with T.Instance(xy) as I:
        I.multistart_two_opt(0, 148)
        I.reset_stats()
        t0 = time.perf_counter(); r = I.multistart_two_opt(0, count); w = time.perf_counter() - t0
        s = I.stats()
        print("K8", name, count, "wall %.3f kernel %.3f us/scan/SM %.1f" % (w, s["pruned_kernel_ms"] * 1e-3, 1e3 * s["pruned_kernel_ms"] * min(count, 148) / max(1, s["pruned_scans"])), flush=True)
P
cat gpurun_out/r02d_nbstats.log
