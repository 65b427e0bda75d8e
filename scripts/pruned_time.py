"""Time NN(0) -> 2-opt local optimum under the pruned / auto / full scan kernels."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import travellingsalesmanoptimization_b200 as T
from oracle.bindings import load_instance
names = sys.argv[1:] or ["pr2392", "usa13509", "d15112", "pla85900"]
for name in names:
    xy, _ = load_instance(name)
    n = xy.shape[0]
    for mode, label in ((T.SCAN_PRUNED, "pruned"), (T.SCAN_AUTO, "auto"), (T.SCAN_FULL, "full")):
        if label == "full" and n > 20000:
            continue
        with T.Instance(xy) as I:
            I.set_scan_mode(mode)
            code, p, c = I.nn(0)
            I.two_opt(p, max_moves=4)          # warm
            I.reset_stats()
            t0 = time.perf_counter()
            code, moves, p2, c2, _ = I.two_opt(p)
            wall = time.perf_counter() - t0
            s = I.stats()
            print("%-9s %-6s moves %6d cost %11d wall %8.4f s  full-kernel %9.2f ms (%d scans)  pruned-kernel %9.2f ms (%d scans, %.0f cand/scan, %.1f us/scan) launches %d"
                  % (name, label, moves, c2, wall, s["scan_kernel_ms"], s["scans"] - s["pruned_scans"], s["pruned_kernel_ms"], s["pruned_scans"],
                     s["candidate_evals"] / max(1, s["pruned_scans"]), 1e3 * s["pruned_kernel_ms"] / max(1, s["pruned_scans"]), s["kernel_launches"]), flush=True)
