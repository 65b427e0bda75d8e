"""NN construction and multi-start timings on the large instances."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import travellingsalesmanoptimization_b200 as T
from oracle.bindings import load_instance
for name, ns in (("pr2392", 2392), ("usa13509", 296), ("pla85900", 148)):
    xy, _ = load_instance(name)
    with T.Instance(xy) as I:
        I.nn(0)
        t0 = time.perf_counter(); I.nn(1); t1 = time.perf_counter() - t0
        t0 = time.perf_counter(); bs, bp, bc, costs = I.nn_multi(0, ns); tm = time.perf_counter() - t0
        I.reset_stats()
        t0 = time.perf_counter(); r = I.multistart_two_opt(0, ns); t2 = time.perf_counter() - t0
        s = I.stats()
        print("%s: nn(1) %.4f s | nn_multi %d starts %.3f s | multistart NN+2opt %d starts %.3f s (descent kernels %.3f s, %d moves) best %d @ %d"
              % (name, t1, ns, tm, ns, t2, (s["scan_kernel_ms"] + s["pruned_kernel_ms"]) * 1e-3, s["moves"], r[2], r[0]), flush=True)
