"""Patching (cx_patching loop) timings: whole `while (ncomp > 1)` loop in one call; pair rate of the scan kernel
(n(n-1)/2 unordered pairs per scan, two fresh edge costs each)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import travellingsalesmanoptimization_b200 as T
from oracle.bindings import Oracle, load_instance
from components import cycles_solution
O = Oracle()
CASES = [("pr2392", 300), ("usa13509", 400), ("d15112", 400), ("pla85900", 60)]
for name, ncomp in [c for c in CASES if len(sys.argv) < 2 or c[0] in sys.argv[1:]]:
    xy, _ = load_instance(name)
    n = xy.shape[0]
    path, comp, nc = cycles_solution(1, n, ncomp, True, xy)
    with T.Instance(xy) as I:
        I.patching(path, comp, nc, 0.0, max_patches=2)
        t0 = time.perf_counter(); code, p, c, n1, k, log = I.patching(path, comp, nc, 0.0); dt = time.perf_counter() - t0
        ms = I.stats()["last_kernel_ms"]
        pairs = len(log) * n * (n - 1) / 2
        print("%-9s n=%6d  %d patches: %.4f s wall, kernels %.4f s = %.1f us per patch, %.3e pairs/s (%.3e edge costs/s)  tour %s" % (
            name, n, len(log), dt, ms * 1e-3, ms * 1e3 / max(1, len(log)), pairs / (ms * 1e-3), 2 * pairs / (ms * 1e-3), O.is_tour(p)), flush=True)
