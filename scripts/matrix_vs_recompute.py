"""Config 2's question: explicit cost matrix (gathers) or recomputation (on-the-fly costs) for the 2-opt scan?
pr2392 (45.8 MB matrix, L2-resident) and usa13509 (1.46 GB, HBM): K10 on the instance's own matrix against K2."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import travellingsalesmanoptimization_b200 as T
from oracle.bindings import load_instance
for name, moves in (("pr2392", 309), ("usa13509", 200)):
    xy, _ = load_instance(name)
    n = xy.shape[0]
    with T.Instance(xy) as I:
        I.set_scan_mode(T.SCAN_FULL)
        M = I.cost_matrix()
        code, p, c = I.nn(0)
        I.two_opt(p, max_moves=4)
        I.reset_stats()
        code, mv, p2, c2, _ = I.two_opt(p, max_moves=moves)
        s = I.stats()
        rate_fly = s["pair_evals"] / (s["scan_kernel_ms"] * 1e-3)
    with T.MatrixInstance(M) as X:
        X.two_opt(p, max_moves=4)
        t0 = time.perf_counter()
        code, mvx, px, cx, _ = X.two_opt(p, max_moves=moves)
        sx = X.stats()
    assert mvx == mv and cx == c2 and np.array_equal(px, p2)
    # the second call's kernel time = total - first call's; use last_kernel_ms of the chunked launches: approximate by scan_kernel_ms delta
    pairs = (mvx + (1 if mvx < moves else 0)) * n * (n - 1) / 2
    print("%-9s n=%5d: recompute (K2) %.1f Gpair/s | explicit matrix (K10, %0.f MB) %.1f Gpair/s  (same %d moves, same tour)"
          % (name, n, rate_fly / 1e9, 8.0 * n * n / 1e6, pairs / (sx["last_kernel_ms"] * 1e-3) / 1e9, mv), flush=True)
