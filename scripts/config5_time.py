"""Config 5 at the helper boundary: K tabu iterations and K VNS iterations on d15112 and
the 20 000-node uniform-random instance (-n 20000 -seed 123), from the NN(0)+2-opt tour.
Tabu: the per-call path (tspcuda_tabu_move, what the reference's own loop drives) against the
device-resident loop (tspcuda_tabu_search); both must leave the same tour."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import travellingsalesmanoptimization_b200 as T
from oracle.bindings import Oracle, load_instance
O = Oracle()
K = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
for name in ("d15112", "rand20000"):
    xy = O.random_points(20000, 123) if name == "rand20000" else load_instance(name)[0]
    n = xy.shape[0]
    with T.Instance(xy) as I:
        code, p, c = I.nn(0)
        t0 = time.perf_counter(); code, moves, p, c, _ = I.two_opt(p); t_opt = time.perf_counter() - t0
        tenure, lo, hi = O.tabu_init(n)
        inc, tenures = True, []
        for _ in range(K):
            tenure, inc = O.tabu_linear_step(tenure, lo, hi, inc)
            tenures.append(tenure)
        pa, ta, ca = p.copy(), np.full(n, -1, dtype=np.int32), c
        I.reset_stats(); t0 = time.perf_counter()
        for k, t in enumerate(tenures):
            ca, mv = I.tabu_move(pa, ca, ta, t, k)
        t_call = time.perf_counter() - t0; s1 = I.stats()
        pb, tb = p.copy(), np.full(n, -1, dtype=np.int32)
        I.reset_stats(); t0 = time.perf_counter()
        cb, trace, bi, bp, bc = I.tabu_search(pb, c, tb, tenures)
        t_loop = time.perf_counter() - t0; s2 = I.stats()
        assert cb == ca and np.array_equal(pa, pb) and np.array_equal(ta, tb)
        pairs = K * n * (n - 1) / 2
        print("%-9s n=%d NN(0)+2opt %d moves %.3f s -> %d | tabu K=%d: per-call %.3f s (kernel %.3f s, %.0f Gpair/s e2e)  device loop %.3f s (kernels %.3f s: %d of %d scans by the pruned scan; %.0f Gpair/s full-scan equivalent)  final %d best %s"
              % (name, n, moves, t_opt, c, K, t_call, s1["scan_kernel_ms"] * 1e-3, pairs / t_call / 1e9, t_loop, (s2["scan_kernel_ms"] + s2["pruned_kernel_ms"]) * 1e-3, s2["pruned_scans"], s2["scans"], pairs / t_loop / 1e9, cb, bc if bi >= 0 else None), flush=True)
        # VNS: K x (kick on the host + ref_2opt on the device), metaheuristic.c:279-320
        import ctypes
        pv, cv = p.copy(), c
        I.reset_stats(); t0 = time.perf_counter(); mv_total = 0
        for k in range(min(K, 200)):
            for _ in range(3):
                O.vns_kick(pv)
            code, m, pv, cv, _ = I.two_opt(pv)
            mv_total += m
        t_vns = time.perf_counter() - t0; s3 = I.stats()
        print("%-9s VNS %d iterations: %.3f s (%d moves, %d scans, %d pruned; kernels %.3f s)" % (name, min(K, 200), t_vns, mv_total, s3["scans"], s3["pruned_scans"], (s3["scan_kernel_ms"] + s3["pruned_kernel_ms"]) * 1e-3), flush=True)
        # the loop of mh_VNS itself (descent, then rand() % 9 - 2 kicks): per-call path (host successor arrays in and out of
        # every ref_2opt, the reference's kick on the host) against the device-resident loop (tspcuda_vns_search)
        libc = ctypes.CDLL(None)
        KV = min(K, 200)
        libc.srand(1); pa = p.copy(); I.reset_stats(); t0 = time.perf_counter()
        for k in range(KV):
            code, m, pa, ca, _ = I.two_opt(pa)
            for _ in range(libc.rand() % 9 - 2):
                O.vns_kick(pa)
        t_call = time.perf_counter() - t0
        libc.srand(1); I.reset_stats(); t0 = time.perf_counter()
        code, pb, cb, trace, bi, bp, bc, done = I.vns_search(p, KV, c)
        t_loop = time.perf_counter() - t0; s4 = I.stats()
        assert np.array_equal(pa, pb) and ca == cb and done == KV
        print("%-9s mh_VNS loop %d iterations: per-call %.3f s  device-resident %.3f s (kernels %.3f s)  best %s" % (
            name, KV, t_call, t_loop, (s4["scan_kernel_ms"] + s4["pruned_kernel_ms"]) * 1e-3, bc), flush=True)
