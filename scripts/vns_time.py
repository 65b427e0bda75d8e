"""The loop of mh_VNS (metaheuristic.c:277-321) through tspcuda_vns_search: 200 iterations from the NN(0) + 2-opt tour, the
process rand() stream seeded with srand(1), for several values of TSPCUDA_NB_GIANT_AGE (how often a giant search box is
evaluated before it is cached, pruned.cuh).  The tour, cost and trace must not depend on the setting."""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import travellingsalesmanoptimization_b200 as T
libc = ctypes.CDLL(None)
K = int(sys.argv[1]) if len(sys.argv) > 1 else 200
ages = sys.argv[2].split(",") if len(sys.argv) > 2 else ["0", "1", "2", "3", "5", "1000"]
for name in ("d15112", "usa13509", "rl1304"):
    xy, _ = T.load_instance(name)
    with T.Instance(xy) as I:
        code, p, c = I.nn(0)
        code, moves, p, c, _ = I.two_opt(p)
        ref = None
        for age in ages:
            os.environ["TSPCUDA_NB_GIANT_AGE"] = age
            best = 1e9
            for rep in range(2):
                libc.srand(1); I.reset_stats(); t0 = time.perf_counter()
                code, pb, cb, trace, bi, bp, bc, done = I.vns_search(p, K, c)
                t = time.perf_counter() - t0; s = I.stats()
                best = min(best, t)
            sig = (int(cb), int(bc), int(np.asarray(trace).sum()), bytes(np.asarray(pb)).__hash__())
            if ref is None: ref = sig
            print("%-9s giant_age %-5s mh_VNS loop %d iterations: %.4f s (kernels %.4f s, %d scans, %d launches)  final %d best %d  %s" % (
                name, age, K, best, (s["scan_kernel_ms"] + s["pruned_kernel_ms"]) * 1e-3, s["pruned_scans"], s["kernel_launches"], cb, bc,
                "same" if sig == ref else "DIFFERENT"), flush=True)
