"""Generate tests/golden/golden_large.json + tests/golden/large_tours.npz: full 2-opt descents at the
config-3/5 sizes that the reference objects need hours for (1.2 - 1.8 s per scan).

Producer = the CPU oracle (oracle/onfly.c, OpenMP), which is pinned move-for-move on the unmodified
reference objects (tests/test_oracle_vs_reference.py) and on the reference-generated goldens for the
SAME instances (usa13509 / d15112: NN(0) cost, the first 20 moves, usa13509's 1983 moves -> 21119812 from the
reference run of BASELINE.md).  Run HERE: `python tests/golden/make_golden_large.py` (about two minutes on 8 cores).

  d15112     NN(0) -> 2-opt local optimum: moves, cost, FNV hash, the tour itself (config 5's start tour)
  rand20000  the reference's `-n 20000 -seed 123` instance (tsp.c:468-481), same
  usa13509   re-derived as a self-check against the reference-run golden (1983 moves -> 21119812)
For the two config-5 instances additionally, from that local optimum:
  tabu       64 iterations of tabu_best_move with the linear tenure schedule (metaheuristic.c:40-59, 188-245):
             cost after every iteration, hash of the final tour and tabu list
  vns        6 iterations of the loop of mh_VNS (metaheuristic.c:277-321) with srand(7): cost trace, final hash
"""
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.bindings import Oracle, fnv1a, load_instance  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "golden_large.json")
TOURS = os.path.join(ROOT, "tests", "golden", "large_tours.npz")
TABU_ITERS, VNS_ITERS, VNS_SEED = 64, 6, 7


def main():
    O = Oracle()
    libc = ctypes.CDLL(None)
    G, tours = {}, {}
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        ref_golden = json.load(f)
    for name in ("usa13509", "d15112", "rand20000"):
        xy = O.random_points(20000, 123) if name == "rand20000" else load_instance(name)[0]
        n = xy.shape[0]
        t0 = time.time()
        e, p, c = O.nn(xy, 0)
        m, p2, c2, la, lb, ld = O.two_opt(xy, p)
        g = G[name] = {"n": n, "nn0": {"cost": int(c), "hash": fnv1a(p)},
                       "nn0_2opt": {"moves": int(m), "cost": int(c2), "hash": fnv1a(p2),
                                    "last_moves": [[int(a), int(b), int(d)] for a, b, d in zip(la[-8:], lb[-8:], ld[-8:])]}}
        print(name, g["nn0"], g["nn0_2opt"]["moves"], g["nn0_2opt"]["cost"], "%.1f s" % (time.time() - t0), flush=True)
        if name in ref_golden:                      # the reference objects' own values for the same instance
            r = ref_golden[name]
            assert r["nn0"] == g["nn0"], name
            assert r["nn0_2opt_20"]["first_moves"] == [[int(a), int(b), int(d)] for a, b, d in zip(la[:20], lb[:20], ld[:20])], name
            if "nn0_2opt_survey" in r:
                assert (r["nn0_2opt_survey"]["moves"], r["nn0_2opt_survey"]["cost"]) == (int(m), int(c2)), name
        if name == "usa13509":
            continue
        tours[name] = p2.astype(np.int32)
        # tabu: linear tenure from tabu_init, K iterations from the local optimum
        tenure, lo, hi = O.tabu_init(n)
        inc, tl, pt, ct, trace = True, np.full(n, -1, dtype=np.int32), p2.copy(), float(c2), []
        for k in range(TABU_ITERS):
            tenure, inc = O.tabu_linear_step(tenure, lo, hi, inc)
            ct, a, b, d = O.tabu_move(xy, pt, ct, tl, tenure, k)
            trace.append(int(ct))
        g["tabu"] = {"iters": TABU_ITERS, "trace": trace, "hash": fnv1a(pt), "tabu_list_hash": fnv1a(tl)}
        print(name, "tabu", trace[:3], trace[-1], flush=True)
        # VNS: the loop of mh_VNS from the local optimum, glibc stream seeded
        libc.srand(VNS_SEED)
        pv, best, bi, vtrace, vmoves = p2.copy(), float(c2), -1, [], []
        for i in range(VNS_ITERS):
            mv, pv, cv, *_ = O.two_opt(xy, pv, log=False)
            if cv < best:
                best, bi = cv, i
            vtrace.append(int(cv))
            vmoves.append(int(mv))
            for _ in range(libc.rand() % 9 - 2):
                O.vns_kick(pv)
        g["vns"] = {"iters": VNS_ITERS, "seed": VNS_SEED, "trace": vtrace, "moves": vmoves, "best_cost": int(best), "best_iter": bi,
                    "hash": fnv1a(pv)}
        print(name, "vns", vtrace, vmoves, flush=True)
    with open(OUT, "w") as f:
        json.dump(G, f, indent=1, sort_keys=True)
    np.savez_compressed(TOURS, **tours)
    print("wrote", OUT, TOURS)


if __name__ == "__main__":
    main()
