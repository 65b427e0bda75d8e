"""Round-2 timings of the pruned-scan kernels through the C ABI (one B200):
  K7  NN(0) -> 2-opt local optimum (whole GPU)            pla85900 / usa13509 / d15112 / pr2392
  K8  multi-start NN + 2-opt (one CTA or cluster per descent)  pr2392 all starts, usa13509, pla85900; TSPCUDA_K8_BLOCK 1024 / 512
  config 5  tabu 1000 iterations, mh_VNS loop 200 iterations   d15112
usage: python scripts/r02_time.py [k7] [k8] [cl] [c5]"""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import travellingsalesmanoptimization_b200 as T
from travellingsalesmanoptimization_b200 import multistart as MS
what = set(sys.argv[1:]) or {"k7", "k8", "cl", "c5"}


def ms_call(I, first, count, stride=1):
    I.reset_stats()
    t0 = time.perf_counter()
    r = I.multistart_two_opt(first, count, stride)
    wall = time.perf_counter() - t0
    s = I.stats()
    return r, wall, s


if "k7" in what:
    for name in ("pr2392", "usa13509", "d15112", "pla85900"):
        xy, _ = T.load_instance(name)
        with T.Instance(xy) as I:
            I.set_scan_mode(T.SCAN_PRUNED)
            code, p, c = I.nn(0)
            I.two_opt(p, max_moves=4)
            for rep in range(2):
                I.reset_stats()
                t0 = time.perf_counter()
                code, moves, p2, c2, _ = I.two_opt(p)
                wall = time.perf_counter() - t0
                s = I.stats()
            print("K7 %-9s moves %6d cost %11d wall %8.4f s kernel %9.2f ms  %.1f us/scan  %.0f records+candidates/scan  launches %d"
                  % (name, moves, c2, wall, s["pruned_kernel_ms"], 1e3 * s["pruned_kernel_ms"] / max(1, s["pruned_scans"]),
                     s["candidate_evals"] / max(1, s["pruned_scans"]), s["kernel_launches"]), flush=True)

if "k8" in what or "k8fast" in what:
    for blk in (("1024", "512") if "k8" in what else ("1024",)):
        os.environ["TSPCUDA_K8_BLOCK"] = blk
        for name, count in (("pr2392", 2392), ("usa13509", 296), ("usa13509", 1184), ("d15112", 296), ("pla85900", 148), ("pla85900", 296)):
            xy, _ = T.load_instance(name)
            with T.Instance(xy) as I:
                ms_call(I, 0, min(count, 4))
                (bs, bp, bc, nn, fin, mv), wall, s = ms_call(I, 0, count)
                print("K8 block %-4s %-9s %5d starts: wall %7.3f s  kernel %8.3f s  moves %8d  %.1f us/scan/SM  best %d @ %d  %.0f rec+cand/scan"
                      % (blk, name, count, wall, s["pruned_kernel_ms"] * 1e-3, int(mv.sum()),
                         1e3 * s["pruned_kernel_ms"] * min(count, 148) / max(1, s["pruned_scans"]), int(bc), bs, s["candidate_evals"] / max(1, s["pruned_scans"])), flush=True)
    del os.environ["TSPCUDA_K8_BLOCK"]

if "cl" in what:
    for name, count in (("usa13509", 9), ("pla85900", 18)):
        xy, _ = T.load_instance(name)
        row = []
        for cl in ("0", "1", "2", "4", "8"):
            os.environ["TSPCUDA_K8_CLUSTER"] = cl
            with T.Instance(xy) as I:
                ms_call(I, 0, 2)
                (bs, bp, bc, nn, fin, mv), wall, s = ms_call(I, 0, count)
                row.append("CL=%s %.3f s (kernel %.3f)" % (cl, wall, (s["pruned_kernel_ms"] + s["scan_kernel_ms"]) * 1e-3))
        del os.environ["TSPCUDA_K8_CLUSTER"]
        print("clusters %-9s %3d starts: %s" % (name, count, "  ".join(row)), flush=True)
    xy, _ = T.load_instance("pla85900")
    with T.Instance(xy) as I:
        for count in (12, 18, 37, 74):
            ms_call(I, 0, 2)
            (bs, bp, bc, nn, fin, mv), wall, s = ms_call(I, 0, count)
            print("library default pla85900 %3d starts: wall %.3f s kernel %.3f s" % (count, wall, s["pruned_kernel_ms"] * 1e-3), flush=True)

if "c5" in what:
    libc = ctypes.CDLL(None)
    xy, _ = T.load_instance("d15112")
    n = xy.shape[0]
    with T.Instance(xy) as I:
        code, p, c = I.nn(0)
        code, moves, p, c, _ = I.two_opt(p)
        tenures = MS.tabu_tenures(n, 1000)
        for rep in range(2):
            pb, tb = p.copy(), np.full(n, -1, dtype=np.int32)
            I.reset_stats(); t0 = time.perf_counter()
            cb, trace, bi, bp, bc = I.tabu_search(pb, c, tb, tenures)
            t_loop = time.perf_counter() - t0; s2 = I.stats()
        print("c5 d15112 tabu 1000 iterations: %.4f s (kernels %.4f s, %d of %d scans pruned) final %d" % (
            t_loop, (s2["scan_kernel_ms"] + s2["pruned_kernel_ms"]) * 1e-3, s2["pruned_scans"], s2["scans"], cb), flush=True)
        for rep in range(2):
            libc.srand(1); I.reset_stats(); t0 = time.perf_counter()
            code, pv, cv, trace, bi, bp, bc, done = I.vns_search(p, 200, c)
            t_loop = time.perf_counter() - t0; s4 = I.stats()
        print("c5 d15112 mh_VNS loop 200 iterations: %.4f s (kernels %.4f s, %d moves, %d scans, %.0f rec+cand/scan) best %s" % (
            t_loop, (s4["scan_kernel_ms"] + s4["pruned_kernel_ms"]) * 1e-3, s4["moves"], s4["scans"], s4["candidate_evals"] / max(1, s4["pruned_scans"]), bc), flush=True)
        for kind, label, iters in ((T.RESTART_TABU, "tabu", 1000), (T.RESTART_VNS, "vns", 100)):
            I.restarts(kind, 0, 1, 1, iters=4)
            t0 = time.perf_counter()
            code, done, bs, bp, bc = I.restarts(kind, 0, 8, 1, iters=iters, seed=1)
            print("c5 d15112 %s restarts: 8 in %.3f s, best %d @ %d" % (label, time.perf_counter() - t0, int(bc), bs), flush=True)
