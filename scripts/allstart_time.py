"""Config 2 timing: all-start NN + 2-opt on pr2392 (reference: 4.68 core-hours)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import travellingsalesmanoptimization_b200 as T
from oracle.bindings import load_instance, fnv1a
for name, cnt in (("pcb442", 442), ("pr2392", 2392)):
    xy, _ = load_instance(name)
    I = T.Instance(xy)
    I.multistart_two_opt(0, 8)
    I.reset_stats()
    t = time.time()
    bs, bp, bc, nn_costs, final, moves = I.multistart_two_opt(0, cnt)
    dt = time.time() - t
    s = I.stats()
    print(name, "all-start NN+2opt: best", bc, "start", bs, "hash", fnv1a(bp), "wall %.2fs" % dt, "descent kernel %.2fs" % (s["scan_kernel_ms"] / 1e3),
          "moves", int(moves.sum()), "Gpairs/s %.0f" % (s["pair_evals"] / s["scan_kernel_ms"] / 1e6), flush=True)
    t = time.time(); bs2, bp2, bc2, costs = I.nn_multi(0, cnt); dt = time.time() - t
    print(name, "all-start NN: best", bc2, "start", bs2, "wall %.2fs" % dt, flush=True)
