"""Extra Mileage timings (EM_MAX initial edge)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import travellingsalesmanoptimization_b200 as T
from oracle.bindings import load_instance
for name in sys.argv[1:] or ["rl1304", "pr2392", "usa13509", "d15112", "pla85900"]:
    xy, _ = load_instance(name)
    with T.Instance(xy) as I:
        I.extra_mileage(0, 1) if xy.shape[0] < 3000 else None
        t0 = time.perf_counter(); p, c, ab = I.extra_mileage(); dt = time.perf_counter() - t0
        print("%-9s n=%6d  extra mileage %.4f s (kernels %.4f s)  cost %d  first edge %s" % (name, xy.shape[0], dt, I.stats()["last_kernel_ms"] * 1e-3, c, ab), flush=True)
