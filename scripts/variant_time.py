"""Time one build variant of libtspcuda.so (TSPCUDA_LIB) on full-scan (SCAN_FULL) descents."""
import hashlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import travellingsalesmanoptimization_b200 as T
from oracle.bindings import load_instance
out = []
for name, moves, reps in (("pla85900", 32, 3), ("usa13509", 256, 2), ("pr2392", 300, 2)):
    xy, _ = load_instance(name)
    I = T.Instance(xy)
    I.set_scan_mode(T.SCAN_FULL)
    code, p, c = I.nn(0)
    I.tour_upload(p); I.tour_descend(8)
    best = 0
    for _ in range(reps):
        I.tour_upload(p); I.tour_descend(4)
        I.reset_stats(); I.tour_descend(moves); s = I.stats()
        best = max(best, s["pair_evals"] / s["scan_kernel_ms"] / 1e6)
    cnt, log = I.tour_movelog(64)
    h = hashlib.md5(str(log).encode()).hexdigest()[:8]
    out.append("%s %.1f Gp/s log %s" % (name, best, h))
    I.close()
print(os.path.basename(os.environ.get("TSPCUDA_LIB", "default")), " | ".join(out), flush=True)
