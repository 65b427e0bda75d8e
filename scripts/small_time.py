"""Full-scan descent rate on small / mid instances (which block configuration below BIG_N?)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import travellingsalesmanoptimization_b200 as T
from oracle.bindings import load_instance
out = []
rng = np.random.default_rng(1)
for name in ("pcb442", "rl1304", "pr2392", "rand4000", "rand6000", "rand8000"):
    xy = np.round(rng.uniform(0, 100000, (int(name[4:]), 2))) if name.startswith("rand") else load_instance(name)[0]
    I = T.Instance(xy); I.set_scan_mode(T.SCAN_FULL)
    code, p, c = I.nn(0)
    I.tour_upload(p); I.tour_descend(8)
    best = 0
    for _ in range(3):
        I.tour_upload(p); I.tour_descend(4)
        I.reset_stats(); I.tour_descend(200); s = I.stats()
        best = max(best, s["pair_evals"] / s["scan_kernel_ms"] / 1e6)
    out.append("%s %.0f (%.1f us/scan)" % (name, best, 1e3 * s["scan_kernel_ms"] / max(1, s["scans"])))
    I.close()
print(os.path.basename(os.environ.get("TSPCUDA_LIB", "default")), " | ".join(out), flush=True)
