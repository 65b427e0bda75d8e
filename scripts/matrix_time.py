"""K1 cost-matrix build: achieved HBM write bandwidth (8 n^2 bytes / kernel time) per instance."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import travellingsalesmanoptimization_b200 as T
from oracle.bindings import Oracle, load_instance
peak = 6551.0
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
O = Oracle()
for name in (sys.argv[1:] or ("pr2392", "usa13509", "d15112", "rand20000", "pla85900")):
    xy = O.random_points(20000, 123) if name == "rand20000" else load_instance(name)[0]
    n = xy.shape[0]
    with T.Instance(xy) as I:
        I.cost_matrix_device()
        ms = min(I.cost_matrix_device() for _ in range(5))
        chk = I.cost_matrix_checksum()
        gb = 8.0 * n * n / 1e9
        print("%-9s n=%6d  %8.3f GB  %8.3f ms  %7.1f GB/s  = %.3f of measured peak %.0f GB/s   checksum %d" % (name, n, gb, ms, gb / (ms * 1e-3), gb / (ms * 1e-3) / peak, peak, chk), flush=True)
