"""Development probe run on the GPU box: parity spot checks + first timings. Not a test."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import travellingsalesmanoptimization_b200 as T  # noqa: E402
from oracle.bindings import Oracle, fnv1a, load_instance  # noqa: E402

O = Oracle()
G = json.load(open(os.path.join(ROOT, "tests/golden/golden.json")))
print("devices", T.device_count(), flush=True)

xy, _ = load_instance("berlin52")
I = T.Instance(xy)
print("sqrt mismatches:", I.selfcheck_sqrt(), flush=True)
names = ["DADD", "DMUL", "DFMA", "F2F pair", "MUFU.RSQ", "FFMA", "edge cost"]
for k in range(7):
    print("pipe rate %-10s %.4g op/s" % (names[k], I.pipe_rate(k)), flush=True)

for name in ["berlin52", "pcb442", "u1060", "pr2392"]:
    xy, _ = load_instance(name)
    n = xy.shape[0]
    I = T.Instance(xy)
    M = I.cost_matrix()
    Mo = O.cost_matrix(xy)
    print(name, "matrix equal:", np.array_equal(M, Mo), flush=True)
    code, p, c = I.nn(0)
    e, po, co = O.nn(xy, 0)
    print(name, "nn0 equal:", np.array_equal(p, po), c, co, G[name]["nn0"]["cost"], flush=True)
    t = time.time()
    code, moves, p2, c2, log = I.two_opt(po, log_cap=64)
    dt = time.time() - t
    g = G[name]["nn0_2opt"]
    print(name, "2opt moves", moves, g["moves"], "cost", c2, g["cost"], "hash ok", fnv1a(p2) == g["hash"],
          "log ok", [list(x) for x in log[:32]] == g["first_moves"][:32], "%.3fs" % dt, I.stats(), flush=True)

for name in ["usa13509", "d15112"]:
    xy, _ = load_instance(name)
    I = T.Instance(xy)
    t = time.time(); code, p, c = I.nn(0); dt = time.time() - t
    print(name, "nn0", c, G[name]["nn0"]["cost"], fnv1a(p) == G[name]["nn0"]["hash"], "%.3fs" % dt, flush=True)
    code, moves, p2, c2, log = I.two_opt(p, max_moves=20, log_cap=20)
    g = G[name]["nn0_2opt_20"]
    print(name, "20 moves cost", c2, g["cost"], fnv1a(p2) == g["hash"], [list(x) for x in log] == g["first_moves"], I.stats(), flush=True)
    I.reset_stats()
    t = time.time(); code, moves, p3, c3, _ = I.two_opt(p); dt = time.time() - t
    s = I.stats()
    print(name, "full descent moves", moves, "cost", c3, "%.2fs" % dt, "kernel ms", s["scan_kernel_ms"],
          "Gpairs/s %.1f" % (s["pair_evals"] / s["scan_kernel_ms"] / 1e6), flush=True)

xy, _ = load_instance("pla85900")
I = T.Instance(xy)
t = time.time(); code, p, c = I.nn(0); dt = time.time() - t
e_ok = O.is_tour(p)
print("pla85900 nn0", c, "tour", e_ok, "%.2fs" % dt, flush=True)
I.tour_upload(p)
I.reset_stats()
t = time.time(); m, opt = I.tour_descend(64); dt = time.time() - t
s = I.stats()
print("pla85900 64 moves %.3fs kernel ms %.2f Gpairs/s %.1f" % (dt, s["scan_kernel_ms"], s["pair_evals"] / s["scan_kernel_ms"] / 1e6), flush=True)
cnt, log = I.tour_movelog(64)
la = [x[0] for x in log]; lb = [x[1] for x in log]; ld = [float(x[2]) for x in log]
t = time.time(); r, _ = O.audit_moves(xy, p, la, lb, ld, [0, 1, 2]); dt = time.time() - t
print("pla85900 audit (3 full CPU scans + 64 delta checks):", r, "%.1fs" % dt, flush=True)
