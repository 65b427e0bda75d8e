"""Generate tests/golden/golden_patching.json from the UNMODIFIED reference objects: cx_patching
(cplex_model.c:693-755) repeated while ncomp > 1 (its caller's loop, :205-212) on seeded multi-cycle
solutions (tests/components.py).  cplex_model.c is compiled unchanged against oracle/stubs/cx/cplex.h
(oracle/Makefile).  Run HERE (needs /root/reference): `make -C oracle ref && python tests/golden/make_golden_patching.py`."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from components import cycles_solution  # noqa: E402
from oracle.bindings import Reference, fnv1a, load_instance  # noqa: E402

CASES = [("berlin52", 1, 2, True), ("berlin52", 2, 6, True), ("berlin52", 3, 17, False), ("kroA100", 4, 33, False),
         ("pcb442", 5, 25, True), ("rl1304", 6, 60, True), ("pr2392", 7, 40, True), ("pr2392", 8, 300, False),
         ("usa13509", 9, 12, True)]


def run_case(R, xy, seed, ncomp, local):
    path, comp, ncomp = cycles_solution(seed, xy.shape[0], ncomp, local, xy)
    cost = float(R.L.refapi_solution_cost(path.ctypes.data_as(R.L.refapi_patching.argtypes[0])))
    start_cost, patches = cost, 0
    while ncomp > 1:
        ncomp, cost = R.patching(path, comp, ncomp, cost)
        patches += 1
    return {"start_cost": start_cost, "patches": patches, "cost": cost, "path_hash": fnv1a(path), "comp_hash": fnv1a(comp)}


if __name__ == "__main__":
    import ctypes
    R = Reference()
    R.L.refapi_solution_cost.restype = ctypes.c_double
    G = {}
    for name, seed, ncomp, local in CASES:
        xy, _ = load_instance(name)
        R.set_points(xy)
        G["%s_%d_%d_%d" % (name, seed, ncomp, int(local))] = g = run_case(R, xy, seed, ncomp, local)
        print(name, seed, ncomp, local, g, flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "golden_patching.json"), "w") as f:
        json.dump(G, f, indent=1)
