"""Generate tests/golden/golden_matrix.json from the UNMODIFIED reference objects: the `costs`-argument
paths (h_greedyutil / ref_2opt on a caller's matrix, h_Greedy_2opt_mod_costs, heuristics.c:118-149) on
the x*-weighted matrices of tests/matrices.py.  Run HERE (needs /root/reference):
`make -C oracle ref && python tests/golden/make_golden_matrix.py`.  FP64 costs are stored as C99 hex
floats (bit-exact); tours as FNV-1a-64 hashes of the successor array."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from matrices import xstar_matrix  # noqa: E402
from oracle.bindings import Reference, fnv1a  # noqa: E402

CASES = [(11, 64), (12, 150), (13, 333), (14, 700)]       # (seed, n), symmetric, zero diagonal

if __name__ == "__main__":
    R = Reference()
    G = {}
    for seed, n in CASES:
        M = xstar_matrix(seed, n)
        R.random_points(n, 1)                       # sets nnodes; the matrix is the caller's
        per = []
        for start in sorted({0, 1, n // 3, n - 2, n - 1}):
            e, p, c = R.nn_costs(M, start)
            e2, p2, c2 = R.ref_2opt_costs(M, p)
            assert e == 0 and e2 == 0
            per.append({"start": start, "nn_cost": c.hex(), "nn_hash": fnv1a(p), "opt_cost": c2.hex(), "opt_hash": fnv1a(p2)})
        e, pm, cm = R.mod_costs(M)
        assert e == 0
        G["%d_%d" % (seed, n)] = {"seed": seed, "n": n, "per_start": per, "mod_costs": {"cost": cm.hex(), "hash": fnv1a(pm)}}
        print(seed, n, per[-1], G["%d_%d" % (seed, n)]["mod_costs"], flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "golden_matrix.json"), "w") as f:
        json.dump(G, f, indent=1)
