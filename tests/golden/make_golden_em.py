"""Generate tests/golden/golden_em.json from the UNMODIFIED reference objects: h_ExtraMileage with the
EM_MAX initial edge (heuristics.c:156-210) and h_extramileage_util from given first edges (:290-367).
Run HERE (needs /root/reference): `make -C oracle ref && python tests/golden/make_golden_em.py`.
The published deterministic column `ExtraMileage` of results/heuristics-ric.csv (rl1304 307898, pr1002
302240, fl1577 25593 ...) is reproduced by the same call."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.bindings import Reference, fnv1a, load_instance  # noqa: E402

if __name__ == "__main__":
    R = Reference()
    G = {}
    for name in ["berlin52", "eil51", "kroA100", "pcb442", "pr1002", "u1060", "rl1304", "fl1577", "pr2392"]:
        xy, _ = load_instance(name)
        n = xy.shape[0]
        R.set_points(xy)
        e, p, c, s = R.run(5)
        assert e == 0
        g = G[name] = {"em_max": {"cost": int(c), "hash": fnv1a(p)}, "from": []}
        if n <= 1100:
            for a, b in ((0, 1), (n - 1, 3), (n // 2, n // 3)):
                e, p, c = R.extra_mileage_from(a, b)
                g["from"].append([a, b, int(c), fnv1a(p)])
        print(name, g, flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "golden_em.json"), "w") as f:
        json.dump(G, f, indent=1)
