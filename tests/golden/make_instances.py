"""Generate the TSPLIB instance fixtures under tests/golden/instances/ (run HERE, where
/root/reference exists; the GPU box only sees the committed .npz files).

Each fixture holds the node coordinates exactly as the reference's reader would produce them
(src/tsp.c:590-599: index = atoi(tok)-1, x = atof(tok), y = atof(tok); Python float() is the same
correctly-rounded strtod) plus the TSPLIB header's EDGE_WEIGHT_TYPE.  The TSPLIB files themselves
are public benchmark data the reference merely redistributes under data/; they are stored here in
a compact binary form, not as copies of the text files.

    python tests/golden/make_instances.py [/root/reference/data]
"""
import os
import sys

import numpy as np

NAMES = [
    "berlin52", "eil51", "kroA100", "pcb442", "pr1002", "u1060", "rl1304", "fl1577", "pr2392",
    "usa13509", "d15112", "pla85900",          # EUC_2D (pla85900: CEIL_2D header)
    "pla7397", "dsj1000",                      # CEIL_2D
    "att48", "att532",                         # ATT
    "burma14", "ulysses16", "gr666",           # GEO
]


def parse(path):
    n, wtype, pts, in_nodes = -1, None, None, False
    with open(path) as f:
        for line in f:
            if len(line) <= 1:
                continue
            toks = line.replace(":", " ").replace(",", " ").split()
            if not toks:
                continue
            key = toks[0]
            if key.startswith("DIMENSION"):
                n = int(toks[1]); pts = np.zeros((n, 2), dtype=np.float64); continue
            if key.startswith("NODE_COORD_SECTION"):
                in_nodes = True; continue
            if key.startswith("EDGE_WEIGHT_TYPE"):
                wtype = toks[1]; continue
            if key.startswith("EOF"):
                break
            if in_nodes:
                i = int(toks[0]) - 1
                pts[i, 0] = float(toks[1]); pts[i, 1] = float(toks[2])
    return n, wtype, pts


def main():
    data = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/data"
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "instances")
    os.makedirs(out, exist_ok=True)
    for name in NAMES:
        n, wtype, pts = parse(os.path.join(data, name + ".tsp"))
        assert n == pts.shape[0] and wtype is not None, name
        np.savez_compressed(os.path.join(out, name + ".npz"), xy=pts, edge_weight_type=np.array(wtype))
        print(f"{name:10s} n={n:6d} {wtype:8s} integer_coords={bool(np.all(pts == np.floor(pts)))} "
              f"x=[{pts[:,0].min()},{pts[:,0].max()}] y=[{pts[:,1].min()},{pts[:,1].max()}]")


if __name__ == "__main__":
    main()
