"""Instance loading for bench.py / scripts (data only, no algorithm).

  read_tsplib(path)    a TSPLIB file with a NODE_COORD_SECTION, by the grammar of the reference's reader
                       (src/tsp.c:527-606: keyword lines split at blanks and ':', node lines "index x y", index = atoi - 1,
                       values through atof -- Python's float() is the same correctly rounded strtod)
  load_instance(name)  the compact fixtures of the public TSPLIB instances the benchmarks name
                       (tests/golden/instances/*.npz, written by tests/golden/make_instances.py)
Both return (xy float64 (n, 2) C-contiguous, EDGE_WEIGHT_TYPE string)."""
import os

import numpy as np

FIXTURES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "instances")


def read_tsplib(path):
    n, wtype, pts, in_nodes = -1, "EUC_2D", None, False
    with open(path) as f:
        for line in f:
            toks = line.replace(":", " ").replace(",", " ").split()
            if not toks:
                continue
            key = toks[0]
            if key.startswith("EOF"):
                break
            if key.startswith("DIMENSION"):
                n = int(toks[1])
                pts = np.zeros((n, 2), dtype=np.float64)
            elif key.startswith("NODE_COORD_SECTION"):
                in_nodes = True
            elif key.startswith("EDGE_WEIGHT_TYPE"):
                wtype = toks[1]
            elif in_nodes and pts is not None:
                i = int(toks[0]) - 1
                if 0 <= i < n:
                    pts[i] = (float(toks[1]), float(toks[2]))
    if pts is None:
        raise ValueError("%s: no DIMENSION / NODE_COORD_SECTION" % path)
    return np.ascontiguousarray(pts), wtype


def load_instance(name):
    z = np.load(os.path.join(FIXTURES, name + ".npz"))
    return np.ascontiguousarray(z["xy"], dtype=np.float64), str(z["edge_weight_type"])
