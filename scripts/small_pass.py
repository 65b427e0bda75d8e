"""A small pass through the changed kernels as a quick end-to-end pass (compute-sanitizer is not available on this pool): K1 (odd and even n, tail rows), NN, the
pruned descent K7 (flat and row-by-row searches, settle pass), the tabu loop K5c, the VNS loop (giant boxes evaluated, then cached),
K8 multi-start.  Instances are small."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import travellingsalesmanoptimization_b200 as T
rng = np.random.default_rng(7)
for n in (7, 515, 1030):
    xy = np.floor(rng.random((n, 2)) * 5000.0)
    with T.Instance(xy) as I:
        M = I.cost_matrix()
        assert M.shape == (n, n) and (np.diag(M) == -1).all() and np.array_equal(M, M.T)
xy = np.floor(rng.random((4500, 2)) * 30000.0)
with T.Instance(xy) as I:
    I.set_scan_mode(T.SCAN_PRUNED)
    code, p, c = I.nn(0)
    code, moves, p2, c2, _ = I.two_opt(p, max_moves=40)       # ends on the move cap: settle pass
    code, moves2, p3, c3, _ = I.two_opt(p2)
    print("K7", moves, moves2, c3)
    tab = np.full(4500, -1, dtype=np.int32)
    out = I.tabu_search(p3.copy(), c3, tab, [600] * 30)
    print("K5c", out[0])
    ctypes.CDLL(None).srand(3)
    out = I.vns_search(p3, 6, c3)
    print("VNS", out[2])
xy = np.floor(rng.random((9000, 2)) * 50000.0)
with T.Instance(xy) as I:
    r = I.multistart_two_opt(0, 6, 1)
    print("K8", r[0], r[2])
print("ok")
