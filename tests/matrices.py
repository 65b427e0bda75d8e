"""Foreign cost matrices for the `costs`-argument paths (h_greedyutil / ref_2opt on a caller's matrix,
h_Greedy_2opt_mod_costs): shaped like the x*-weighted matrices of cplex_model.c:1193-1201 --
c_ij * (1 - x*_ij), symmetric, zero diagonal, many exact zeros (x* = 1) and ties -- plus an asymmetric
variant and one with NOT_CONNECTED (-1.0) entries, which h_greedyutil skips (heuristics.c:257)."""
import numpy as np


def xstar_matrix(seed, n, symmetric=True, not_connected=0.0):
    rng = np.random.default_rng(seed)
    xy = np.round(rng.uniform(0, 1000, size=(n, 2)))
    d = np.floor(np.sqrt(((xy[:, None, :] - xy[None, :, :]) ** 2).sum(-1)) + 0.5)
    x = rng.uniform(0, 1, size=(n, n))
    kind = rng.uniform(0, 1, size=(n, n))
    x = np.where(kind < 0.15, 1.0, np.where(kind < 0.55, 0.0, x))       # LP solutions are mostly 0/1
    if symmetric:
        iu = np.triu_indices(n, 1)
        x[(iu[1], iu[0])] = x[iu]
    m = d * (1.0 - x)
    if not_connected > 0:
        m = np.where(rng.uniform(0, 1, size=(n, n)) < not_connected, -1.0, m)
    np.fill_diagonal(m, 0.0)
    return np.ascontiguousarray(m, dtype=np.float64)
