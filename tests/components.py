"""Successor arrays made of several disjoint cycles with component labels, shaped like what cx_build_sol
(cplex_model.c:182) hands to cx_patching: a degree-2 solution of the TSP model without subtour constraints
is a set of cycles of length >= 3, labelled 1..ncomp in order of their lowest node."""
import numpy as np


def cycles_solution(seed, n, ncomp, local=True, xy=None):
    """(path, comp, ncomp): n nodes cut into `ncomp` cycles (each >= 3 nodes).  local=True groups nodes that are
    close in the plane (what LP solutions look like); otherwise the grouping is random."""
    rng = np.random.default_rng(seed)
    ncomp = max(1, min(ncomp, n // 3))
    if local and xy is not None:
        order = np.lexsort((xy[:, 1], np.floor(xy[:, 0] / (np.ptp(xy[:, 0]) / max(1, int(np.sqrt(ncomp))) + 1e-9))))
    else:
        order = rng.permutation(n)
    cuts = np.sort(rng.choice(np.arange(1, n // 3), size=ncomp - 1, replace=False)) * 3 if ncomp > 1 else np.array([], dtype=int)
    bounds = np.concatenate(([0], cuts, [n])).astype(int)
    path = np.zeros(n, dtype=np.int32)
    comp = np.zeros(n, dtype=np.int32)
    groups = [order[bounds[k]:bounds[k + 1]] for k in range(ncomp)]
    groups.sort(key=lambda g: int(g.min()))
    for label, g in enumerate(groups, start=1):
        g = g[rng.permutation(len(g))] if not local else g
        path[g] = np.roll(g, -1)
        comp[g] = label
    return path, comp, ncomp


def solution_cost(oracle, xy, path, wt=0):
    return float(sum(oracle.edge(wt, xy[i, 0], xy[i, 1], xy[path[i], 0], xy[path[i], 1]) for i in range(len(path))))
