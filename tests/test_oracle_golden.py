"""CPU: the oracle (oracle/onfly.c) against the golden values generated from the unmodified
reference objects (tests/golden/golden.json, tests/golden/make_golden.py) and against the
reference's published deterministic columns (results/heuristics-ric.csv:2-15)."""
import numpy as np
import pytest

from oracle.bindings import ATT, CEIL_2D, EUC_REF, GEO, fnv1a, load_instance

SMALL = ["berlin52", "eil51", "kroA100", "pcb442", "pr1002", "u1060", "rl1304"]


@pytest.mark.parametrize("name", SMALL)
def test_nn0_and_descent(oracle, golden, name):
    xy, _ = load_instance(name)
    g = golden[name]
    e, p, c = oracle.nn(xy, 0)
    assert e == 0 and int(c) == g["nn0"]["cost"] and fnv1a(p) == g["nn0"]["hash"]
    m, p2, c2, la, lb, ld = oracle.two_opt(xy, p)
    assert m == g["nn0_2opt"]["moves"] and int(c2) == g["nn0_2opt"]["cost"] and fnv1a(p2) == g["nn0_2opt"]["hash"]
    k = len(g["nn0_2opt"]["first_moves"])
    assert [[int(a), int(b), int(d)] for a, b, d in zip(la[:k], lb[:k], ld[:k])] == g["nn0_2opt"]["first_moves"]
    assert oracle.is_tour(p2) and oracle.tour_cost(xy, p2) == c2


@pytest.mark.parametrize("name", ["berlin52", "eil51", "kroA100", "pcb442"])
def test_all_start_nn(oracle, golden, name):
    xy, _ = load_instance(name)
    best, bp, bc, costs = oracle.nn_all(xy, 0, xy.shape[0])
    g = golden[name]["allnn"]
    assert (int(bc), best, fnv1a(bp)) == (g["cost"], g["start"], g["hash"])
    per = golden[name]["greedy_2opt"]["per_start"]
    assert [int(c) for c in costs] == [r[0] for r in per]


def test_published_nn_columns(oracle, golden):
    """results/heuristics-ric.csv NN / allNN columns (deterministic)."""
    for name, row in golden["published_heuristics_ric"].items():
        xy, _ = load_instance(name)
        e, p, c = oracle.nn(xy, 0)
        assert int(c) == row["NN"], name
    xy, _ = load_instance("pr1002")
    best, bp, bc, _ = oracle.nn_all(xy, 0, xy.shape[0])
    assert int(bc) == golden["published_heuristics_ric"]["pr1002"]["allNN"] and best == golden["pr1002"]["allnn"]["start"]


@pytest.mark.parametrize("name", ["berlin52", "kroA100"])
def test_greedy_2opt_per_start(oracle, golden, name):
    xy, _ = load_instance(name)
    g = golden[name]["greedy_2opt"]
    best, best_start, best_hash = None, -1, None
    for s, (nn_cost, moves, final) in enumerate(g["per_start"]):
        e, p, c = oracle.nn(xy, s)
        m, p2, c2, *_ = oracle.two_opt(xy, p, log=False)
        assert (int(c), m, int(c2)) == (nn_cost, moves, final), (name, s)
        if best is None or c2 < best:
            best, best_start, best_hash = c2, s, fnv1a(p2)
    assert (int(best), best_start, best_hash) == (g["cost"], g["first_start"], g["hash"])


@pytest.mark.parametrize("name", ["berlin52", "kroA100", "pcb442"])
def test_tabu_steps(oracle, golden, name):
    xy, _ = load_instance(name)
    n = xy.shape[0]
    g = golden[name]["tabu_steps"]
    e, p, c = oracle.nn(xy, 0)
    m, p, c, *_ = oracle.two_opt(xy, p, log=False)
    tenure, lo, hi = oracle.tabu_init(n)
    assert [tenure, lo, hi] == g["tenure0"]
    inc = True
    tl = np.full(n, -1, dtype=np.int32)
    for k, (g_tenure, g_cost, g_hash) in enumerate(g["log"]):
        tenure, inc = oracle.tabu_linear_step(tenure, lo, hi, inc)
        c, a, b, d = oracle.tabu_move(xy, p, c, tl, tenure, k)
        assert (tenure, int(c), fnv1a(p)) == (g_tenure, g_cost, g_hash), (name, k)
    assert fnv1a(tl) == g["tabu_list_hash"]


def test_large_first_moves(oracle, golden):
    for name in ["usa13509"]:
        xy, _ = load_instance(name)
        g = golden[name]
        e, p, c = oracle.nn(xy, 0)
        assert int(c) == g["nn0"]["cost"] and fnv1a(p) == g["nn0"]["hash"]
        m, p2, c2, la, lb, ld = oracle.two_opt(xy, p, max_moves=5)
        assert [[int(a), int(b), int(d)] for a, b, d in zip(la, lb, ld)] == g["nn0_2opt_20"]["first_moves"][:5]


def test_random_points_stream(oracle, golden):
    """-n 2000 -seed 123 (src/tsp.c:468-481): glibc rand() stream and the descent from NN(0)."""
    g = golden["random_n2000_seed123"]
    xy = oracle.random_points(2000, 123)
    assert [float(xy[0, 0]), float(xy[0, 1])] == g["xy_first"] and fnv1a(xy.view(np.uint32).ravel()) == g["xy_hash"]
    e, p, c = oracle.nn(xy, 0)
    assert int(c) == g["nn0"]["cost"] and fnv1a(p) == g["nn0"]["hash"]
    m, p2, c2, la, lb, ld = oracle.two_opt(xy, p, max_moves=32)
    assert [[int(a), int(b), int(d)] for a, b, d in zip(la, lb, ld)] == g["nn0_2opt"]["first_moves"]


def test_cost_rule_vs_tsplib_nint(oracle):
    """The reference's float-sqrt idiom is NOT TSPLIB nint(sqrt()) in double: pcb442 differs on 10 of
    97 461 pairs -- the oracle must reproduce the reference, not TSPLIB."""
    xy, _ = load_instance("pcb442")
    M = oracle.cost_matrix(xy)
    d = xy[:, None, :] - xy[None, :, :]
    T = np.floor(np.sqrt(d[..., 0] ** 2 + d[..., 1] ** 2) + 0.5)
    iu = np.triu_indices(442, 1)
    assert int((M[iu] != T[iu]).sum()) == 10
    assert np.all(np.diag(M) == -1.0) and np.array_equal(M, M.T)
    # canonical tour 1..n: TSPLIB95's documented check value for pcb442
    canon = np.roll(np.arange(442, dtype=np.int32), -1)
    assert oracle.tour_cost(xy, canon) == 221440


def test_tsplib_weight_types_unpinned(oracle):
    """CEIL_2D / ATT / GEO have no reference behaviour (tsp.c:576-583); the oracle follows TSPLIB95.
    External KATs: canonical-tour lengths documented by TSPLIB95 (att532 309636, gr666 423710)."""
    xy, wt = load_instance("att532")
    assert wt == "ATT"
    canon = np.roll(np.arange(xy.shape[0], dtype=np.int32), -1)
    assert oracle.tour_cost(xy, canon, ATT) == 309636
    xy, wt = load_instance("gr666")
    canon = np.roll(np.arange(xy.shape[0], dtype=np.int32), -1)
    assert oracle.tour_cost(xy, canon, GEO) == 423710
    xy, wt = load_instance("pla7397")
    assert wt == "CEIL_2D"
    M = oracle.cost_matrix(xy[:200], CEIL_2D)
    d = xy[:200, None, :] - xy[None, :200, :]
    E = np.ceil(np.sqrt(d[..., 0] ** 2 + d[..., 1] ** 2))
    np.fill_diagonal(E, -1.0)
    assert np.array_equal(M, E)


def test_edge_cases(oracle):
    # n = 1, 2, 3: every pair is skipped by refinment.c:55; NN closes the tour
    for n in (1, 2, 3):
        xy = np.array([[0.0, 0.0], [3.0, 4.0], [10.0, 0.0]])[:n].copy()
        e, p, c = oracle.nn(xy, 0)
        assert e == 0 and oracle.is_tour(p)
        m, p2, c2, *_ = oracle.two_opt(xy, p)
        assert m == 0 and np.array_equal(p, p2)
    xy = np.array([[0.0, 0.0], [3.0, 4.0], [10.0, 0.0]])
    assert oracle.nn(xy, 3)[0] == 14 and oracle.nn(xy, -1)[0] == 14      # UNAVAILABLE, heuristics.c:223-226
    # duplicate points: zero-cost edges, ties broken by lowest id
    xy = np.array([[0.0, 0.0], [5.0, 5.0], [0.0, 0.0], [5.0, 5.0], [9.0, 1.0]])
    e, p, c = oracle.nn(xy, 0)
    assert p[0] == 2 and oracle.is_tour(p)


def test_foreign_matrix_golden(oracle):
    """Oracle restatement of the `costs`-argument paths against tests/golden/golden_matrix.json
    (generated from the unmodified reference objects by make_golden_matrix.py): FP64 costs bit-equal
    (hex floats), tours by hash."""
    import json
    import os
    from matrices import xstar_matrix
    with open(os.path.join(os.path.dirname(__file__), "golden", "golden_matrix.json")) as f:
        G = json.load(f)
    for key, g in G.items():
        M = xstar_matrix(g["seed"], g["n"])
        for r in g["per_start"]:
            e, p, c = oracle.nn_matrix(M, r["start"])
            assert e == 0 and c.hex() == r["nn_cost"] and fnv1a(p) == r["nn_hash"], (key, r["start"])
            m, p2, c2, *_ = oracle.two_opt_matrix(M, p)
            assert c2.hex() == r["opt_cost"] and fnv1a(p2) == r["opt_hash"], (key, r["start"])
        if g["n"] <= 333:                        # all-start loop: O(n) descents on the CPU
            pm, cm = oracle.mod_costs(M)
            assert cm.hex() == g["mod_costs"]["cost"] and fnv1a(pm) == g["mod_costs"]["hash"], key


def test_extra_mileage_golden(oracle):
    """Oracle restatement of h_ExtraMileage / h_extramileage_util against tests/golden/golden_em.json."""
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "golden_em.json")) as f:
        G = json.load(f)
    for name in ["berlin52", "eil51", "kroA100", "pcb442", "pr1002"]:
        xy, _ = load_instance(name)
        p, c = oracle.extra_mileage(xy)
        assert (int(c), fnv1a(p)) == (G[name]["em_max"]["cost"], G[name]["em_max"]["hash"])
        for a, b, cost, h in G[name]["from"][:2]:
            p, c = oracle.extra_mileage(xy, a, b)
            assert (int(c), fnv1a(p)) == (cost, h)


def test_patching_golden(oracle):
    """Oracle restatement of cx_patching (cplex_model.c:693-755, repeated while ncomp > 1) against
    tests/golden/golden_patching.json (the unmodified reference cplex_model.c, compiled with a stub <cplex.h>)."""
    import json
    import os
    from components import cycles_solution, solution_cost
    with open(os.path.join(os.path.dirname(__file__), "golden", "golden_patching.json")) as f:
        G = json.load(f)
    for key, g in G.items():
        name, seed, ncomp, local = key.rsplit("_", 3)
        xy, _ = load_instance(name)
        if xy.shape[0] > 1400:
            continue                                   # the larger cases are checked on the GPU only (CPU time)
        path, comp, nc = cycles_solution(int(seed), xy.shape[0], int(ncomp), bool(int(local)), xy)
        cost = solution_cost(oracle, xy, path)
        assert cost == g["start_cost"]
        patches = 0
        while nc > 1:
            ok, nc, cost, a, b, d = oracle.patching(xy, path, comp, nc, cost)
            assert ok and d > 0 and a < b
            patches += 1
        assert (patches, cost, fnv1a(path), fnv1a(comp)) == (g["patches"], g["cost"], g["path_hash"], g["comp_hash"]), key
        assert oracle.is_tour(path) and len(set(comp.tolist())) == 1


def test_patching_no_admissible_pair(oracle):
    """Every cross pair has delta <= 0 (coincident 2-cycles): the reference would index path[-1]
    (cplex_model.c:730-735); the restatement reports 'no patch' and changes nothing."""
    xy = np.array([[0, 0], [10, 0], [0, 0], [10, 0]], dtype=np.float64)
    path = np.array([1, 0, 3, 2], dtype=np.int32)
    comp = np.array([1, 1, 2, 2], dtype=np.int32)
    ok, nc, cost, a, b, d = oracle.patching(xy, path, comp, 2, 40.0)
    assert not ok and nc == 2 and cost == 40.0 and path.tolist() == [1, 0, 3, 2] and comp.tolist() == [1, 1, 2, 2]
