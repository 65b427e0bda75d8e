"""CPU: pin the oracle restatement (oracle/onfly.c) against the UNMODIFIED reference objects
(oracle/_ref/libtspref.so through oracle/ref_api.c), move for move, on fresh seeded inputs.
Skipped where oracle/_ref was never built (it needs /root/reference at build time)."""
import os
import tempfile

import numpy as np
import pytest

from oracle.bindings import fnv1a, load_instance


@pytest.fixture(scope="module", autouse=True)
def _cwd_with_results():
    old = os.getcwd()
    td = tempfile.mkdtemp()
    os.makedirs(os.path.join(td, "results"))      # metaheuristic.c:97 writes results/TabuResults.dat
    os.chdir(td)
    yield
    os.chdir(old)


def _rand_instance(rng, n, integer):
    xy = rng.uniform(-5000, 5000, size=(n, 2))
    return np.ascontiguousarray(np.round(xy) if integer else xy)


@pytest.mark.parametrize("n,integer,seed", [(5, True, 1), (17, False, 2), (200, True, 3), (333, False, 4), (700, False, 5)])
def test_costs_nn_descent_random(oracle, reference, n, integer, seed):
    rng = np.random.default_rng(seed)
    xy = _rand_instance(rng, n, integer)
    reference.set_points(xy)
    assert np.array_equal(oracle.cost_matrix(xy), reference.costs())
    for start in sorted(set([0, n // 2, n - 1])):
        e_r, p_r, c_r = reference.nn(start)
        e_o, p_o, c_o = oracle.nn(xy, start)
        assert e_r == e_o == 0 and np.array_equal(p_r, p_o) and c_r == c_o
        m_r, q_r, d_r, la, lb, ld = reference.two_opt(p_r)
        m_o, q_o, d_o, oa, ob, od = oracle.two_opt(xy, p_o)
        assert m_r == m_o and d_r == d_o and np.array_equal(q_r, q_o)
        assert np.array_equal(la, oa) and np.array_equal(lb, ob) and np.array_equal(ld, od)


def test_descent_from_random_tours(oracle, reference):
    """Random (non-NN) tours exercise wrap-around segments and long reversals."""
    rng = np.random.default_rng(11)
    xy = _rand_instance(rng, 150, True)
    reference.set_points(xy)
    for _ in range(3):
        perm = rng.permutation(150).astype(np.int32)
        path = np.empty(150, dtype=np.int32)
        path[perm] = np.roll(perm, -1)
        m_r, q_r, d_r, la, lb, ld = reference.two_opt(path)
        m_o, q_o, d_o, oa, ob, od = oracle.two_opt(xy, path)
        assert m_r == m_o and d_r == d_o and np.array_equal(q_r, q_o) and np.array_equal(la, oa) and np.array_equal(lb, ob)


@pytest.mark.parametrize("name,K", [("berlin52", 150), ("pcb442", 25)])
def test_tabu_moves(oracle, reference, name, K):
    xy, _ = load_instance(name)
    n = xy.shape[0]
    reference.set_points(xy)
    e, p, c = reference.nn(0)
    m, p, c, *_ = reference.two_opt(p)
    pr, po = p.copy(), p.copy()
    cr, co = c, c
    tr = np.full(n, -1, dtype=np.int32)
    to = tr.copy()
    ten_r, lo_r, hi_r = reference.tabu_init(n)
    assert (ten_r, lo_r, hi_r) == oracle.tabu_init(n)
    ten_o, inc_r, inc_o = ten_r, True, True
    for k in range(K):
        ten_r, inc_r = reference.tabu_linear_step(ten_r, lo_r, hi_r, inc_r)
        ten_o, inc_o = oracle.tabu_linear_step(ten_o, lo_r, hi_r, inc_o)
        assert (ten_r, inc_r) == (ten_o, inc_o)
        cr = reference.tabu_move(pr, cr, tr, ten_r, k)
        co, a, b, d = oracle.tabu_move(xy, po, co, to, ten_o, k)
        assert cr == co and np.array_equal(pr, po) and np.array_equal(tr, to), (name, k)


def test_vns_kick_and_descent(oracle, reference):
    """vns_kick (metaheuristic.c:344-409) under the same glibc rand() stream, then ref_2opt."""
    xy, _ = load_instance("kroA100")
    reference.set_points(xy)
    e, p, c = reference.nn(0)
    m, p, c, *_ = reference.two_opt(p)
    pr, po = p.copy(), p.copy()
    for it in range(12):
        reference.srand(1000 + it)
        reference.vns_kick(pr)
        reference.srand(1000 + it)          # same process, same libc: identical stream for the restatement
        oracle.vns_kick(po)
        assert np.array_equal(pr, po), it
        m_r, pr, c_r, *_ = reference.two_opt(pr)
        m_o, po, c_o, *_ = oracle.two_opt(xy, po, log=False)
        assert m_r == m_o and c_r == c_o and np.array_equal(pr, po)


def test_whole_algorithms_match_golden(reference, golden):
    """The reference's own entry points reproduce golden.json (guards the harness itself)."""
    xy, _ = load_instance("berlin52")
    reference.set_points(xy)
    e, p, c, s = reference.run(2)
    g = golden["berlin52"]["greedy_2opt"]
    assert (int(c), fnv1a(p), s) == (g["cost"], g["hash"], g["starting_node_field"])
    e, p, c, s = reference.run(3, 200)
    assert (int(c), fnv1a(p)) == (golden["berlin52"]["tabu_search_k200"]["cost"], golden["berlin52"]["tabu_search_k200"]["hash"])


@pytest.mark.parametrize("seed,n,symmetric,nc", [(1, 40, True, 0.0), (2, 97, True, 0.0), (3, 60, False, 0.0), (4, 50, True, 0.02)])
def test_foreign_matrix_paths(oracle, reference, seed, n, symmetric, nc):
    """h_greedyutil / ref_2opt with a caller-supplied matrix and h_Greedy_2opt_mod_costs
    (heuristics.c:118-149) on x*-weighted fractional matrices: tours and FP64 costs bit-equal."""
    from matrices import xstar_matrix
    M = xstar_matrix(seed, n, symmetric, nc)
    reference.random_points(n, 1)                 # only sets nnodes; the matrix is the caller's
    for start in (0, n // 2, n - 1):
        er, pr, cr = reference.nn_costs(M, start)
        eo, po, co = oracle.nn_matrix(M, start)
        assert er == eo == 0 and cr == co and np.array_equal(pr, po), start
        # with -1.0 entries NN may close early (not a tour); on an asymmetric matrix the reference's 2-opt
        # has no potential function (reversed segments change direction) and need not terminate
        if nc == 0.0 and symmetric:
            e2, p2, c2 = reference.ref_2opt_costs(M, pr)
            m, p3, c3, *_ = oracle.two_opt_matrix(M, po)
            assert e2 == 0 and c2 == c3 and np.array_equal(p2, p3), start
    if nc == 0.0 and symmetric:
        e, pm, cm = reference.mod_costs(M)
        po, co = oracle.mod_costs(M)
        assert e == 0 and cm == co and np.array_equal(pm, po)
        e1, p1, c1 = reference.nn_costs(M, n - 1)
        e2, p2, c2 = reference.ref_2opt_costs(M, p1)
        assert c2 == cm and np.array_equal(p2, pm)      # = the LAST start's result (nothing is kept between starts)


def test_extra_mileage_matches_reference(oracle, reference):
    for name in ("berlin52", "kroA100", "pcb442"):
        xy, _ = load_instance(name)
        n = xy.shape[0]
        reference.set_points(xy)
        e, p, c, s_ = reference.run(5)
        po, co = oracle.extra_mileage(xy)
        assert e == 0 and c == co and np.array_equal(p, po), name
        for a, b in ((0, 1), (n - 1, 2), (5, 4)):
            e, p, c = reference.extra_mileage_from(a, b)
            po, co = oracle.extra_mileage(xy, a, b)
            assert c == co and np.array_equal(p, po), (name, a, b)


@pytest.mark.parametrize("n,integer,ncomp,local,seed", [(9, True, 3, False, 1), (60, False, 20, False, 2), (300, True, 12, True, 3),
                                                        (500, False, 150, False, 4), (900, False, 9, True, 5)])
def test_patching_step_by_step(oracle, reference, n, integer, ncomp, local, seed):
    """cx_patching (cplex_model.c:693-755; the reference's file compiled unchanged with a stub <cplex.h>) against
    the restatement, patch by patch: successor array, component labels, ncomp and cost after every call."""
    from components import cycles_solution, solution_cost
    rng = np.random.default_rng(seed)
    xy = _rand_instance(rng, n, integer)
    if integer:
        xy = np.ascontiguousarray(np.round(xy / 100.0))       # coarse lattice: many equal deltas, the tie rule decides
    reference.set_points(xy)
    path, comp, nc = cycles_solution(seed, n, ncomp, local, xy)
    cost = solution_cost(oracle, xy, path)
    p2, c2, n2, k2 = path.copy(), comp.copy(), nc, cost
    while nc > 1:
        ok, nc, cost, a, b, d = oracle.patching(xy, path, comp, nc, cost)
        n2, k2 = reference.patching(p2, c2, n2, k2)
        assert ok and nc == n2 and cost == k2 and np.array_equal(path, p2) and np.array_equal(comp, c2), (a, b, d)
    assert oracle.is_tour(path)


@pytest.mark.parametrize("n", [7, 8, 9, 11, 12, 13, 16, 21])
def test_vns_kick_small_n(oracle, reference, n):
    """vns_kick's rejection test reads tour[-1] / tour[n] out of bounds (metaheuristic.c:372); under glibc both are 0
    unless n = 2 (mod 4), so a draw of 0 is rejected once index n-1 has been accepted.  Frequent on small n."""
    xy = np.ascontiguousarray(np.random.default_rng(n).uniform(0, 100, (n, 2)))
    reference.set_points(xy)
    for seed in range(300):
        p = np.roll(np.arange(n), -1).astype(np.int32)
        q = p.copy()
        reference.srand(seed); reference.vns_kick(p)
        reference.srand(seed); oracle.vns_kick(q)
        assert np.array_equal(p, q), seed
