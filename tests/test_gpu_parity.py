"""GPU parity tests (-m gpu): libtspcuda.so through its C ABI against the CPU oracle on the same
seeded inputs, against the golden fixtures generated from the unmodified reference objects
(tests/golden/golden.json), and -- at BASELINE.json's full sizes -- through the move-log audit
and size-independent properties.  Bar: bit-exact integer tours (successor arrays), identical move
logs (a, b, delta), cost equal (integer-valued doubles; 1e-9 relative is trivially met)."""
import numpy as np
import pytest

from oracle.bindings import ATT, CEIL_2D, EUC_REF, GEO, fnv1a, load_instance

pytestmark = pytest.mark.gpu


def _rand_xy(seed, n, integer, lo=-5000, hi=5000):
    rng = np.random.default_rng(seed)
    xy = rng.uniform(lo, hi, size=(n, 2))
    return np.ascontiguousarray(np.round(xy) if integer else xy)


def _rand_tour(seed, n):
    rng = np.random.default_rng(seed)
    perm = rng.permutation(n).astype(np.int32)
    path = np.empty(n, dtype=np.int32)
    path[perm] = np.roll(perm, -1)
    return path


SCANS = ["full", "pruned"]


def _inst(tsp, xy, scan, wt=EUC_REF):
    """Instance with the 2-opt scan kernel pinned: 'full' = every pair as the reference (descend_kernel),
    'pruned' = exact candidate pruning (descend_pruned_kernel), 'auto' = the library's own choice."""
    I = tsp.Instance(xy, wt)
    I.set_scan_mode({"auto": tsp.SCAN_AUTO, "full": tsp.SCAN_FULL, "pruned": tsp.SCAN_PRUNED}[scan])
    return I


def _log3(la, lb, ld):
    return [(int(a), int(b), int(d)) for a, b, d in zip(la, lb, ld)]


# ---- arithmetic -----------------------------------------------------------------------------
def test_device_sqrt_is_ieee(tsp):
    """The in-kernel float sqrt equals sqrt.rn.f32 on every float in [2^-60, 2^64)."""
    with tsp.Instance(np.zeros((4, 2))) as I:
        assert I.selfcheck_sqrt() == 0


@pytest.mark.parametrize("name", ["berlin52", "u1060", "fl1577", "pr2392"])
def test_cost_matrix_instances(tsp, oracle, name):
    xy, _ = load_instance(name)
    with tsp.Instance(xy) as I:
        M = I.cost_matrix()
    assert np.array_equal(M, oracle.cost_matrix(xy))


@pytest.mark.parametrize("n,integer", [(1, True), (2, False), (3, True), (7, False), (64, True), (129, False), (1000, False), (1001, True)])
def test_cost_matrix_random(tsp, oracle, n, integer):
    xy = _rand_xy(100 + n, n, integer)
    with tsp.Instance(xy) as I:
        M = I.cost_matrix()
        ms = I.cost_matrix_device()
        assert ms >= 0 and I.cost_matrix_checksum() == int(M.sum())
    E = oracle.cost_matrix(xy)
    assert np.array_equal(M, E) and np.all(np.diag(M) == -1.0)


@pytest.mark.parametrize("n,integer", [(1030, True), (2049, False), (2055, True), (4104, False)])
def test_cost_matrix_unit_shapes(tsp, oracle, monkeypatch, n, integer):
    """K1's work units (8 rows x 1024 columns, one per CTA, kernels.cuh): several column chunks per row, a last row group
    with fewer than 8 rows, odd n (row groups of every other row, unpaired first / last cells), the diagonal in a partial
    trip -- against the oracle cell by cell, and the same matrix from the round-1 launch shape (units dealt out grid-stride,
    TSPCUDA_K1_CTAS) and from whole-row units (TSPCUDA_K1_CHUNK=0)."""
    xy = _rand_xy(300 + n, n, integer)
    E = oracle.cost_matrix(xy)
    with tsp.Instance(xy) as I:
        M = I.cost_matrix()
    assert np.array_equal(M, E) and np.all(np.diag(M) == -1.0)
    for key, val in (("TSPCUDA_K1_CTAS", "3"), ("TSPCUDA_K1_CHUNK", "0"), ("TSPCUDA_K1_CHUNK", "1")):
        monkeypatch.setenv(key, val)
        with tsp.Instance(xy) as I:
            assert np.array_equal(I.cost_matrix(), E)
        monkeypatch.delenv(key)


def test_cost_matrix_large_coordinates(tsp, oracle):
    """Magnitudes of usa13509 / pla85900 (where float-sqrt rounding differs most from TSPLIB)."""
    xy = _rand_xy(5, 600, False, 245552.778, 1244961.111)
    with tsp.Instance(xy) as I:
        assert np.array_equal(I.cost_matrix(), oracle.cost_matrix(xy))
    xy = _rand_xy(6, 600, True, 548000, 1452000)
    with tsp.Instance(xy) as I:
        assert np.array_equal(I.cost_matrix(), oracle.cost_matrix(xy))


def test_usa13509_matrix_checksum(tsp, oracle):
    """Full-size K1: 1.46 GB matrix stays in HBM; its checksum equals the oracle's sum of a sample of
    rows scaled is not possible, so compare the exact total against row sums from the oracle."""
    xy, _ = load_instance("usa13509")
    with tsp.Instance(xy) as I:
        I.cost_matrix_device()
        total = I.cost_matrix_checksum()
    E = oracle.cost_matrix(xy)
    assert total == int(E.sum())


# ---- nearest neighbour ----------------------------------------------------------------------
@pytest.mark.parametrize("name", ["berlin52", "pcb442", "u1060", "rl1304"])
def test_nn_instances(tsp, oracle, golden, name):
    xy, _ = load_instance(name)
    n = xy.shape[0]
    with tsp.Instance(xy) as I:
        code, p, c = I.nn(0)
        assert code == 0 and int(c) == golden[name]["nn0"]["cost"] and fnv1a(p) == golden[name]["nn0"]["hash"]
        for s in (1, n // 3, n - 1):
            code, p, c = I.nn(s)
            e, po, co = oracle.nn(xy, s)
            assert code == 0 and np.array_equal(p, po) and c == co
        assert I.nn(n)[0] == tsp.UNAVAILABLE and I.nn(-1)[0] == tsp.UNAVAILABLE      # heuristics.c:223-226


@pytest.mark.parametrize("name", ["berlin52", "kroA100", "pcb442", "pr1002"])
def test_all_start_nn(tsp, golden, name):
    xy, _ = load_instance(name)
    g = golden[name]["allnn"]
    with tsp.Instance(xy) as I:
        bs, bp, bc, costs = I.nn_multi(0, xy.shape[0])
    assert (int(bc), bs, fnv1a(bp)) == (g["cost"], g["start"], g["hash"])
    if "greedy_2opt" in golden[name]:
        assert [int(c) for c in costs] == [r[0] for r in golden[name]["greedy_2opt"]["per_start"]]


def test_nn_ties_and_duplicates(tsp, oracle):
    xy = np.array([[0.0, 0.0], [5.0, 5.0], [0.0, 0.0], [5.0, 5.0], [9.0, 1.0], [0.0, 0.0]])
    with tsp.Instance(xy) as I:
        for s in range(6):
            code, p, c = I.nn(s)
            e, po, co = oracle.nn(xy, s)
            assert np.array_equal(p, po) and c == co
    # lattice: massive distance ties, lowest node id must win every step
    g = np.stack(np.meshgrid(np.arange(12.0), np.arange(11.0)), -1).reshape(-1, 2)
    with tsp.Instance(g) as I:
        for s in (0, 17, 131):
            code, p, c = I.nn(s)
            e, po, co = oracle.nn(g, s)
            assert np.array_equal(p, po) and c == co


# ---- 2-opt ----------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["berlin52", "eil51", "kroA100", "pcb442", "pr1002", "u1060", "rl1304", "fl1577", "pr2392"])
@pytest.mark.parametrize("scan", SCANS)
def test_descent_golden(tsp, oracle, golden, name, scan):
    xy, _ = load_instance(name)
    g = golden[name]["nn0_2opt"]
    e, p, c = oracle.nn(xy, 0)
    with _inst(tsp, xy, scan) as I:
        code, moves, p2, c2, log = I.two_opt(p, log_cap=64)
    assert code == 0 and moves == g["moves"] and int(c2) == g["cost"] and fnv1a(p2) == g["hash"]
    assert [list(x) for x in log[:len(g["first_moves"])]] == g["first_moves"]


@pytest.mark.parametrize("n", [4, 5, 6, 9, 31, 126, 127, 128, 129, 130, 131, 254, 255, 256, 257, 383, 509, 777])
@pytest.mark.parametrize("integer", [True, False])
@pytest.mark.parametrize("scan", SCANS)
def test_descent_random_vs_oracle(tsp, oracle, n, integer, scan):
    """Sizes straddle the 127-row strip and 32-column block boundaries of the scan kernel."""
    xy = _rand_xy(n * 2 + integer, n, integer)
    path = _rand_tour(n, n)                     # random tours: long and wrap-around reversals
    m_o, p_o, c_o, la, lb, ld = oracle.two_opt(xy, path)
    with _inst(tsp, xy, scan) as I:
        code, moves, p2, c2, log = I.two_opt(path, log_cap=len(la) + 8)
    assert code == 0 and moves == m_o and c2 == c_o
    assert log == _log3(la, lb, ld)
    assert np.array_equal(p2, p_o)


@pytest.mark.parametrize("force", ["", "general"])
@pytest.mark.parametrize("dups", [False, True])
@pytest.mark.parametrize("scan", SCANS)
def test_metric_variants_agree(tsp, oracle, monkeypatch, force, dups, scan):
    """Both implementations of tsp.c:629 (integer coordinates: FMA + 2^-60 addend, no clamp; general:
    separate roundings + clamp) must give the same move log as the oracle -- with and without
    coincident nodes (zero-cost edges)."""
    n = 700
    xy = _rand_xy(91, n, True, lo=-300000, hi=900000)
    if dups:
        xy[5] = xy[400]; xy[6] = xy[400]; xy[699] = xy[0]
    if force:
        monkeypatch.setenv("TSPCUDA_FORCE_METRIC", force)
    path = _rand_tour(17, n)
    m_o, p_o, c_o, la, lb, ld = oracle.two_opt(xy, path)
    with _inst(tsp, xy, scan) as I:
        code, moves, p2, c2, log = I.two_opt(path, log_cap=len(la) + 8)
        assert np.array_equal(I.cost_matrix(), oracle.cost_matrix(xy))
    assert code == 0 and moves == m_o and c2 == c_o
    assert log == _log3(la, lb, ld)
    assert np.array_equal(p2, p_o)


@pytest.mark.parametrize("scan", SCANS + ["auto"])
def test_descent_lattice_ties(tsp, oracle, scan):
    """Integer lattice: many equal deltas; the lexicographically smallest (a,b) must win each scan."""
    g = np.stack(np.meshgrid(np.arange(20.0), np.arange(15.0)), -1).reshape(-1, 2) * 10.0
    for seed in (1, 2):
        path = _rand_tour(seed, g.shape[0])
        m_o, p_o, c_o, la, lb, ld = oracle.two_opt(g, path)
        with _inst(tsp, g, scan) as I:
            code, moves, p2, c2, log = I.two_opt(path, log_cap=len(la) + 8)
        assert log == _log3(la, lb, ld) and np.array_equal(p2, p_o) and c2 == c_o


def test_two_opt_once_steps(tsp, oracle):
    """ref_2opt_once semantics: one scan+move per call, returns best_delta, 0 at the local optimum."""
    xy, _ = load_instance("kroA100")
    e, p, c = oracle.nn(xy, 0)
    po, co = p.copy(), c
    pg, cg = p.copy(), c
    with tsp.Instance(xy) as I:
        for _ in range(40):
            d_o, co, a_o, b_o = oracle.two_opt_once(xy, po, co)
            d_g, cg, (a, b) = I.two_opt_once(pg, cg)
            assert d_g == d_o and np.array_equal(pg, po) and cg == co
            if d_g == 0:
                assert (a, b) == (-1, -1)
            else:
                assert (a, b) == (a_o, b_o)
            if d_g == 0:
                break
        else:
            pytest.fail("no local optimum within 40 moves")


def test_small_and_degenerate(tsp, oracle):
    for n in (1, 2, 3):
        xy = np.array([[0.0, 0.0], [3.0, 4.0], [10.0, 0.0]])[:n].copy()
        with tsp.Instance(xy) as I:
            code, p, c = I.nn(0)
            e, po, co = oracle.nn(xy, 0)
            assert np.array_equal(p, po) and c == co
            code, moves, p2, c2, _ = I.two_opt(p)
            assert code == 0 and moves == 0 and np.array_equal(p2, p) and c2 == oracle.tour_cost(xy, p)
    # all points identical: every edge costs 0, no improving move
    xy = np.zeros((40, 2)) + 7.0
    with tsp.Instance(xy) as I:
        code, p, c = I.nn(5)
        e, po, co = oracle.nn(xy, 5)
        assert np.array_equal(p, po) and c == co == 0
        code, moves, p2, c2, _ = I.two_opt(p)
        assert moves == 0 and c2 == 0


def test_invalid_inputs(tsp):
    xy = _rand_xy(1, 50, True)
    with tsp.Instance(xy) as I:
        bad = np.arange(50, dtype=np.int32)            # self loops: not a tour
        with pytest.raises(tsp.TspCudaError) as ei:
            I.two_opt(bad)
        assert ei.value.code == tsp.INVALID_ARGUMENT
        two = np.roll(np.arange(50, dtype=np.int32), -1)
        two[24] = 0; two[49] = 25                      # two cycles of 25
        with pytest.raises(tsp.TspCudaError) as ei:
            I.two_opt(two)
        assert ei.value.code == tsp.INVALID_ARGUMENT
    with pytest.raises(tsp.TspCudaError) as ei:
        tsp.Instance(np.zeros((131073, 2)))
    assert ei.value.code == tsp.OUT_OF_RANGE
    with pytest.raises(tsp.TspCudaError) as ei:
        tsp.Instance(np.array([[0.0, 0.0], [9e6, 9e6], [1.0, 2.0], [5.0, 5.0]]))
    assert ei.value.code == tsp.OUT_OF_RANGE


def test_time_limit_returns_deadline_exceeded(tsp, oracle):
    """refinment.c:17-24: DEADLINE_EXCEEDED with a valid partial result; err_ok() treats it as success."""
    xy, _ = load_instance("d15112")
    e, p, c = oracle.nn(xy, 0)
    with tsp.Instance(xy) as I:
        code, moves, p2, c2, _ = I.two_opt(p, time_limit=0.0)
    assert code == tsp.DEADLINE_EXCEEDED and moves > 0 and oracle.is_tour(p2) and c2 == oracle.tour_cost(xy, p2) and c2 < c


# ---- tabu -----------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["berlin52", "kroA100", "pcb442"])
def test_tabu_steps_golden(tsp, oracle, golden, name):
    xy, _ = load_instance(name)
    n = xy.shape[0]
    g = golden[name]["tabu_steps"]
    e, p, c = oracle.nn(xy, 0)
    with tsp.Instance(xy) as I:
        code, moves, p, c, _ = I.two_opt(p)
        tenure, lo, hi = oracle.tabu_init(n)
        inc = True
        tl = np.full(n, -1, dtype=np.int32)
        for k, (g_tenure, g_cost, g_hash) in enumerate(g["log"]):
            tenure, inc = oracle.tabu_linear_step(tenure, lo, hi, inc)
            c, mv = I.tabu_move(p, c, tl, tenure, k)
            assert (tenure, int(c), fnv1a(p)) == (g_tenure, g_cost, g_hash), (name, k)
        assert fnv1a(tl) == g["tabu_list_hash"]


def test_tabu_vs_oracle_random(tsp, oracle):
    xy = _rand_xy(77, 300, False)
    path = _rand_tour(3, 300)
    m, p, c, *_ = oracle.two_opt(xy, path, log=False)
    po, pg, co, cg = p.copy(), p.copy(), c, c
    to = np.full(300, -1, dtype=np.int32)
    tg = to.copy()
    with tsp.Instance(xy) as I:
        for k in range(60):
            tenure = 20 + (k % 7)
            co, a, b, d = oracle.tabu_move(xy, po, co, to, tenure, k)
            cg, mv = I.tabu_move(pg, cg, tg, tenure, k)
            assert mv == (a, b, int(d)) and cg == co and np.array_equal(pg, po) and np.array_equal(tg, to), k
    # everything tabu: no admissible pair, nothing changes (metaheuristic.c:225 guard)
    with tsp.Instance(xy) as I:
        tl = np.zeros(300, dtype=np.int32)
        before = pg.copy()
        cg2, mv = I.tabu_move(pg, cg, tl, 1000, 5)
        assert mv[0] == -1 and cg2 == cg and np.array_equal(pg, before) and np.all(tl == 0)


@pytest.mark.parametrize("name", ["berlin52", "kroA100", "pcb442"])
def test_tabu_search_device_loop_golden(tsp, oracle, golden, name):
    """tspcuda_tabu_search (the k-loop of mh_TabuSearch in one device-resident run) against the step
    log generated from the unmodified reference objects: cost after every iteration, final tour,
    final tabu list; best tour = the first iteration reaching the minimum cost (tsp.c:671)."""
    xy, _ = load_instance(name)
    n = xy.shape[0]
    g = golden[name]["tabu_steps"]
    e, p, c = oracle.nn(xy, 0)
    tenure, lo, hi = oracle.tabu_init(n)
    inc, tenures = True, []
    for _ in g["log"]:
        tenure, inc = oracle.tabu_linear_step(tenure, lo, hi, inc)
        tenures.append(tenure)
    assert tenures == [t for t, _, _ in g["log"]]
    with tsp.Instance(xy) as I:
        code, moves, p, c, _ = I.two_opt(p)
        p_step, tl_step, c_step = p.copy(), np.full(n, -1, dtype=np.int32), c
        tl = np.full(n, -1, dtype=np.int32)
        cost, trace, best_iter, best_path, best_cost = I.tabu_search(p, c, tl, tenures)
        # the same iterations one call at a time (tabu_best_move semantics), for the best tour
        seen_best, seen_iter, seen_path = c, -1, None
        for k, t in enumerate(tenures):
            c_step, mv = I.tabu_move(p_step, c_step, tl_step, t, k)
            if c_step < seen_best:
                seen_best, seen_iter, seen_path = c_step, k, p_step.copy()
    assert [int(x) for x in trace] == [gc for _, gc, _ in g["log"]]
    assert int(cost) == g["log"][-1][1] and fnv1a(p) == g["log"][-1][2] and fnv1a(tl) == g["tabu_list_hash"]
    assert np.array_equal(p, p_step) and np.array_equal(tl, tl_step)
    assert best_iter == seen_iter and (best_iter < 0 or (best_cost == seen_best and np.array_equal(best_path, seen_path)))


def test_tabu_search_device_loop_vs_oracle(tsp, oracle):
    """Random instance, non-integer coordinates, 300 iterations split over two calls (first_iter
    continues), tenures that make every pair tabu for a while (iterations without a move)."""
    n = 400
    xy = _rand_xy(123, n, False)
    path = _rand_tour(9, n)
    m, p, c, *_ = oracle.two_opt(xy, path, log=False)
    tenures = [3 + (k * 7) % 90 for k in range(300)]
    tenures[50:60] = [100000] * 10
    po, co, to = p.copy(), c, np.full(n, -1, dtype=np.int32)
    trace_o = []
    for k, t in enumerate(tenures):
        co, a, b, d = oracle.tabu_move(xy, po, co, to, t, k)
        trace_o.append(co)
    pg, tg = p.copy(), np.full(n, -1, dtype=np.int32)
    with tsp.Instance(xy) as I:
        c1, tr1, bi1, bp1, bc1 = I.tabu_search(pg, c, tg, tenures[:170])
        c2, tr2, bi2, bp2, bc2 = I.tabu_search(pg, c1, tg, tenures[170:], first_iter=170, incumbent_cost=bc1 if bi1 >= 0 else c)
        # a per-call tabu move right after the device loop, same context, tour still resident (regression: the
        # host mirror of the device move counter must not drift)
        c3, mv3 = I.tabu_move(pg, c2, tg, 25, 300)
    co3, a3, b3, d3 = oracle.tabu_move(xy, po, co, to, 25, 300)
    assert mv3 == (a3, b3, int(d3)) and c3 == co3
    assert list(tr1) + list(tr2) == trace_o and c2 == co
    assert np.array_equal(pg, po) and np.array_equal(tg, to)
    best = min(trace_o)
    if best < c:
        assert min(bc1 if bi1 >= 0 else c, bc2 if bi2 >= 0 else c) == best


# ---- VNS local search (metaheuristic.c:279-320): host kicks, GPU descents ---------------------
def test_vns_loop(tsp, oracle):
    xy, _ = load_instance("kroA100")
    e, p, c = oracle.nn(xy, 0)
    po, pg = p.copy(), p.copy()
    import ctypes
    libc = ctypes.CDLL(None)
    with tsp.Instance(xy) as I:
        for it in range(10):
            m_o, po, c_o, *_ = oracle.two_opt(xy, po, log=False)
            code, moves, pg, c_g, _ = I.two_opt(pg)
            assert moves == m_o and c_g == c_o and np.array_equal(pg, po), it
            libc.srand(99 + it); oracle.vns_kick(po)
            libc.srand(99 + it); oracle.vns_kick(pg)      # the kick is host code in the reference (unchanged driver)
            assert np.array_equal(po, pg)


@pytest.mark.parametrize("name,k,seed", [("berlin52", 40, 1), ("eil51", 25, 7), ("kroA100", 20, 3), ("pcb442", 10, 5), ("rl1304", 5, 2)])
def test_vns_search_device_resident_loop(tsp, oracle, name, k, seed):
    """tspcuda_vns_search (the loop of mh_VNS, metaheuristic.c:277-321, tour resident on the device) against the same
    loop made of the oracle's ref_2opt and its vns_kick (pinned on the reference objects), same glibc rand() stream:
    cost trace, incumbent and the final kicked tour."""
    import ctypes
    libc = ctypes.CDLL(None)
    xy, _ = load_instance(name)
    e, p0, c0 = oracle.nn(xy, 0)
    libc.srand(seed)
    po, best, best_path, bi, trace = p0.copy(), c0, None, -1, []
    for i in range(k):
        m, po, co, *_ = oracle.two_opt(xy, po, log=False)
        if co < best:
            best, best_path, bi = co, po.copy(), i
        trace.append(co)
        for _ in range(libc.rand() % 9 - 2):
            oracle.vns_kick(po)
    libc.srand(seed)
    with tsp.Instance(xy) as I:
        code, pg, cg, tr, gbi, gbp, gbc, done = I.vns_search(p0, k, c0)
    assert code == 0 and done == k and tr.tolist() == trace and cg == trace[-1]
    assert np.array_equal(pg, po)
    assert (gbi, gbc) == (bi, best) and (bi < 0 or np.array_equal(gbp, best_path))


def test_vns_restarts_shard_independent(tsp):
    """multistart.vns_restarts: a restart depends on (instance, start node, seed) only, so the merged best of two shards
    equals the best of one shard holding every start."""
    from travellingsalesmanoptimization_b200.multistart import shard_starts, vns_restarts
    xy, _ = load_instance("kroA100")
    with tsp.Instance(xy) as I:
        whole = vns_restarts(I, 0, 6, 1, 8, seed=5)
        parts = [vns_restarts(I, *shard_starts(0, 6, r, 2), 8, seed=5) for r in range(2)]
    merged = min((p for p in parts if p[0] is not None), key=lambda p: (p[0], p[1]))
    assert whole[0] == merged[0] and whole[1] == merged[1] and np.array_equal(whole[2], merged[2])


def test_vns_search_limits(tsp, oracle):
    xy, _ = load_instance("kroA100")
    e, p0, c0 = oracle.nn(xy, 0)
    with tsp.Instance(xy) as I:
        code, pg, cg, tr, bi, bp, bc, done = I.vns_search(p0, 0, c0)                 # k = 0: nothing happens
        assert code == 0 and done == 0 and np.array_equal(pg, p0) and bi == -1 and bc == c0
        code, pg, cg, tr, bi, bp, bc, done = I.vns_search(p0, 10**6, c0, time_limit=0.05)
        assert code == tsp.DEADLINE_EXCEEDED and 0 < done < 10**6 and oracle.is_tour(pg) and bc <= c0
        code, pg, cg, tr, bi, bp, bc, done = I.vns_search(p0, 3, 0.0)                # nothing beats a cost of 0
        assert bi == -1 and bp is None and bc == 0.0 and done == 3


# ---- TSPLIB95 weight types (no reference behaviour: parity unpinned, oracle = TSPLIB95) ---------
def test_ceil2d_and_att(tsp, oracle):
    xy, _ = load_instance("pla7397")
    xy = np.ascontiguousarray(xy[:900])
    with tsp.Instance(xy, tsp.CEIL_2D) as I:
        assert np.array_equal(I.cost_matrix(), oracle.cost_matrix(xy, CEIL_2D))
        e, po, co = oracle.nn(xy, 0, CEIL_2D)
        code, p, c = I.nn(0)
        assert np.array_equal(p, po) and c == co
        m_o, p_o, c_o, la, lb, ld = oracle.two_opt(xy, po, CEIL_2D)
        code, moves, p2, c2, log = I.two_opt(p, log_cap=len(la) + 4)
        assert log == _log3(la, lb, ld) and np.array_equal(p2, p_o) and c2 == c_o
    xy, _ = load_instance("att532")
    with tsp.Instance(xy, tsp.ATT) as I:
        assert np.array_equal(I.cost_matrix(), oracle.cost_matrix(xy, ATT))
        e, po, co = oracle.nn(xy, 0, ATT)
        m_o, p_o, c_o, la, lb, ld = oracle.two_opt(xy, po, ATT)
        code, moves, p2, c2, log = I.two_opt(po, log_cap=len(la) + 4)
        assert log == _log3(la, lb, ld) and np.array_equal(p2, p_o) and c2 == c_o


@pytest.mark.parametrize("name", ["burma14", "ulysses16", "gr666"])
def test_geo(tsp, oracle, name):
    """TSPLIB95 GEO on the device (cos/acos from CUDA's libdevice) against the oracle (glibc): cost
    matrix, NN, descent, all-start multi-start; gr666's canonical tour 1..n costs 423710 (TSPLIB95 KAT).
    Parity unpinned: the reference rejects GEO files (tsp.c:576-583)."""
    xy, ewt = load_instance(name)
    assert ewt == "GEO"
    n = xy.shape[0]
    with tsp.Instance(xy, tsp.GEO) as I:
        assert np.array_equal(I.cost_matrix(), oracle.cost_matrix(xy, GEO))
        e, po, co = oracle.nn(xy, 0, GEO)
        code, p, c = I.nn(0)
        assert np.array_equal(p, po) and c == co
        m_o, p_o, c_o, la, lb, ld = oracle.two_opt(xy, po, GEO)
        code, moves, p2, c2, log = I.two_opt(p, log_cap=len(la) + 4)
        assert log == _log3(la, lb, ld) and np.array_equal(p2, p_o) and c2 == c_o
        if name == "gr666":
            canon = np.roll(np.arange(n, dtype=np.int32), -1)
            code, mv, pc, cc, _ = I.two_opt(canon, max_moves=0)
            assert cc == 423710.0
        else:
            bs, bp, bc, nn_costs, final, mvs = I.multistart_two_opt(0, n)
            per = [oracle.two_opt(xy, oracle.nn(xy, s, GEO)[1], GEO, log=False) for s in range(n)]
            assert [int(f) for f in final] == [int(r[2]) for r in per]


# ---- multi-start (h_greedy_2opt, heuristics.c:74-116) -----------------------------------------
@pytest.mark.parametrize("name", ["berlin52", "eil51", "kroA100"])
@pytest.mark.parametrize("scan", ["full", "auto"])
def test_multistart_golden(tsp, golden, name, scan):
    """All starts: 'full' = K6 (CTA per descent, every pair), 'auto' = K8 (CTA per descent, pruned scan)."""
    xy, _ = load_instance(name)
    g = golden[name]["greedy_2opt"]
    with _inst(tsp, xy, scan) as I:
        bs, bp, bc, nn_costs, final, moves = I.multistart_two_opt(0, xy.shape[0])
    assert [[int(a), int(m), int(f)] for a, m, f in zip(nn_costs, moves, final)] == g["per_start"]
    assert (int(bc), bs, fnv1a(bp)) == (g["cost"], g["first_start"], g["hash"])


def test_multistart_pr2392_first8_and_strided(tsp, golden):
    xy, _ = load_instance("pr2392")
    g = golden["pr2392"]["greedy_2opt_first8"]
    with tsp.Instance(xy) as I:
        bs, bp, bc, nn_costs, final, moves = I.multistart_two_opt(0, 8)
        assert [[int(a), int(m), int(f)] for a, m, f in zip(nn_costs, moves, final)] == [r[:3] for r in g]
        best = min(r[2] for r in g)
        first = [r[2] for r in g].index(best)
        assert int(bc) == best and bs == first and fnv1a(bp) == g[first][3]
        # block-cyclic shard of rank 1 of 4 (starts 1, 5): what one GPU of a 4-GPU run computes
        bs2, bp2, bc2, nn2, fin2, mv2 = I.multistart_two_opt(1, 2, 4)
        assert [int(f) for f in fin2] == [g[1][2], g[5][2]]


@pytest.mark.parametrize("scan", ["full", "auto"])
def test_multistart_cta_paths_vs_grid_path(tsp, oracle, scan):
    """The CTA-per-descent kernels (K6: full scan, n <= 7600, 512- and 1024-thread variants; K8: pruned
    scan, any n, used from num_sms/4 starts up) and the grid-wide descent kernels must produce the same
    tours: compare multistart with nn + two_opt start by start, and one start of each size with the
    CPU oracle."""
    for n, starts, integer in ((130, 6, True), (131, 6, False), (1500, 4, False), (3700, 3, True), (900, 60, False), (901, 60, True),
                               (9000, 40, False), (9001, 40, True)):
        if scan == "full" and n >= 9000:
            continue
        xy = _rand_xy(900 + n, n, integer)                  # integer: EucInt kernels, otherwise EucGeneral (no FMA contraction)
        with _inst(tsp, xy, scan) as I:
            bs, bp, bc, nn_costs, final, moves = I.multistart_two_opt(2, starts, 5)
            per = []
            I.set_scan_mode(tsp.SCAN_FULL if n < 5000 else tsp.SCAN_AUTO)
            for k in range(starts):
                code, p, c = I.nn(2 + 5 * k)
                code, mv, p2, c2, _ = I.two_opt(p)
                per.append((c, mv, c2, p2))
            assert [(a, m, f) for a, m, f in zip(nn_costs, moves, final)] == [(c, mv, c2) for c, mv, c2, _ in per]
            best = min(range(starts), key=lambda k: (per[k][2], k))
            assert bs == 2 + 5 * best and bc == per[best][2] and np.array_equal(bp, per[best][3])
        e, po, co = oracle.nn(xy, 2)
        m_o, p_o, c_o, *_ = oracle.two_opt(xy, po, log=False)
        assert (co, m_o, c_o) == per[0][:3] and np.array_equal(p_o, per[0][3])


@pytest.mark.parametrize("n,starts,integer", [(700, 5, True), (700, 5, False), (8500, 9, True), (8500, 9, False)])
def test_multistart_cluster_sizes_agree(tsp, oracle, monkeypatch, n, starts, integer):
    """K8 with a thread-block cluster per descent (2/4/8 SMs share one descent, pruned.cuh) against one CTA per
    descent (1) and against K7 with one start after the other (0): same per-start costs, move counts and
    best tour; the first start additionally against the CPU oracle.  n = 700: scans of a few microseconds,
    i.e. the cluster barriers and the per-CTA argmin keys under the tightest timing."""
    xy = _rand_xy(4100 + n, n, integer)
    ref = None
    for cl in (0, 1, 2, 4, 8):
        monkeypatch.setenv("TSPCUDA_K8_CLUSTER", str(cl))
        with tsp.Instance(xy) as I:
            bs, bp, bc, nn_costs, final, moves = I.multistart_two_opt(1, starts, 3)
        key = (bs, int(bc), bp.tobytes(), nn_costs.tobytes(), final.tobytes(), moves.tobytes())
        if ref is None:
            ref = key
            e, po, co = oracle.nn(xy, 1)
            m_o, p_o, c_o, *_ = oracle.two_opt(xy, po, log=False)
            assert (co, m_o, c_o) == (int(nn_costs[0]), int(moves[0]), int(final[0]))
        assert key == ref, cl


def test_pr2392_all_start_config2(tsp, golden):
    """Config 2: all 2392 starts NN + 2-opt on one B200.  The reference objects need 4.68 core-hours
    (BASELINE.md section 3): 393339, first reached at start 359; per-range bests for sharding checks."""
    xy, _ = load_instance("pr2392")
    with tsp.Instance(xy) as I:
        bs, bp, bc, nn_costs, final, moves = I.multistart_two_opt(0, 2392)
        bs_nn, bp_nn, bc_nn, _ = I.nn_multi(0, 2392)
        st = I.stats()
    assert st["pruned_scans"] == int(moves.sum()) + 2392        # K8 answered every scan
    assert (int(bc), bs) == (golden["pr2392"]["greedy_2opt_survey"]["cost"], golden["pr2392"]["greedy_2opt_survey"]["first_start"])
    for lo, hi, cost, first in golden["pr2392"]["greedy_2opt_survey_ranges"]:
        seg = final[lo:hi]
        assert int(seg.min()) == cost and lo + int(np.argmin(seg)) == first, (lo, hi)
    assert [[int(a), int(m), int(f)] for a, m, f in zip(nn_costs[:8], moves[:8], final[:8])] == \
        [r[:3] for r in golden["pr2392"]["greedy_2opt_first8"]]
    assert (int(bc_nn), bs_nn, fnv1a(bp_nn)) == (golden["pr2392"]["allnn"]["cost"], golden["pr2392"]["allnn"]["start"], golden["pr2392"]["allnn"]["hash"])


# ---- full-size configs ------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["usa13509", "d15112"])
def test_large_first_moves_golden(tsp, golden, name):
    xy, _ = load_instance(name)
    g = golden[name]
    with tsp.Instance(xy) as I:
        code, p, c = I.nn(0)
        assert int(c) == g["nn0"]["cost"] and fnv1a(p) == g["nn0"]["hash"]
        code, moves, p2, c2, log = I.two_opt(p, max_moves=20, log_cap=20)
    g2 = g["nn0_2opt_20"]
    assert moves == 20 and int(c2) == g2["cost"] and fnv1a(p2) == g2["hash"] and [list(x) for x in log] == g2["first_moves"]


@pytest.mark.parametrize("scan", SCANS + ["auto"])
def test_usa13509_full_descent(tsp, oracle, golden, scan):
    """Config 3: NN(0) tour to the 2-opt local optimum with on-the-fly FP64 distances.  The
    reference objects need 2567 s for this: 1983 moves -> 21119812."""
    xy, _ = load_instance("usa13509")
    with _inst(tsp, xy, scan) as I:
        code, p, c = I.nn(0)
        code, moves, p2, c2, log = I.two_opt(p, log_cap=4096)
    s = golden["usa13509"]["nn0_2opt_survey"]
    assert (moves, int(c2)) == (s["moves"], s["cost"])
    assert oracle.is_tour(p2) and oracle.tour_cost(xy, p2) == c2
    # audit: every logged delta re-derived on the CPU, full CPU scans at a few iterations, local optimality
    la = [x[0] for x in log]; lb = [x[1] for x in log]; ld = [float(x[2]) for x in log]
    r, pend = oracle.audit_moves(xy, p, la, lb, ld, [0, 1, 500, 1982])
    assert r == 0 and np.array_equal(pend, p2)
    d, a, b = oracle.two_opt_scan(xy, p2)
    assert d >= -1e-7                                   # local optimum (refinment.c:27)


@pytest.mark.parametrize("scan", SCANS)
def test_pla85900_descent_audit(tsp, oracle, scan):
    """Config 4 size (n = 85 900, 3.689e9 pairs per scan): the reference cannot run it (int overflow,
    tsp.c:614) so parity is the SURVEY 8c audit against the 64-bit oracle, plus invariants."""
    xy, _ = load_instance("pla85900")
    with _inst(tsp, xy, scan) as I:
        code, p, c = I.nn(0)
        assert code == 0 and oracle.is_tour(p) and oracle.tour_cost(xy, p) == c
        code, moves, p2, c2, log = I.two_opt(p, max_moves=48, log_cap=48)
    assert moves == 48 and oracle.is_tour(p2) and oracle.tour_cost(xy, p2) == c2
    la = [x[0] for x in log]; lb = [x[1] for x in log]; ld = [float(x[2]) for x in log]
    assert c2 == c + sum(ld) and all(d < 0 for d in ld)
    r, pend = oracle.audit_moves(xy, p, la, lb, ld, [0, 47])
    assert r == 0 and np.array_equal(pend, p2)


def test_pla85900_pruned_descent_matches_full_scan(tsp, oracle):
    """The whole pla85900 NN(0) descent through the pruned kernel: same first 64 moves as the full
    scan, and the end state the full-scan descent reaches (5589 moves -> 148543747, measured with
    descend_kernel in profiles/r01_bench_first.json); every logged delta is re-derived by the oracle,
    full CPU scans confirm the argmin at sampled iterations and the final tour is a local optimum."""
    xy, _ = load_instance("pla85900")
    with _inst(tsp, xy, "full") as I:
        code, p, c = I.nn(0)
        code, m64, pf, cf, log_full = I.two_opt(p, max_moves=64, log_cap=64)
    with _inst(tsp, xy, "pruned") as I:
        code, moves, p2, c2, log = I.two_opt(p, log_cap=8192)
        st = I.stats()
    assert log[:64] == log_full
    assert (moves, int(c2)) == (5589, 148543747)
    assert st["pruned_scans"] == moves + 1 and st["pair_evals"] == 0
    la = [x[0] for x in log]; lb = [x[1] for x in log]; ld = [float(x[2]) for x in log]
    r, pend = oracle.audit_moves(xy, p, la, lb, ld, [2500, 5588])
    assert r == 0 and np.array_equal(pend, p2)
    d, a, b = oracle.two_opt_scan(xy, p2)
    assert d >= -1e-7


@pytest.mark.parametrize("wt", [CEIL_2D, ATT])
def test_pruned_scan_other_metrics(tsp, oracle, wt):
    """Search radius of the pruned scan under the TSPLIB95 metrics (parity unpinned, oracle only)."""
    xy = _rand_xy(5, 600, True, lo=0, hi=3000)
    path = _rand_tour(3, 600)
    m_o, p_o, c_o, la, lb, ld = oracle.two_opt(xy, path, wt)
    with _inst(tsp, xy, "pruned", wt) as I:
        code, moves, p2, c2, log = I.two_opt(path, log_cap=len(la) + 8)
    assert moves == m_o and c2 == c_o and log == _log3(la, lb, ld) and np.array_equal(p2, p_o)


def test_auto_scan_mode_random_tour(tsp, oracle):
    """AUTO starts on a random tour (every search box is huge: the full scan is faster), probes the
    pruned kernel and switches as the tour shortens; the result never depends on the choice."""
    n = 3000
    xy = _rand_xy(77, n, False)
    path = _rand_tour(5, n)
    m_o, p_o, c_o, la, lb, ld = oracle.two_opt(xy, path)
    with _inst(tsp, xy, "auto") as I:
        code, moves, p2, c2, log = I.two_opt(path, log_cap=len(la) + 8)
        st = I.stats()
    assert moves == m_o and c2 == c_o and log == _log3(la, lb, ld) and np.array_equal(p2, p_o)
    assert st["scans"] == moves + 1 and st["pruned_scans"] > 0


# ---- the `costs`-argument paths: caller-supplied matrices (heuristics.c:118-149) -----------------
def test_foreign_matrix_golden_gpu(tsp, oracle):
    """tspcuda_mx_* against tests/golden/golden_matrix.json (unmodified reference objects on x*-weighted
    fractional matrices): NN tours, 2-opt local optima and h_Greedy_2opt_mod_costs results; FP64 costs
    must be bit-equal (hex floats) -- the final costs are ~1e-13 rounding residues of zero-cost tours."""
    import json
    import os
    from matrices import xstar_matrix
    with open(os.path.join(os.path.dirname(__file__), "golden", "golden_matrix.json")) as f:
        G = json.load(f)
    for key, g in G.items():
        M = xstar_matrix(g["seed"], g["n"])
        with tsp.MatrixInstance(M) as I:
            for r in g["per_start"]:
                code, p, c, ln = I.nn(r["start"])
                assert code == 0 and ln == g["n"] and c.hex() == r["nn_cost"] and fnv1a(p) == r["nn_hash"], (key, r["start"])
                code, moves, p2, c2, _ = I.two_opt(p)
                assert code == 0 and c2.hex() == r["opt_cost"] and fnv1a(p2) == r["opt_hash"], (key, r["start"])
            pm, cm = I.mod_costs()
            assert cm.hex() == g["mod_costs"]["cost"] and fnv1a(pm) == g["mod_costs"]["hash"], key


@pytest.mark.parametrize("seed,n,symmetric,nc", [(21, 5, True, 0.0), (22, 33, True, 0.0), (23, 257, True, 0.0), (24, 1200, True, 0.0),
                                                 (25, 180, False, 0.0), (26, 90, True, 0.03)])
def test_foreign_matrix_vs_oracle(tsp, oracle, seed, n, symmetric, nc):
    """Move-for-move against the oracle: log (a, b, FP64 delta), tour, cost; asymmetric matrices (the
    entries are read at the reference's own indices; bounded number of moves, the descent has no
    potential function there); NOT_CONNECTED entries in the NN step; ref_2opt_once stepping."""
    from matrices import xstar_matrix
    M = xstar_matrix(seed, n, symmetric, nc)
    with tsp.MatrixInstance(M) as I:
        for start in sorted({0, n // 2, n - 1}):
            eo, po, co = oracle.nn_matrix(M, start)
            code, p, c, ln = I.nn(start)
            assert code == eo == 0 and c == co and np.array_equal(p, po), start
        if nc > 0:
            return
        path = _rand_tour(seed, n)
        cap = -1 if symmetric else 40
        m_o, p_o, c_o, la, lb, ld = oracle.two_opt_matrix(M, path, max_moves=cap)
        code, moves, p2, c2, log = I.two_opt(path, max_moves=cap, log_cap=len(la) + 8)
        assert moves == m_o and c2 == c_o and np.array_equal(p2, p_o)
        assert log == [(int(a), int(b), float(d)) for a, b, d in zip(la, lb, ld)]
        # ref_2opt_once: one scan per call
        ps, cs = path.copy(), oracle.two_opt_matrix(M, path, max_moves=0)[2]
        for k in range(min(6, len(ld))):
            d, cs = I.two_opt_once(ps, cs)
            assert d == ld[k]
        if symmetric:
            d, c_end = I.two_opt_once(p2.copy(), c2)
            assert d > -1e-7 and c_end == c2              # local optimum: nothing applied


# ---- Extra Mileage (heuristics.c:156-210, :290-367) -------------------------------
@pytest.mark.parametrize("name", ["berlin52", "eil51", "kroA100", "pcb442", "pr1002", "u1060", "rl1304", "fl1577", "pr2392"])
def test_extra_mileage_golden(tsp, name):
    """tests/golden/golden_em.json (unmodified reference objects; the EM_MAX values also are the published
    `ExtraMileage` column of results/heuristics-ric.csv: rl1304 307898, pr1002 302240, fl1577 26944)."""
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "golden_em.json")) as f:
        g = json.load(f)[name]
    xy, _ = load_instance(name)
    with tsp.Instance(xy) as I:
        p, c, ab = I.extra_mileage()
        assert (int(c), fnv1a(p)) == (g["em_max"]["cost"], g["em_max"]["hash"])
        for a, b, cost, h in g["from"]:
            p, c, ab = I.extra_mileage(a, b)
            assert ab == (a, b) and (int(c), fnv1a(p)) == (cost, h), (a, b)


@pytest.mark.parametrize("n,integer", [(2, True), (3, True), (4, False), (33, True), (257, False), (700, True)])
def test_extra_mileage_vs_oracle(tsp, oracle, n, integer):
    """Random instances, lattice ties, coincident nodes, the other metrics."""
    xy = _rand_xy(3 * n + integer, n, integer, lo=0, hi=60 if integer else 1000)      # small integer range: many ties
    if n >= 33:
        xy[7] = xy[20]
    for wt, twt in ((EUC_REF, tsp.EUC_REF), (CEIL_2D, tsp.CEIL_2D), (ATT, tsp.ATT)):
        with tsp.Instance(xy, twt) as I:
            p, c, ab = I.extra_mileage()
            po, co = oracle.extra_mileage(xy, wt=wt)
            assert c == co and np.array_equal(p, po), (n, wt)
            if n >= 4:
                p, c, ab = I.extra_mileage(n - 1, 1)
                po, co = oracle.extra_mileage(xy, n - 1, 1, wt)
                assert c == co and np.array_equal(p, po)


@pytest.mark.parametrize("scan", ["pruned", "auto"])
def test_pruned_scan_degenerate_geometry(tsp, oracle, scan):
    """Grids with one row / one column / one cell: collinear nodes, a single cluster of coincident nodes
    plus outliers, two far-apart clusters (empty cells in between)."""
    rng = np.random.default_rng(5)
    n = 300
    line_x = np.stack([np.round(rng.uniform(0, 5000, n)), np.full(n, 17.0)], -1)
    line_y = np.stack([np.full(n, -3.0), rng.uniform(0, 900, n)], -1)
    blob = np.zeros((n, 2)); blob[:10] = np.round(rng.uniform(0, 50, (10, 2)))
    two = np.round(np.concatenate([rng.uniform(0, 100, (n // 2, 2)), rng.uniform(1e6, 1e6 + 100, (n - n // 2, 2))]))
    for k, xy in enumerate((line_x, line_y, blob, two)):
        xy = np.ascontiguousarray(xy)
        path = _rand_tour(k, n)
        m_o, p_o, c_o, la, lb, ld = oracle.two_opt(xy, path)
        with _inst(tsp, xy, scan) as I:
            code, moves, p2, c2, log = I.two_opt(path, log_cap=len(la) + 8)
        assert moves == m_o and c2 == c_o and log == _log3(la, lb, ld) and np.array_equal(p2, p_o), k


@pytest.mark.parametrize("mode", ["grid", "list"])
def test_nn_kernels_agree(tsp, oracle, monkeypatch, mode):
    """K4 (every unvisited node per step) and K4b (grid squares, cost_floor stopping rule) must build the
    reference's tour: random / clustered / lattice (ties) instances against the oracle, several starts."""
    monkeypatch.setenv("TSPCUDA_NN", mode)
    rng = np.random.default_rng(11)
    clustered = np.round(np.concatenate([rng.normal(c, 30, (150, 2)) for c in ((0, 0), (5000, 200), (300, 9000), (7000, 7000))]))
    lattice = np.stack(np.meshgrid(np.arange(25.0), np.arange(20.0)), -1).reshape(-1, 2) * 7.0
    for xy in (_rand_xy(3, 900, False), np.ascontiguousarray(clustered), np.ascontiguousarray(lattice)):
        n = xy.shape[0]
        with tsp.Instance(xy) as I:
            for start in (0, n // 2, n - 1):
                e, po, co = oracle.nn(xy, start)
                code, p, c = I.nn(start)
                assert code == 0 and c == co and np.array_equal(p, po), (mode, n, start)
            bs, bp, bc, costs = I.nn_multi(0, 40)
            assert [int(x) for x in costs] == [int(oracle.nn(xy, s)[2]) for s in range(40)]
    for wt, twt in ((CEIL_2D, tsp.CEIL_2D), (ATT, tsp.ATT)):
        xy = _rand_xy(8, 700, True, lo=0, hi=4000)
        with tsp.Instance(xy, twt) as I:
            e, po, co = oracle.nn(xy, 5, wt)
            code, p, c = I.nn(5)
            assert c == co and np.array_equal(p, po), (mode, wt)


def test_maximum_size(tsp, oracle):
    """n = 131072 (TSPCUDA_MAX_NODES: 17-bit node ids in the packed key): NN over the grid, the first 2-opt moves
    by both scan kernels, audited by the oracle (every logged delta; one full CPU scan, 8.6e9 pairs); one more
    node is refused with OUT_OF_RANGE."""
    n = 131072
    rng = np.random.default_rng(2024)
    xy = np.ascontiguousarray(np.round(rng.uniform(0, 2_000_000, size=(n, 2))))
    with _inst(tsp, xy, "full") as I:
        code, p, c = I.nn(n - 1)
        assert code == 0 and oracle.is_tour(p) and oracle.tour_cost(xy, p) == c
        code, mv, p_full, c_full, log_full = I.two_opt(p, max_moves=12, log_cap=12)
        I.set_scan_mode(tsp.SCAN_PRUNED)
        code, mv2, p_pr, c_pr, log_pr = I.two_opt(p, max_moves=12, log_cap=12)
    assert mv == mv2 == 12 and log_full == log_pr and c_full == c_pr and np.array_equal(p_full, p_pr)
    la = [x[0] for x in log_full]; lb = [x[1] for x in log_full]; ld = [float(x[2]) for x in log_full]
    r, pend = oracle.audit_moves(xy, p, la, lb, ld, [0])
    assert r == 0 and np.array_equal(pend, p_full)
    with pytest.raises(tsp.TspCudaError) as ei:
        tsp.Instance(np.zeros((n + 1, 2)))
    assert ei.value.code == tsp.OUT_OF_RANGE


@pytest.mark.parametrize("kind", ["random", "lattice", "mostly_tabu"])
def test_tabu_search_pruned_loop_matches_per_call_path(tsp, oracle, kind):
    """n >= 4096: tspcuda_tabu_search runs K5c (pruned scan with a bound, falling back to the full-scan loop when the
    bound is exhausted).  It must reproduce, iteration by iteration, what tabu_best_move gives through
    tspcuda_tabu_move (full scan, one call per iteration): cost trace, tour, tabu list; the first iterations are
    also checked against the oracle."""
    if kind == "lattice":
        xy = np.ascontiguousarray(np.stack(np.meshgrid(np.arange(70.0), np.arange(64.0)), -1).reshape(-1, 2) * 5.0)
    else:
        xy = _rand_xy(41, 5000, kind == "mostly_tabu")
    n = xy.shape[0]
    K = 260
    tenures = [40 + (k * 13) % 200 for k in range(K)]
    if kind == "mostly_tabu":
        tenures = [n] * K                               # nothing ever leaves the list: admissible pairs run out
    with tsp.Instance(xy) as I:
        code, p, c = I.nn(3)
        code, mv, p, c, _ = I.two_opt(p)
        pa, ta, ca, trace_a = p.copy(), np.full(n, -1, dtype=np.int32), c, []
        for k, t in enumerate(tenures):
            ca, mvk = I.tabu_move(pa, ca, ta, t, k)
            trace_a.append(ca)
        pb, tb = p.copy(), np.full(n, -1, dtype=np.int32)
        I.reset_stats()
        cb, trace_b, bi, bp, bc = I.tabu_search(pb, c, tb, tenures)
        st = I.stats()
    assert list(trace_b) == trace_a and cb == ca
    assert np.array_equal(pa, pb) and np.array_equal(ta, tb)
    assert st["pruned_scans"] > 0
    best = min(trace_a)
    assert (bi >= 0) == (best < c) and (bi < 0 or (bc == best and bi == trace_a.index(best)))
    po, to, co = p.copy(), np.full(n, -1, dtype=np.int32), c
    for k in range(12):
        co, a, b, d = oracle.tabu_move(xy, po, co, to, tenures[k], k)
        assert co == trace_a[k], k


def test_tabu_search_pruned_vs_full_loop_pla85900(tsp):
    """K5c against K5b at the headline size (more tour positions than phase-1 threads: several trips per thread)."""
    xy, _ = load_instance("pla85900")
    n = xy.shape[0]
    tenures = [30 + (k * 7) % 50 for k in range(40)]
    with tsp.Instance(xy) as I:
        code, p, c = I.nn(0)
        code, mv, p, c, _ = I.two_opt(p)
        res = {}
        for mode in (tsp.SCAN_FULL, tsp.SCAN_AUTO):
            I.set_scan_mode(mode)
            pm, tm = p.copy(), np.full(n, -1, dtype=np.int32)
            I.reset_stats()
            cm, trace, bi, bp, bc = I.tabu_search(pm, c, tm, tenures)
            res[mode] = (cm, list(trace), bi, bc, pm, tm, I.stats()["pruned_scans"])
    f, a = res[tsp.SCAN_FULL], res[tsp.SCAN_AUTO]
    assert f[:4] == a[:4] and np.array_equal(f[4], a[4]) and np.array_equal(f[5], a[5])
    assert f[6] == 0 and a[6] > 0


def test_tabu_restarts_helper(tsp, oracle):
    """multistart.tabu_restarts (config 5: restarts = start nodes, sharded like the multi-start drivers): the tenure
    schedule equals the reference's tabu_init + tabu_linear_policy, the shard result equals the restarts run by hand."""
    from travellingsalesmanoptimization_b200 import multistart as MS
    xy, _ = load_instance("kroA100")
    n = xy.shape[0]
    tenure, lo, hi = oracle.tabu_init(n)
    inc, ref = True, []
    for _ in range(150):
        tenure, inc = oracle.tabu_linear_step(tenure, lo, hi, inc)
        ref.append(tenure)
    assert MS.tabu_tenures(n, 150) == ref
    with tsp.Instance(xy) as I:
        bc, bs, bp = MS.tabu_restarts(I, 1, 3, 4, 60)             # starts 1, 5, 9: rank 1 of 4
        by_hand = []
        for start in (1, 5, 9):
            code, p, c = I.nn(start)
            code, mv, p, c, _ = I.two_opt(p)
            tl, best = np.full(n, -1, dtype=np.int32), c
            for k, t in enumerate(ref[:60]):
                c, _ = I.tabu_move(p, c, tl, t, k)
                best = min(best, c)
            by_hand.append(best)
    assert bc == min(by_hand) and bs == (1, 5, 9)[by_hand.index(min(by_hand))]
    assert oracle.is_tour(bp) and oracle.tour_cost(xy, bp) == bc


@pytest.mark.parametrize("name", ["d15112", "rand20000"])
def test_full_size_properties(tsp, oracle, name):
    """Size-independent properties at config-5 sizes: successor-array round trip through the device tour slot;
    a random tour descended 80 moves gives the same log under FULL, PRUNED and AUTO (which switches kernels on the
    way); cost bookkeeping equals the oracle's recomputation; a local optimum is a fixed point (0 moves, same tour)."""
    xy = oracle.random_points(20000, 123) if name == "rand20000" else load_instance(name)[0]
    n = xy.shape[0]
    rnd = _rand_tour(99, n)
    logs = {}
    with tsp.Instance(xy) as I:
        I.tour_upload(rnd)
        back, cost0 = I.tour_download()
        assert np.array_equal(back, rnd) and cost0 == oracle.tour_cost(xy, rnd)
        for scan, mode in (("full", tsp.SCAN_FULL), ("pruned", tsp.SCAN_PRUNED), ("auto", tsp.SCAN_AUTO)):
            I.set_scan_mode(mode)
            code, mv, p2, c2, log = I.two_opt(rnd, max_moves=80, log_cap=80)
            logs[scan] = (mv, c2, log, p2)
        assert logs["full"][:3] == logs["pruned"][:3] == logs["auto"][:3]
        assert np.array_equal(logs["full"][3], logs["pruned"][3]) and np.array_equal(logs["full"][3], logs["auto"][3])
        assert logs["full"][1] == oracle.tour_cost(xy, logs["full"][3]) == cost0 + sum(d for _, _, d in logs["full"][2])
        code, p, c = I.nn(7)
        code, mv, popt, copt, _ = I.two_opt(p)
        code, mv2, popt2, copt2, _ = I.two_opt(popt)
        assert mv2 == 0 and copt2 == copt and np.array_equal(popt2, popt)
        d, a, b = oracle.two_opt_scan(xy, popt)
        assert d >= -1e-7


def test_scan_mode_environment(tsp, oracle, monkeypatch):
    """TSPCUDA_SCAN pins the kernel for contexts created afterwards (how the unchanged CLI selects it)."""
    xy = _rand_xy(12, 1500, True)
    path = _rand_tour(4, 1500)
    m_o, p_o, c_o, *_ = oracle.two_opt(xy, path, log=False)
    for val, pruned_expected in (("full", False), ("pruned", True)):
        monkeypatch.setenv("TSPCUDA_SCAN", val)
        with tsp.Instance(xy) as I:
            code, mv, p2, c2, _ = I.two_opt(path)
            st = I.stats()
        assert mv == m_o and c2 == c_o and np.array_equal(p2, p_o)
        assert (st["pruned_scans"] == st["scans"]) == pruned_expected and (st["pruned_scans"] == 0) == (not pruned_expected)


# ---- patching (cplex_model.c:693-755 cx_patching) ---------------------------------------
def _golden_patching():
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "golden_patching.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("key", sorted(_golden_patching().keys()))
def test_patching_golden(tsp, oracle, key):
    """The whole `while (ncomp > 1) cx_patching` loop in one call against tests/golden/golden_patching.json
    (unmodified reference cplex_model.c), and the same solution one cx_patching at a time (max_patches = 1)."""
    from components import cycles_solution, solution_cost
    g = _golden_patching()[key]
    name, seed, ncomp, local = key.rsplit("_", 3)
    xy, _ = load_instance(name)
    path, comp, nc = cycles_solution(int(seed), xy.shape[0], int(ncomp), bool(int(local)), xy)
    with tsp.Instance(xy) as I:
        cost = float(g["start_cost"])
        code, p, c, n1, k1, log = I.patching(path, comp, nc, cost)
        assert code == 0 and n1 == 1 and len(log) == g["patches"] == nc - 1
        assert (k1, fnv1a(p), fnv1a(c)) == (g["cost"], g["path_hash"], g["comp_hash"])
        assert all(a < b and d > 0 for a, b, d in log) and k1 == cost + sum(d for _, _, d in log)
        p2, c2, n2, k2, steps = path, comp, nc, cost, []
        for _ in range(min(nc - 1, 12)):
            code, p2, c2, n2, k2, lg = I.patching(p2, c2, n2, k2, max_patches=1)
            assert code == 0 and len(lg) == 1
            steps += lg
        assert steps == log[:len(steps)] and n2 == nc - len(steps)


@pytest.mark.parametrize("n,integer,ncomp,local", [(6, True, 2, False), (9, True, 3, False), (60, False, 20, False), (257, True, 30, True),
                                                   (700, False, 233, False), (1500, True, 40, True)])
def test_patching_vs_oracle(tsp, oracle, n, integer, ncomp, local):
    """Patch by patch against the CPU restatement: lattice ties, non-integer coordinates, every metric."""
    from components import cycles_solution, solution_cost
    xy = _rand_xy(7 * n + integer, n, integer, lo=0, hi=40 if integer else 3000)
    for wt, twt in ((EUC_REF, tsp.EUC_REF), (CEIL_2D, tsp.CEIL_2D), (ATT, tsp.ATT)):
        path, comp, nc = cycles_solution(n, n, ncomp, local, xy)
        cost = solution_cost(oracle, xy, path, wt)
        with tsp.Instance(xy, twt) as I:
            code, p, c, n1, k1, log = I.patching(path, comp, nc, cost, allow=(tsp.FAILED_PRECONDITION,))
            po, co, no, ko = path.copy(), comp.copy(), nc, cost
            for a, b, d in log:
                ok, no, ko, ao, bo, do = oracle.patching(xy, po, co, no, ko, wt)
                assert ok and (a, b, d) == (ao, bo, do), (wt, a, b, d, ao, bo, do)
            if code == tsp.FAILED_PRECONDITION:                 # coarse lattices can run out of pairs with delta > 0
                assert not oracle.patching(xy, po.copy(), co.copy(), no, ko, wt)[0]
            else:
                assert n1 == 1 and len(log) == nc - 1
            assert n1 == no and k1 == ko and np.array_equal(p, po) and np.array_equal(c, co)


def test_patching_geo(tsp, oracle):
    xy, ewt = load_instance("gr666")
    from components import cycles_solution, solution_cost
    path, comp, nc = cycles_solution(3, xy.shape[0], 25, True, xy)
    cost = solution_cost(oracle, xy, path, GEO)
    with tsp.Instance(xy, tsp.GEO) as I:
        code, p, c, n1, k1, log = I.patching(path, comp, nc, cost)
    for a, b, d in log:
        ok, nc, cost, ao, bo, do = oracle.patching(xy, path, comp, nc, cost, GEO)
        assert ok and (a, b, d) == (ao, bo, do)
    assert n1 == 1 and k1 == cost and np.array_equal(p, path) and np.array_equal(c, comp)


def test_patching_edge_cases(tsp):
    xy = np.array([[0, 0], [10, 0], [0, 0], [10, 0], [5, 5], [6, 6]], dtype=np.float64)
    with tsp.Instance(xy[:4].copy()) as I:
        # coincident 2-cycles: every cross pair has delta <= 0 (the reference indexes path[-1], cplex_model.c:730-735)
        code, p, c, nc, k, log = I.patching([1, 0, 3, 2], [1, 1, 2, 2], 2, 40.0, allow=(tsp.FAILED_PRECONDITION,))
        assert code == tsp.FAILED_PRECONDITION and nc == 2 and k == 40.0 and p.tolist() == [1, 0, 3, 2] and log == []
        # one component: nothing to do
        code, p, c, nc, k, log = I.patching([1, 2, 3, 0], [7, 7, 7, 7], 1, 40.0)
        assert code == 0 and nc == 1 and k == 40.0 and p.tolist() == [1, 2, 3, 0] and log == []
        for bad_path, bad_comp, bad_nc in (([1, 0, 3, 2], [1, 1, 1, 2], 2),      # label not constant on a cycle
                                           ([1, 0, 3, 2], [1, 1, 1, 1], 2),      # two cycles share a label
                                           ([1, 0, 3, 2], [1, 1, 2, 2], 3),      # wrong ncomp
                                           ([1, 1, 3, 2], [1, 1, 2, 2], 2),      # not a permutation
                                           ([0, 2, 3, 1], [1, 2, 2, 2], 2)):     # self-loop
            with pytest.raises(tsp.TspCudaError) as ei:
                I.patching(bad_path, bad_comp, bad_nc, 0.0)
            assert ei.value.code == tsp.INVALID_ARGUMENT


# ---- multi-GPU inside the library (tspcuda_mgpu: worker thread + context per GPU, NCCL incumbent reduction) ------
def _ndev_params(tsp_mod=None):
    return [1, 2, 4, 8]


@pytest.mark.parametrize("ndev", _ndev_params())
def test_mgpu_multistart_matches_single_gpu(tsp, ndev):
    """Sharding the starts over the GPUs of the box must not change the incumbent: same (cost, start, tour) as the
    single-GPU calls, including cost ties (lowest start wins, tsp.c:671) and shards that own no start."""
    if ndev > tsp.device_count():
        pytest.skip("needs %d GPUs" % ndev)
    for name, first, nstarts, stride in (("berlin52", 0, 52, 1), ("berlin52", 3, 5, 7), ("berlin52", 10, 1, 1), ("pr2392", 0, 400, 5)):
        xy, _ = load_instance(name)
        with tsp.Instance(xy) as I, tsp.MultiGpuInstance(xy, ndev=ndev) as M:
            assert M.ndev == ndev
            bs, bp, bc, *_ = I.multistart_two_opt(first, nstarts, stride)
            ms, mp, mc, t_c, t_r = M.multistart(True, first, nstarts, stride)
            assert (ms, mc) == (bs, bc) and np.array_equal(mp, bp), (name, first, nstarts, stride)
            bs, bp, bc, _ = I.nn_multi(first, nstarts, stride)
            ms, mp, mc, t_c, t_r = M.multistart(False, first, nstarts, stride)
            assert (ms, mc) == (bs, bc) and np.array_equal(mp, bp), (name, first, nstarts, stride)
    # lattice instance: many starts reach the same cost, the lowest start must win on every GPU count
    xy = _rand_xy(11, 64, True, lo=0, hi=6)
    with tsp.Instance(xy) as I, tsp.MultiGpuInstance(xy, ndev=ndev) as M:
        bs, bp, bc, *_ = I.multistart_two_opt(0, 64, 1)
        ms, mp, mc, *_ = M.multistart(True, 0, 64, 1)
        assert (ms, mc) == (bs, bc) and np.array_equal(mp, bp)


def test_mgpu_more_gpus_than_visible_fails_loudly(tsp):
    xy, _ = load_instance("berlin52")
    with pytest.raises(tsp.TspCudaError) as ei:
        tsp.MultiGpuInstance(xy, ndev=tsp.device_count() + 1)
    assert ei.value.code == tsp.UNAVAILABLE
