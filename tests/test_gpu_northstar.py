"""GPU parity tests (-m gpu) at the north_star's OWN sizes -- the configs of BASELINE.json that the small-instance
tests of test_gpu_parity.py only reach through timing runs:

  config 3 / 4   usa13509 (non-integer coordinates, EucGeneral) and pla85900 through the MULTI-START kernels
                 (K4b batched NN + K8 multi_pruned_kernel, one CTA or a cluster of 2/4/8 CTAs per descent), start by
                 start against nn + two_opt on the whole GPU (K7), the reference-run golden of start 0 and the CPU
                 oracle's move audit (refinment.c:39-93, heuristics.c:74-116)
  config 5       d15112 and the reference's `-n 20000 -seed 123` instance: full NN(0) descents against
                 tests/golden/golden_large.json (oracle pinned on the reference objects, tests/golden/make_golden_large.py),
                 64 tabu iterations and 6 VNS iterations from that local optimum against the golden traces AND against
                 the oracle run live (metaheuristic.c:188-245, :277-321)
Bar: bit-exact successor arrays, identical move counts / cost traces."""
import ctypes
import json
import os

import numpy as np
import pytest

from oracle.bindings import fnv1a, load_instance

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def large():
    with open(os.path.join(HERE, "golden", "golden_large.json")) as f:
        g = json.load(f)
    tours = np.load(os.path.join(HERE, "golden", "large_tours.npz"))
    return g, {k: np.ascontiguousarray(tours[k], dtype=np.int32) for k in tours.files}


def _xy(oracle, name):
    return oracle.random_points(20000, 123) if name == "rand20000" else load_instance(name)[0]


def _k7_reference(tsp, xy, starts, log_for=None):
    """nn + two_opt per start on the whole GPU (library default: K4b/K4 + K7): {start: (nn_cost, moves, cost, path, log)}."""
    out = {}
    with tsp.Instance(xy) as I:
        for s in starts:
            code, p, c = I.nn(s)
            assert code == 0
            cap = 16384 if s == log_for else 0
            code, mv, p2, c2, log = I.two_opt(p, log_cap=cap)
            assert code == 0
            out[s] = (c, mv, c2, p2, log, p)
    return out


def _multistart_modes(tsp, monkeypatch, xy, starts_spec, ref, modes):
    """tspcuda_multistart_two_opt under every TSPCUDA_K8_CLUSTER mode: per-start (nn cost, moves, final cost) and the best
    tour against the K7 results; then every start alone (so that each start's TOUR is compared, not only the best one)."""
    first, count, stride = starts_spec
    starts = [first + k * stride for k in range(count)]
    best = min(starts, key=lambda s: (ref[s][2], s))
    for cl in modes:
        monkeypatch.setenv("TSPCUDA_K8_CLUSTER", str(cl))
        with tsp.Instance(xy) as I:
            bs, bp, bc, nn_costs, final, moves = I.multistart_two_opt(first, count, stride)
            st = I.stats()
        assert st["pruned_scans"] == int(moves.sum()) + count, cl          # K8 answered every scan
        assert [(a, int(m), f) for a, m, f in zip(nn_costs, moves, final)] == [ref[s][:3] for s in starts], cl
        assert (bs, bc) == (best, ref[best][2]) and np.array_equal(bp, ref[best][3]), cl
    monkeypatch.setenv("TSPCUDA_K8_CLUSTER", str(modes[-1]))
    with tsp.Instance(xy) as I:
        for s in starts:
            bs, bp, bc, nn_costs, final, moves = I.multistart_two_opt(s, 1, 1)
            assert (bs, nn_costs[0], int(moves[0]), bc) == (s, ref[s][0], ref[s][1], ref[s][2]), s
            assert np.array_equal(bp, ref[s][3]), s


def test_pla85900_multistart_kernels(tsp, oracle, monkeypatch):
    """Config 4's own kernels at config-4 size: starts 0, 24, 48 of pla85900 through K4b + K8 under cluster sizes 1/2/4/8.
    Start 0 = 5589 moves -> 148543747 (the full-scan descent of bench.py / profiles/r01_bench_first.json); start 24 is
    audited by the oracle: every logged delta re-derived, full CPU scans at sampled iterations, final local optimality."""
    xy, _ = load_instance("pla85900")
    ref = _k7_reference(tsp, xy, (0, 24, 48), log_for=24)
    assert (ref[0][1], int(ref[0][2])) == (5589, 148543747)
    c, mv, c2, p2, log, p = ref[24]
    assert len(log) == mv
    la = [x[0] for x in log]; lb = [x[1] for x in log]; ld = [float(x[2]) for x in log]
    r, pend = oracle.audit_moves(xy, p, la, lb, ld, [0, mv // 2, mv - 1])
    assert r == 0 and np.array_equal(pend, p2)
    d, a, b = oracle.two_opt_scan(xy, p2)
    assert d >= -1e-7
    _multistart_modes(tsp, monkeypatch, xy, (0, 3, 24), ref, (1, 2, 4, 8))


def test_usa13509_multistart_kernels(tsp, oracle, golden, monkeypatch):
    """Config 3's instance (13 337 of 13 509 nodes carry decimals: EucGeneral, where an FMA contraction would change
    costs) through K4b + K8 and the clusters.  Start 0 = the reference objects' own run: 1983 moves -> 21119812."""
    xy, _ = load_instance("usa13509")
    assert np.any(xy != np.floor(xy))
    ref = _k7_reference(tsp, xy, (0, 7, 14, 21, 28))
    s = golden["usa13509"]["nn0_2opt_survey"]
    assert (ref[0][1], int(ref[0][2])) == (s["moves"], s["cost"]) and int(ref[0][0]) == golden["usa13509"]["nn0"]["cost"]
    for st_ in (7, 28):                                       # two more starts against the oracle's own NN tour + cost recomputation
        e, po, co = oracle.nn(xy, st_)
        assert co == ref[st_][0] and np.array_equal(po, ref[st_][5])
        assert oracle.is_tour(ref[st_][3]) and oracle.tour_cost(xy, ref[st_][3]) == ref[st_][2]
    _multistart_modes(tsp, monkeypatch, xy, (0, 5, 7), ref, (1, 2, 4, 8))
    monkeypatch.delenv("TSPCUDA_K8_CLUSTER")
    with tsp.Instance(xy) as I:                               # library default for a batch of this size (>= 12 starts: clusters)
        bs, bp, bc, nn_costs, final, moves = I.multistart_two_opt(0, 15, 2)
    for s_ in (0, 14, 28):
        k = s_ // 2
        assert (nn_costs[k], int(moves[k]), final[k]) == ref[s_][:3]


@pytest.mark.parametrize("name", ["d15112", "rand20000"])
@pytest.mark.parametrize("scan", ["full", "pruned", "auto"])
def test_config5_full_descent_golden(tsp, oracle, large, name, scan):
    """NN(0) -> 2-opt local optimum at the config-5 sizes against the pinned oracle's run (d15112: 2060 moves -> 1659198;
    reference objects: NN(0) and the first 20 moves of the same descent, tests/golden/golden.json)."""
    g, tours = large
    xy = _xy(oracle, name)
    with tsp.Instance(xy) as I:
        I.set_scan_mode({"auto": tsp.SCAN_AUTO, "full": tsp.SCAN_FULL, "pruned": tsp.SCAN_PRUNED}[scan])
        code, p, c = I.nn(0)
        assert (int(c), fnv1a(p)) == (g[name]["nn0"]["cost"], g[name]["nn0"]["hash"])
        code, moves, p2, c2, log = I.two_opt(p, log_cap=8)
    e = g[name]["nn0_2opt"]
    assert (moves, int(c2), fnv1a(p2)) == (e["moves"], e["cost"], e["hash"])
    assert np.array_equal(p2, tours[name])


@pytest.mark.parametrize("name", ["d15112", "rand20000"])
def test_config5_tabu_vs_oracle(tsp, oracle, large, name):
    """64 iterations of the tabu loop (linear tenure, metaheuristic.c:40-59) from the local optimum, at config-5 size: the
    device-resident loop (K5c pruned scan + K5b fallback) and the per-call path against the golden trace and against the
    oracle's tabu_best_move run live (tour and tabu list compared after every 16th iteration and at the end)."""
    g, tours = large
    xy = _xy(oracle, name)
    n = xy.shape[0]
    p0, c0 = tours[name], float(g[name]["nn0_2opt"]["cost"])
    K = g[name]["tabu"]["iters"]
    tenure, lo, hi = oracle.tabu_init(n)
    inc, tenures = True, []
    for _ in range(K):
        tenure, inc = oracle.tabu_linear_step(tenure, lo, hi, inc)
        tenures.append(tenure)
    po, co, to, trace_o, snaps = p0.copy(), c0, np.full(n, -1, dtype=np.int32), [], {}
    for k, t in enumerate(tenures):
        co, a, b, d = oracle.tabu_move(xy, po, co, to, t, k)
        trace_o.append(int(co))
        if k % 16 == 15:
            snaps[k] = (po.copy(), to.copy())
    assert trace_o == g[name]["tabu"]["trace"] and fnv1a(po) == g[name]["tabu"]["hash"] and fnv1a(to) == g[name]["tabu"]["tabu_list_hash"]
    for mode in (tsp.SCAN_AUTO, tsp.SCAN_FULL):
        with tsp.Instance(xy) as I:
            I.set_scan_mode(mode)
            pg, tg, cg, trace_g = p0.copy(), np.full(n, -1, dtype=np.int32), c0, []
            for k0 in range(0, K, 16):                      # four device-resident runs of 16 iterations (first_iter continues)
                cg, tr, bi, bp, bc = I.tabu_search(pg, cg, tg, tenures[k0:k0 + 16], first_iter=k0, incumbent_cost=c0)
                trace_g += [int(x) for x in tr]
                assert np.array_equal(pg, snaps[k0 + 15][0]) and np.array_equal(tg, snaps[k0 + 15][1]), (mode, k0)
            st = I.stats()
        assert trace_g == trace_o
        assert (st["pruned_scans"] > 0) == (mode == tsp.SCAN_AUTO)
    with tsp.Instance(xy) as I:                             # per-call path: tabu_best_move, host tour in and out
        pg, tg, cg = p0.copy(), np.full(n, -1, dtype=np.int32), c0
        for k, t in enumerate(tenures[:8]):
            cg, mv = I.tabu_move(pg, cg, tg, t, k)
            assert int(cg) == trace_o[k]


@pytest.mark.parametrize("name", ["d15112", "rand20000"])
def test_config5_vns_vs_oracle(tsp, oracle, large, name):
    """6 iterations of the loop of mh_VNS (metaheuristic.c:277-321: descent, incumbent test, rand() % 9 - 2 kicks) from the
    local optimum at config-5 size, glibc stream seeded: tspcuda_vns_search against the golden trace and the oracle's loop."""
    g, tours = large
    xy = _xy(oracle, name)
    libc = ctypes.CDLL(None)
    v = g[name]["vns"]
    p0, c0 = tours[name], float(g[name]["nn0_2opt"]["cost"])
    libc.srand(v["seed"])
    po, trace_o, moves_o = p0.copy(), [], []
    for i in range(v["iters"]):
        m, po, co, *_ = oracle.two_opt(xy, po, log=False)
        trace_o.append(int(co)); moves_o.append(int(m))
        for _ in range(libc.rand() % 9 - 2):
            oracle.vns_kick(po)
    assert (trace_o, moves_o, fnv1a(po)) == (v["trace"], v["moves"], v["hash"])
    libc.srand(v["seed"])
    with tsp.Instance(xy) as I:
        code, pg, cg, tr, bi, bp, bc, done = I.vns_search(p0, v["iters"], c0)
        st = I.stats()
    assert code == 0 and done == v["iters"] and [int(x) for x in tr] == trace_o and np.array_equal(pg, po)
    assert st["moves"] == sum(moves_o)
    assert (bi, int(bc)) == (v["best_iter"], v["best_cost"])


def test_multistart_time_limit(tsp, oracle):
    """time_limit_s inside the batched multi-start (heuristics.c:83-90, refinment.c:17-24): a limit that cannot be met returns
    DEADLINE_EXCEEDED with the best over the PREFIX of starts handed out -- complete descents equal their unlimited runs,
    cut ones hold valid, partially descended tours; a generous limit changes nothing; a limit of 0 takes no start."""
    import time
    xy, _ = load_instance("usa13509")
    S = 1184
    with tsp.Instance(xy) as I:
        I.multistart_two_opt(0, 8)                                                         # warm (module load, scratch growth)
        t0 = time.perf_counter()
        bs0, bp0, bc0, nn0, fin0, mv0 = I.multistart_two_opt(0, S)                        # unlimited reference run
        wall0 = time.perf_counter() - t0
        code, done, bs, bp, bc, nn, fin, mv = I.multistart_two_opt(0, S, time_limit=600.0)
        assert code == 0 and done == S and (bs, bc) == (bs0, bc0) and np.array_equal(bp, bp0) and np.array_equal(fin, fin0)
        code, done, bs, bp, bc, nn, fin, mv = I.multistart_two_opt(0, S, time_limit=0.5 * wall0)
        assert code == tsp.DEADLINE_EXCEEDED and 0 < done <= S, (code, done, wall0)
        assert np.all(fin[done:] == -1) and np.all(mv[done:] == -1) and np.all(fin[:done] > 0)
        assert np.all(mv[:done] <= mv0[:done]) and np.all(fin[:done] >= fin0[:done])       # a cut descent stopped early
        assert np.any(mv[:done] < mv0[:done]) or done < S
        complete = mv[:done] == mv0[:done]
        assert np.array_equal(fin[:done][complete], fin0[:done][complete])
        k = int(np.argmin(fin[:done]))                                                     # first minimum = lowest start
        assert (bs, bc) == (k, fin[k]) and oracle.is_tour(bp) and oracle.tour_cost(xy, bp) == bc
        code, done, bs, bp, bc, nn, fin, mv = I.multistart_two_opt(0, S, time_limit=0.0)
        assert code == tsp.DEADLINE_EXCEEDED and done == 0 and bs == -1
        # all-start NN under a limit: a prefix of whole batches
        bsn0, bpn0, bcn0, costs0 = I.nn_multi(0, 2000)
        code, done, bsn, bpn, bcn, costs = I.nn_multi(0, 2000, time_limit=600.0)
        assert code == 0 and done == 2000 and (bsn, bcn) == (bsn0, bcn0) and np.array_equal(costs, costs0)
        code, done, bsn, bpn, bcn, costs = I.nn_multi(0, 2000, time_limit=0.0)
        assert code == tsp.DEADLINE_EXCEEDED and done == 0 and bsn == -1
    # the full-scan multi-start kernel (K6, tours in shared memory) honours the same flag
    xy = np.ascontiguousarray(np.round(np.random.default_rng(5).uniform(0, 20000, size=(3000, 2))))
    with tsp.Instance(xy) as I:
        I.set_scan_mode(tsp.SCAN_FULL)
        I.multistart_two_opt(0, 8)
        t0 = time.perf_counter()
        bs0, bp0, bc0, nn0, fin0, mv0 = I.multistart_two_opt(0, 600)
        wall0 = time.perf_counter() - t0
        code, done, bs, bp, bc, nn, fin, mv = I.multistart_two_opt(0, 600, time_limit=0.5 * wall0)
        assert code == tsp.DEADLINE_EXCEEDED and 0 < done <= 600 and np.all(fin[done:] == -1), (code, done, wall0)
        complete = mv[:done] == mv0[:done]
        assert np.array_equal(fin[:done][complete], fin0[:done][complete]) and oracle.is_tour(bp) and oracle.tour_cost(xy, bp) == bc


def test_restarts_c_abi(tsp, oracle):
    """tspcuda_restarts (config 5): the C entry point against the same restarts made by hand from the public calls --
    tabu: NN -> 2-opt -> tabu loop with the reference's tenure schedule; VNS: srand(seed + start), NN -> mh_VNS loop --
    and shard independence (the merged best of two shards = the best of one call over every start)."""
    from travellingsalesmanoptimization_b200 import multistart as MS
    xy, _ = load_instance("rl1304")
    with tsp.Instance(xy) as I:
        code, done, bs, bp, bc = I.restarts(tsp.RESTART_TABU, 3, 4, 5, iters=120)
        hc, hs, hp = MS.tabu_restarts(I, 3, 4, 5, 120)
        assert code == 0 and done == 4 and (bs, bc) == (hs, hc) and np.array_equal(bp, hp)
        assert oracle.is_tour(bp) and oracle.tour_cost(xy, bp) == bc
        code, done, bs, bp, bc = I.restarts(tsp.RESTART_VNS, 0, 6, 1, iters=10, seed=5)
        hc, hs, hp = MS.vns_restarts(I, 0, 6, 1, 10, seed=5)
        assert code == 0 and done == 6 and (bs, bc) == (hs, hc) and np.array_equal(bp, hp)
        parts = [I.restarts(tsp.RESTART_VNS, *MS.shard_starts(0, 6, r, 2), iters=10, seed=5) for r in range(2)]
        merged = min(parts, key=lambda p: (p[4], p[2]))
        assert (merged[2], merged[4]) == (bs, bc) and np.array_equal(merged[3], bp)
        code, done, bs2, bp2, bc2 = I.restarts(tsp.RESTART_TABU, 0, 50, 1, iters=2000, time_limit=0.2)
        assert code == tsp.DEADLINE_EXCEEDED and 0 < done < 50 and oracle.is_tour(bp2) and oracle.tour_cost(xy, bp2) == bc2


@pytest.mark.parametrize("per_node", ["0", "2"])
def test_delta_cache_under_heap_pressure(tsp, monkeypatch, per_node):
    """The rare paths of the delta cache (csrc/pruned.cuh): with TSPCUDA_NB_HEAP the record heap holds 4096 (+ 2 n)
    records instead of 24 n, so long lists find no room (they stay stale and are searched again in every scan, giant
    search boxes are walked by one warp without being cached) and the heap restarts every few moves.  Slower, and
    move for move the reference's descent: usa13509 NN(0) -> 1983 moves -> 21119812 (golden), a kicked tour, and the
    multi-start kernel start by start against the default configuration."""
    xy, _ = load_instance("usa13509")
    with tsp.Instance(xy) as I:                                   # default heap
        I.set_scan_mode(tsp.SCAN_PRUNED)
        code, p, c = I.nn(0)
        code, m0, p0, c0, log0 = I.two_opt(p, log_cap=4096)
        bs0, bp0, bc0, nn0, fin0, mv0 = I.multistart_two_opt(0, 40)
        # a tour with three instance-spanning edges (what a VNS kick makes): giant boxes in the first scans
        order = np.empty(len(p0), dtype=np.int64)
        cur = 0
        for k in range(len(p0)):
            order[k] = cur
            cur = p0[cur]
        kicked = np.concatenate([order[:3000], order[9000:12000], order[3000:9000], order[12000:]])
        pk = np.empty_like(p0)
        pk[kicked] = np.roll(kicked, -1)
        code, mk0, pk0, ck0, logk0 = I.two_opt(pk, log_cap=4096)
    assert (m0, c0) == (1983, 21119812.0)
    monkeypatch.setenv("TSPCUDA_NB_HEAP", per_node)
    with tsp.Instance(xy) as I:
        I.set_scan_mode(tsp.SCAN_PRUNED)
        code, m1, p1, c1, log1 = I.two_opt(p, log_cap=4096)
        assert (m1, c1) == (m0, c0) and log1 == log0 and np.array_equal(p1, p0)
        code, mk1, pk1, ck1, logk1 = I.two_opt(pk, log_cap=4096)
        assert (mk1, ck1) == (mk0, ck0) and logk1 == logk0 and np.array_equal(pk1, pk0)
        bs1, bp1, bc1, nn1, fin1, mv1 = I.multistart_two_opt(0, 40)
        assert (bs1, bc1) == (bs0, bc0) and np.array_equal(fin1, fin0) and np.array_equal(mv1, mv0) and np.array_equal(bp1, bp0)


@pytest.mark.parametrize("knob,value", [("TSPCUDA_NB_GIANT_AGE", "0"), ("TSPCUDA_NB_GIANT_AGE", "2"), ("TSPCUDA_NB_GIANT_AGE", "1000"),
                                         ("TSPCUDA_NB_RESTART_PCT", "25"), ("TSPCUDA_NB_RESTART_PCT", "400")])
def test_delta_cache_policies_do_not_change_results(tsp, monkeypatch, knob, value):
    """Two policies of the delta cache (csrc/pruned.cuh) decide only WHEN work is done, never what the scan returns:
    giant_age (a search box of more than 512 nodes is evaluated k times before it is cached: the VNS loop, whose kicks make
    instance-spanning edges) and the heap restart rule (how many abandoned records are streamed before the heap starts
    over; 25 % restarts every few moves, 400 % almost never -- with giant runs kept in place across searches).  d15112: the
    mh_VNS loop (metaheuristic.c:277-321) with a fixed rand() stream, the NN(0) descent and a multi-start batch, against the
    library defaults: same traces, tours, move logs and incumbents."""
    libc = ctypes.CDLL(None)
    xy, _ = load_instance("d15112")

    def run():
        with tsp.Instance(xy) as I:
            I.set_scan_mode(tsp.SCAN_PRUNED)
            code, p, c = I.nn(0)
            code, m, p1, c1, log = I.two_opt(p, log_cap=4096)
            libc.srand(11)
            code, pv, cv, tr, bi, bp, bc, done = I.vns_search(p1, 12, c1)
            ms = I.multistart_two_opt(0, 12)
        return (m, c1, log, fnv1a(p1), cv, tr.tolist(), bi, bc, fnv1a(pv), ms[0], ms[2], ms[4].tolist(), ms[5].tolist(), fnv1a(ms[1]))

    ref = run()
    monkeypatch.setenv(knob, value)
    assert run() == ref
