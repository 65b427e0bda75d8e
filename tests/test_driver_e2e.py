"""The drop-in boundary end to end: the reference's UNCHANGED driver (src/main.c compiled by
oracle/Makefile) linked with libtspcuda_interpose.so + libtspcuda.so ahead of libtspref.so
(build/bin/tsp), run with the reference's own CLI flags, compared with the golden values and with
the reference CPU CLI (oracle/_ref/tsp_ref) on the same arguments.

The binaries are built where /root/reference exists (__graft_entry__.build()) and travel to the GPU
box as built artefacts; where they are absent these tests skip."""
import os
import subprocess

import numpy as np
import pytest

from oracle.bindings import load_instance

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TSP_GPU = os.path.join(ROOT, "build", "bin", "tsp")
TSP_REF = os.path.join(ROOT, "oracle", "_ref", "tsp_ref")


def write_tsplib(path, name, xy, ewt="EUC_2D"):
    with open(path, "w") as f:
        f.write("NAME : %s\nTYPE : TSP\nDIMENSION : %d\nEDGE_WEIGHT_TYPE : %s\nNODE_COORD_SECTION\n" % (name, xy.shape[0], ewt))
        for i, (x, y) in enumerate(xy):
            f.write("%d %s %s\n" % (i + 1, repr(float(x)), repr(float(y))))   # repr round-trips through atof
        f.write("EOF\n")


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    d = tmp_path_factory.mktemp("driver")
    os.makedirs(d / "results")          # metaheuristic.c:97,275 write results/*.dat relative to cwd
    for name in ["berlin52", "kroA100", "rl1304", "pr2392", "usa13509"]:
        xy, _ = load_instance(name)
        write_tsplib(d / (name + ".tsp"), name, xy)
    for name in ["att48", "att532", "dsj1000", "burma14"]:       # ATT, ATT, CEIL_2D, GEO headers
        xy, ewt = load_instance(name)
        write_tsplib(d / (name + ".tsp"), name, xy, ewt)
    return d


def run_cli(exe, workdir, args, env_extra=None, check=True):
    if not os.path.exists(exe):
        pytest.skip("%s not built (needs /root/reference at build time)" % os.path.relpath(exe, ROOT))
    env = dict(os.environ)
    env["PATH"] = os.path.join(ROOT, "oracle", "stubs") + ":" + env["PATH"]      # no-op gnuplot
    env["TSPCUDA_STATS"] = "1"
    env.update(env_extra or {})
    r = subprocess.run([exe] + args, cwd=workdir, env=env, capture_output=True, text=True, timeout=600)
    if check:
        assert r.returncode == 0, (r.returncode, r.stdout[-500:], r.stderr[-2000:])
    return r


def calls(stderr):
    line = [ln for ln in stderr.splitlines() if ln.startswith("[tspcuda] interposed calls:")][-1]
    return {k: int(v) for k, v in (kv.split("=") for kv in line.split(":", 1)[1].split())}


def test_cpu_box_fails_loudly(workdir):
    """No GPU => the drop-in CLI must NOT produce a result (no CPU fallback), yet must show that the
    unchanged driver reached the interposed symbols."""
    import travellingsalesmanoptimization_b200 as T
    if T.device_count() > 0:
        pytest.skip("a CUDA device is present")
    r = run_cli(TSP_GPU, workdir, ["-f", "berlin52.tsp", "-alg", "GREEDY", "-q"], check=False)
    assert r.returncode != 0 and "Cost:" not in r.stdout
    c = calls(r.stderr)
    assert c["tsp_compute_costs"] == 1


@pytest.mark.gpu
@pytest.mark.parametrize("alg,extra,cost", [
    ("GREEDY", [], "8980.00"), ("GREEDY_ITER", [], "8181.00"), ("2OPT_GREEDY", [], "7657.00"),
    ("TABU_SEARCH", ["-k", "200"], "7542.00"), ("VNS", ["-k", "50"], "7670.00"), ("EXTRA_MILEAGE", [], "9021.00")])
def test_berlin52_cli_golden(workdir, alg, extra, cost):
    """Golden values of BASELINE.md section 3 through the unchanged CLI: make/bin/tsp -f F -q -alg A."""
    r = run_cli(TSP_GPU, workdir, ["-f", "berlin52.tsp", "-q", "-alg", alg] + extra)
    assert r.stdout.strip() == "Cost: " + cost          # stdout contract of errors.c:159-164
    c = calls(r.stderr)
    assert c["tsp_compute_costs"] == 1
    if alg == "GREEDY":
        assert c["h_greedyutil"] == 1
    if alg == "TABU_SEARCH":
        assert c["mh_TabuSearch"] == 1 and c["tabu_best_move"] == 0 and c["h_greedy_2opt"] == 1
        with open(os.path.join(workdir, "results", "TabuResults.dat")) as f:
            lines = f.read().split()
        assert len(lines) == 200 and lines[0].startswith("0,") and min(float(x.split(",")[1]) for x in lines) == 7542.0
        # the reference's own loop over the interposed tabu_best_move: same result, same results file
        r0 = run_cli(TSP_GPU, workdir, ["-f", "berlin52.tsp", "-q", "-alg", alg] + extra, {"TSPCUDA_BATCH": "0"})
        c0 = calls(r0.stderr)
        assert r0.stdout == r.stdout and c0["tabu_best_move"] == 200 and c0["mh_TabuSearch"] == 0
        with open(os.path.join(workdir, "results", "TabuResults.dat")) as f:
            assert f.read().split() == lines
    if alg == "VNS":
        assert c["mh_VNS"] == 1 and c["ref_2opt"] == 0 and c["h_Greedy_iterative"] == 1
        with open(os.path.join(workdir, "results", "VNSResults.dat")) as f:
            lines = f.read().split()
        assert len(lines) == 50
        # the reference's own loop (its vns_kick, the interposed ref_2opt once per iteration): same result, same file
        r0 = run_cli(TSP_GPU, workdir, ["-f", "berlin52.tsp", "-q", "-alg", alg] + extra, {"TSPCUDA_BATCH": "0"})
        c0 = calls(r0.stderr)
        assert r0.stdout == r.stdout and c0["ref_2opt"] == 50 and c0["mh_VNS"] == 0
        with open(os.path.join(workdir, "results", "VNSResults.dat")) as f:
            assert f.read().split() == lines
    if alg == "EXTRA_MILEAGE":
        assert c["h_ExtraMileage"] == 1
        # the reference's own h_ExtraMileage (EM_MAX loop over the host matrix) over the interposed util
        r0 = run_cli(TSP_GPU, workdir, ["-f", "berlin52.tsp", "-q", "-alg", alg], {"TSPCUDA_BATCH": "0"})
        c0 = calls(r0.stderr)
        assert r0.stdout == r.stdout and c0["h_extramileage_util"] == 1 and c0["h_ExtraMileage"] == 0


def test_mislink_is_fatal(workdir):
    """a mis-link must not masquerade as a pass.  With the reference objects AHEAD of the interposer in the
    lookup order, ref_2opt_once resolves into libtspref.so -- every "GPU" result would be the reference's CPU code -- and
    the interposer's load-time self-check ends the process; TSPCUDA_ALLOW_MISLINK=1 downgrades that to a warning."""
    import sys
    ref = os.path.join(ROOT, "build", "ref", "libtspref.so")
    itp = os.path.join(ROOT, "travellingsalesmanoptimization_b200", "libtspcuda_interpose.so")
    if not (os.path.exists(ref) and os.path.exists(itp)):
        pytest.skip("drop-in artefacts not built (needs /root/reference at build time)")
    code = ("import ctypes; ctypes.CDLL(%r, mode=ctypes.RTLD_GLOBAL); ctypes.CDLL(%r, mode=ctypes.RTLD_GLOBAL); print('survived')" % (ref, itp))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 70 and "survived" not in r.stdout and "FATAL" in r.stderr and "NOT active" in r.stderr
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120, env=dict(os.environ, TSPCUDA_ALLOW_MISLINK="1"))
    assert r.returncode == 0 and "survived" in r.stdout and "WARNING" in r.stderr
    # the right order (what build/bin/tsp links): the check is silent (with or without a GPU)
    r = run_cli(TSP_GPU, workdir, ["-f", "berlin52.tsp", "-alg", "GREEDY", "-q"], check=False)
    assert "NOT active" not in r.stderr and r.returncode != 70


def test_cli_does_not_link_the_oracle_directory():
    """The product binary is the unchanged reference driver + the interposer + libtspcuda; nothing of it may come from
    oracle/ (test infrastructure)."""
    if not os.path.exists(TSP_GPU):
        pytest.skip("build/bin/tsp not built (needs /root/reference at build time)")
    out = subprocess.run(["readelf", "-d", TSP_GPU], capture_output=True, text=True).stdout
    assert "oracle" not in out and "libtspcuda_interpose.so" in out and "libtspref.so" in out


@pytest.mark.gpu
def test_time_limit_keeps_the_batched_paths(workdir, golden):
    """The reference's own tooling always passes -t 600 (scripts/compare_algs.py:18,67).  The limit is honoured INSIDE the
    batched calls, so the whole-loop overrides stay active: config 2 (pr2392, all 2392 starts) under -t 600 gives the
    golden 393339 with h_greedy_2opt = 1 and no per-start helper call."""
    r = run_cli(TSP_GPU, workdir, ["-f", "pr2392.tsp", "-alg", "2OPT_GREEDY", "-t", "600", "-q"])
    assert r.stdout.strip() == "Cost: %d.00" % golden["pr2392"]["greedy_2opt_survey"]["cost"]
    c = calls(r.stderr)
    assert c["h_greedy_2opt"] == 1 and c["h_greedyutil"] == 0 and c["ref_2opt"] == 0
    r = run_cli(TSP_GPU, workdir, ["-f", "pr2392.tsp", "-alg", "GREEDY_ITER", "-t", "600", "-q"])
    assert r.stdout.strip() == "Cost: %d.00" % golden["pr2392"]["allnn"]["cost"] and calls(r.stderr)["h_Greedy_iterative"] == 1


@pytest.mark.gpu
def test_time_limit_cuts_a_batch_short(workdir, golden):
    """-t 1 on usa13509's all-start NN + 2-opt (13 509 descents; the driver's clock starts after the instance is read,
    main.c:177): the call comes back at the limit with a VALID tour -- the best over the prefix of starts reached, descents in
    flight cut where they were (heuristics.c:83-90, refinment.c:17-24,29-34) -- and the driver prints its cost and exits 0,
    as the reference does after DEADLINE_EXCEEDED (main.c:180).  (The exact semantics are tested through the C ABI in
    test_gpu_northstar.py::test_multistart_time_limit.)"""
    import time
    t0 = time.perf_counter()
    r = run_cli(TSP_GPU, workdir, ["-f", "usa13509.tsp", "-alg", "2OPT_GREEDY", "-t", "1", "-q"], {"TSPCUDA_MATRIX": "off"})
    wall = time.perf_counter() - t0
    assert calls(r.stderr)["h_greedy_2opt"] == 1
    cost = float(r.stdout.strip().split()[1])
    assert 19.9e6 < cost < golden["usa13509"]["nn0"]["cost"]           # 2-opt tours: below NN(0), above the optimum 19 982 859
    assert wall < 40.0                                                  # the whole job would take far longer than the limit


@pytest.mark.gpu
def test_reference_loops_over_interposed_helpers(workdir, golden):
    """TSPCUDA_BATCH=0: the reference's own h_greedy_2opt loop runs and reaches the device through
    h_greedyutil / ref_2opt start by start -- same result as the batched override."""
    r = run_cli(TSP_GPU, workdir, ["-f", "kroA100.tsp", "-q", "-alg", "2OPT_GREEDY"], {"TSPCUDA_BATCH": "0"})
    assert r.stdout.strip() == "Cost: %d.00" % golden["kroA100"]["greedy_2opt"]["cost"]
    c = calls(r.stderr)
    assert c["h_greedyutil"] == 100 and c["ref_2opt"] == 100 and c["h_greedy_2opt"] == 0
    r2 = run_cli(TSP_GPU, workdir, ["-f", "kroA100.tsp", "-q", "-alg", "2OPT_GREEDY"])
    assert r2.stdout == r.stdout and calls(r2.stderr)["h_greedy_2opt"] == 1


@pytest.mark.gpu
def test_matches_reference_cli(workdir, golden):
    """Same arguments through the reference CPU build and the GPU drop-in: identical stdout."""
    args = ["-f", "rl1304.tsp", "-q", "-alg", "GREEDY_ITER"]
    g = run_cli(TSP_GPU, workdir, args)
    r = run_cli(TSP_REF, workdir, args)
    assert g.stdout == r.stdout == "Cost: %d.00\n" % golden["rl1304"]["allnn"]["cost"]


@pytest.mark.gpu
@pytest.mark.parametrize("n,seed,alg,algid,k", [(300, 7, "2OPT_GREEDY", 2, None), (200, 11, "TABU_SEARCH", 3, 40), (150, 5, "GREEDY_ITER", 1, None),
                                                 (250, 3, "VNS", 4, 12), (180, 9, "EXTRA_MILEAGE", 5, None), (90, 2, "GREEDY", 0, None)])
def test_random_instances_match_reference_objects(workdir, reference, n, seed, alg, algid, k):
    """`-n N -seed S` (src/tsp.c:468-481).  The reference's own CLI aborts at exit for -n on this glibc
    (free(): invalid pointer, the utils_format_title overflow on its 40-byte plot-name buffer), so the
    GPU drop-in CLI is compared with the same reference objects driven through oracle/ref_api.c."""
    cwd = os.getcwd()
    os.chdir(workdir)
    try:
        reference.random_points(n, seed)
        e, p, c, s = reference.run(algid, k if k is not None else 2**31 - 1)
    finally:
        os.chdir(cwd)
    g = run_cli(TSP_GPU, workdir, ["-n", str(n), "-seed", str(seed), "-q", "-alg", alg] + (["-k", str(k)] if k else []))
    assert g.stdout.strip() == "Cost: %.2f" % c


@pytest.mark.gpu
def test_usa13509_cli(workdir, golden):
    """Config-3 instance through the CLI: 1.46 GB host matrix filled from the device, NN on the device."""
    r = run_cli(TSP_GPU, workdir, ["-f", "usa13509.tsp", "-q", "-alg", "GREEDY"])
    assert r.stdout.strip() == "Cost: %d.00" % golden["usa13509"]["nn0"]["cost"]


@pytest.mark.gpu
def test_reader_extension_att_ceil2d_geo(workdir, oracle):
    """the interposed tsp_read_input accepts CEIL_2D and ATT files (the reference
    stops at tsp.c:576-583).  No reference behaviour exists: expected values come from the TSPLIB95
    restatement in the oracle ("parity unpinned")."""
    from oracle.bindings import ATT, CEIL_2D, EUC_REF
    for name, wt, alg in (("att48", ATT, "2OPT_GREEDY"), ("att532", ATT, "GREEDY_ITER"), ("dsj1000", CEIL_2D, "GREEDY")):
        xy, ewt = load_instance(name)
        n = xy.shape[0]
        if alg == "GREEDY":
            expect = oracle.nn(xy, 0, wt)[2]
        elif alg == "GREEDY_ITER":
            expect = min(oracle.nn(xy, s, wt)[2] for s in range(n))
        else:
            expect = min(oracle.two_opt(xy, oracle.nn(xy, s, wt)[1], wt, log=False)[2] for s in range(n))
        r = run_cli(TSP_GPU, workdir, ["-f", name + ".tsp", "-q", "-alg", alg])
        assert r.stdout.strip() == "Cost: %.2f" % expect, (name, alg)
        assert calls(r.stderr)["tsp_read_input"] == 1
    # TSPCUDA_WEIGHT=ref: the reference's own rule (tsp.c:629) on a CEIL_2D file -- how bench.py treats pla85900
    xy, _ = load_instance("dsj1000")
    r = run_cli(TSP_GPU, workdir, ["-f", "dsj1000.tsp", "-q", "-alg", "GREEDY"], {"TSPCUDA_WEIGHT": "ref"})
    assert r.stdout.strip() == "Cost: %.2f" % oracle.nn(xy, 0, EUC_REF)[2]
    # GEO (burma14): all-start NN + 2-opt
    from oracle.bindings import GEO
    xy, _ = load_instance("burma14")
    expect = min(oracle.two_opt(xy, oracle.nn(xy, s, GEO)[1], GEO, log=False)[2] for s in range(xy.shape[0]))
    r = run_cli(TSP_GPU, workdir, ["-f", "burma14.tsp", "-q", "-alg", "2OPT_GREEDY"])
    assert r.stdout.strip() == "Cost: %.2f" % expect
    # an EXPLICIT-matrix file is rejected the way the reference rejects a foreign type: no result
    with open(workdir / "explicit.tsp", "w") as f:
        f.write("NAME : x\nTYPE : TSP\nDIMENSION : 3\nEDGE_WEIGHT_TYPE : EXPLICIT\nEDGE_WEIGHT_SECTION\n0 1 2\n1 0 3\n2 3 0\nEOF\n")
    r = run_cli(TSP_GPU, workdir, ["-f", "explicit.tsp", "-q", "-alg", "GREEDY"], check=False)
    assert "Cost:" not in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("batch", ["1", "0"])
def test_mod_costs_through_the_reference_objects(workdir, batch):
    """h_Greedy_2opt_mod_costs (heuristics.c:118-149; its only caller is the CPLEX relaxation callback,
    which cannot run here) reached the way the callback reaches it: build/bin/harness_mod_costs is linked
    like the drop-in CLI (interposer + libtspcuda ahead of the unmodified reference objects).
    TSPCUDA_BATCH=1: the whole-loop override (computes the last start only); TSPCUDA_BATCH=0: the
    reference's own loop over the interposed h_greedyutil / ref_2opt with the foreign matrix.  Both must
    give the result the unmodified reference objects produced (golden_matrix.json)."""
    import json
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from matrices import xstar_matrix
    exe = os.path.join(ROOT, "build", "bin", "harness_mod_costs")
    xstar_matrix(11, 64).tofile(workdir / "m64.f64")
    r = run_cli(exe, workdir, ["64", "m64.f64"], {"TSPCUDA_BATCH": batch})
    with open(os.path.join(ROOT, "tests", "golden", "golden_matrix.json")) as f:
        g = json.load(f)["11_64"]["mod_costs"]
    e, cost, h = r.stdout.split()
    assert e == "0" and float.fromhex(cost).hex() == g["cost"] and h == g["hash"]
    c = calls(r.stderr)
    if batch == "0":
        assert c["h_greedyutil"] == 64 and c["ref_2opt"] == 64
    else:
        assert c["h_greedyutil"] == 0 and c["ref_2opt"] == 0


@pytest.mark.gpu
def test_patching_through_the_reference_objects(workdir, reference):
    """cx_patching (cplex_model.c:693-755; its only caller is the CPLEX-backed Benders loop, :205-212) reached the way
    that loop reaches it: build/bin/harness_patching is linked like the drop-in CLI, so the reference-side call lands
    in the interposer.  Result = what the unmodified reference objects give in this process on the same solution."""
    from components import cycles_solution
    from oracle.bindings import fnv1a
    exe = os.path.join(ROOT, "build", "bin", "harness_patching")
    n, seed = 400, 5
    xy = reference.random_points(n, seed)
    path, comp, nc = cycles_solution(seed, n, 37, True, xy)
    np.concatenate([path, comp]).astype(np.int32).tofile(workdir / "sol.i32")
    r = run_cli(exe, workdir, [str(n), str(seed), "sol.i32"])
    patches, cost, h = r.stdout.split()
    k = sum(float(reference.costs()[i, path[i]]) for i in range(n))
    count = 0
    while nc > 1:
        nc, k = reference.patching(path, comp, nc, k)
        count += 1
    assert int(patches) == count == 36 and float.fromhex(cost) == k
    assert h == fnv1a(np.concatenate([path, comp]).astype(np.int32))
    assert calls(r.stderr)["cx_patching"] == 36


@pytest.mark.gpu
def test_cli_multi_gpu(workdir):
    """TSPCUDA_GPUS: the whole-loop overrides shard their starts over the GPUs of the box (NCCL incumbent
    reduction inside libtspcuda); the Cost line must not change.  More GPUs than visible: loud failure."""
    import travellingsalesmanoptimization_b200 as T
    r = run_cli(TSP_GPU, workdir, ["-f", "berlin52.tsp", "-q", "-alg", "2OPT_GREEDY"], {"TSPCUDA_GPUS": "99"}, check=False)
    assert "Cost:" not in r.stdout or r.returncode != 0
    g = T.device_count()
    if g < 2:
        pytest.skip("needs 2 GPUs")
    for alg in ("GREEDY_ITER", "2OPT_GREEDY"):
        one = run_cli(TSP_GPU, workdir, ["-f", "rl1304.tsp", "-q", "-alg", alg])
        many = run_cli(TSP_GPU, workdir, ["-f", "rl1304.tsp", "-q", "-alg", alg], {"TSPCUDA_GPUS": "all"})
        assert one.stdout == many.stdout and "Cost:" in one.stdout
        assert calls(many.stderr)["gpus"] == g and calls(one.stderr)["gpus"] == 1


@pytest.mark.gpu
def test_pla85900_opens_through_the_unchanged_driver(workdir):
    """The headline instance through the drop-in CLI: CEIL_2D header (the reference reader stops there,
    tsp.c:576-583), n > 46340 (no host matrix: the reference's int n*n overflows, tsp.c:614).
    TSPCUDA_WEIGHT=ref = the reference's own cost rule, as in bench.py: NN(0) costs 163503308
    (the start cost of the bench descent); with the file's own CEIL_2D rule the oracle's value."""
    from oracle.bindings import CEIL_2D, Oracle
    xy, ewt = load_instance("pla85900")
    assert ewt == "CEIL_2D"
    write_tsplib(workdir / "pla85900.tsp", "pla85900", xy, ewt)
    r = run_cli(TSP_GPU, workdir, ["-f", "pla85900.tsp", "-q", "-alg", "GREEDY"], {"TSPCUDA_WEIGHT": "ref"})
    assert r.stdout.strip() == "Cost: 163503308.00"
    r = run_cli(TSP_GPU, workdir, ["-f", "pla85900.tsp", "-q", "-alg", "GREEDY"])
    assert r.stdout.strip() == "Cost: %.2f" % Oracle().nn(xy, 0, CEIL_2D)[2]
