"""ctypes bindings for the parity oracle -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (travellingsalesmanoptimization_b200) never does.

  Oracle     -> oracle/liboracle.so      CPU restatement (oracle/onfly.c), always available
  Reference  -> oracle/_ref/librefapi.so the UNMODIFIED reference objects behind oracle/ref_api.c
                (built here from /root/reference by `make -C oracle ref`; travels to the GPU box
                as a built artefact; absent => Reference.available() is False)
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
EUC_REF, CEIL_2D, ATT, GEO = 0, 1, 2, 3
WT_BY_NAME = {"EUC_2D": EUC_REF, "EUC_REF": EUC_REF, "CEIL_2D": CEIL_2D, "ATT": ATT, "GEO": GEO}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_long)


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip) if a is not None else None


def build_oracle(ref=False):
    """Compile liboracle.so (and oracle/_ref when asked and /root/reference is present)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "all"])
    if ref and os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def fnv1a(path):
    h = 0xCBF29CE484222325
    for v in np.asarray(path, dtype=np.uint32).tolist():
        h ^= v
        h = (h * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h


def load_instance(name):
    """(xy float64 (n,2) C-contiguous, TSPLIB edge weight type string) from tests/golden/instances."""
    z = np.load(os.path.join(HERE, "..", "tests", "golden", "instances", name + ".npz"))
    return np.ascontiguousarray(z["xy"], dtype=np.float64), str(z["edge_weight_type"])


class Oracle:
    def __init__(self):
        so = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(so):
            build_oracle()
        L = self.L = C.CDLL(so)
        L.orc_edge_cost.restype = C.c_double
        L.orc_edge_cost.argtypes = [C.c_int] + [C.c_double] * 4
        L.orc_fnv1a.restype = C.c_uint64
        L.orc_fnv1a.argtypes = [_ip, C.c_int]
        L.orc_random_points.argtypes = [C.c_int, C.c_int, _dp]
        L.orc_cost_matrix.argtypes = [_dp, C.c_int, C.c_int, _dp]
        L.orc_is_tour.argtypes = [_ip, C.c_int]
        L.orc_tour_cost.restype = C.c_double
        L.orc_tour_cost.argtypes = [_dp, C.c_int, C.c_int, _ip]
        L.orc_nn.argtypes = [_dp, C.c_int, C.c_int, C.c_int, _ip, _dp]
        L.orc_nn_all.argtypes = [_dp, C.c_int, C.c_int, C.c_int, C.c_int, _ip, _dp, _dp]
        L.orc_two_opt_scan.restype = C.c_double
        L.orc_two_opt_scan.argtypes = [_dp, C.c_int, C.c_int, _ip, _ip, _ip]
        L.orc_two_opt_scan_rows.restype = C.c_longlong
        L.orc_two_opt_scan_rows.argtypes = [_dp, C.c_int, C.c_int, _ip, C.c_int, C.c_int, _ip, _ip, _dp]
        L.orc_two_opt_once.restype = C.c_double
        L.orc_two_opt_once.argtypes = [_dp, C.c_int, C.c_int, _ip, _dp, _ip, _ip]
        L.orc_two_opt.restype = C.c_long
        L.orc_two_opt.argtypes = [_dp, C.c_int, C.c_int, _ip, _dp, C.c_long, _ip, _ip, _dp]
        L.orc_two_opt_matrix.restype = C.c_long
        L.orc_two_opt_matrix.argtypes = [_dp, C.c_int, _ip, _dp, C.c_long, _ip, _ip, _dp]
        L.orc_extra_mileage.argtypes = [_dp, C.c_int, C.c_int, C.c_int, C.c_int, _ip, _dp]
        L.orc_nn_matrix.argtypes = [_dp, C.c_int, C.c_int, _ip, _dp]
        L.orc_mod_costs.argtypes = [_dp, C.c_int, _ip, _dp]
        L.orc_tabu_move.argtypes = [_dp, C.c_int, C.c_int, _ip, _dp, _ip, C.c_int, C.c_int, _ip, _ip, _dp]
        L.orc_tabu_init.argtypes = [C.c_int, _ip, _ip, _ip]
        L.orc_tabu_linear_step.argtypes = [_ip, C.c_int, C.c_int, _ip]
        L.orc_vns_kick.argtypes = [C.c_int, _ip]
        L.orc_patching.argtypes = [_dp, C.c_int, C.c_int, _ip, _ip, _ip, _dp, _ip, _ip, _dp]
        L.orc_audit_moves.restype = C.c_long
        L.orc_audit_moves.argtypes = [_dp, C.c_int, C.c_int, _ip, C.c_long, _ip, _ip, _dp, _lp, C.c_long]
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]

    # -- helpers -------------------------------------------------------------------------------
    def edge(self, wt, xi, yi, xj, yj):
        return self.L.orc_edge_cost(wt, xi, yi, xj, yj)

    def random_points(self, n, seed):
        xy = np.empty((n, 2), dtype=np.float64)
        self.L.orc_random_points(n, seed, _d(xy))
        return xy

    def cost_matrix(self, xy, wt=EUC_REF):
        n = xy.shape[0]
        out = np.empty((n, n), dtype=np.float64)
        assert self.L.orc_cost_matrix(_d(xy), n, wt, _d(out)) == 0
        return out

    def is_tour(self, path):
        path = np.ascontiguousarray(path, dtype=np.int32)
        return bool(self.L.orc_is_tour(_i(path), path.shape[0]))

    def tour_cost(self, xy, path, wt=EUC_REF):
        path = np.ascontiguousarray(path, dtype=np.int32)
        return self.L.orc_tour_cost(_d(xy), xy.shape[0], wt, _i(path))

    def nn(self, xy, start, wt=EUC_REF):
        n = xy.shape[0]
        path = np.full(n, -7, dtype=np.int32)
        cost = C.c_double(0)
        e = self.L.orc_nn(_d(xy), n, wt, start, _i(path), C.byref(cost))
        return e, path, cost.value

    def nn_all(self, xy, s0, count, wt=EUC_REF):
        n = xy.shape[0]
        best_path = np.zeros(n, dtype=np.int32)
        costs = np.zeros(count, dtype=np.float64)
        bc = C.c_double(0)
        best = self.L.orc_nn_all(_d(xy), n, wt, s0, count, _i(best_path), C.byref(bc), _d(costs))
        return best, best_path, bc.value, costs

    def two_opt_scan(self, xy, path, wt=EUC_REF):
        a, b = C.c_int(-1), C.c_int(-1)
        path = np.ascontiguousarray(path, dtype=np.int32)
        d = self.L.orc_two_opt_scan(_d(xy), xy.shape[0], wt, _i(path), C.byref(a), C.byref(b))
        return d, a.value, b.value

    def two_opt_scan_rows(self, xy, path, a_lo, a_hi, wt=EUC_REF):
        """Rows [a_lo, a_hi) of one scan: returns (pair iterations covered, delta, a, b)."""
        a, b, d = C.c_int(-1), C.c_int(-1), C.c_double(0)
        path = np.ascontiguousarray(path, dtype=np.int32)
        pairs = self.L.orc_two_opt_scan_rows(_d(xy), xy.shape[0], wt, _i(path), a_lo, a_hi, C.byref(a), C.byref(b), C.byref(d))
        return pairs, d.value, a.value, b.value

    def two_opt_once(self, xy, path, cost, wt=EUC_REF):
        """ref_2opt_once in place on path (int32 array): returns (best_delta, new cost, a, b)."""
        c = C.c_double(cost)
        a, b = C.c_int(-1), C.c_int(-1)
        d = self.L.orc_two_opt_once(_d(xy), xy.shape[0], wt, _i(path), C.byref(c), C.byref(a), C.byref(b))
        return d, c.value, a.value, b.value

    def two_opt(self, xy, path, wt=EUC_REF, max_moves=-1, log=True):
        """Returns (moves, path, cost, log_a, log_b, log_delta)."""
        n = xy.shape[0]
        path = np.array(path, dtype=np.int32, copy=True)
        cap = max_moves if max_moves >= 0 else 4 * n + 64
        la = np.zeros(cap, dtype=np.int32) if log else None
        lb = np.zeros(cap, dtype=np.int32) if log else None
        ld = np.zeros(cap, dtype=np.float64) if log else None
        cost = C.c_double(0)
        if max_moves < 0 and log:
            # unknown move count: run capped chunks so the log cannot overflow
            total, cur = 0, cap
            m = self.L.orc_two_opt(_d(xy), n, wt, _i(path), C.byref(cost), cur, _i(la), _i(lb), _d(ld))
            total = m
            while m == cur:   # cap hit: continue with a fresh log chunk
                la2 = np.zeros(cap, dtype=np.int32); lb2 = np.zeros(cap, dtype=np.int32); ld2 = np.zeros(cap, dtype=np.float64)
                m = self.L.orc_two_opt(_d(xy), n, wt, _i(path), C.byref(cost), cur, _i(la2), _i(lb2), _d(ld2))
                la = np.concatenate([la[:total], la2]); lb = np.concatenate([lb[:total], lb2]); ld = np.concatenate([ld[:total], ld2])
                total += m
            return total, path, cost.value, la[:total], lb[:total], ld[:total]
        m = self.L.orc_two_opt(_d(xy), n, wt, _i(path), C.byref(cost), max_moves, _i(la), _i(lb),
                               _d(ld) if log else None)
        if log:
            return m, path, cost.value, la[:m], lb[:m], ld[:m]
        return m, path, cost.value, None, None, None

    def two_opt_matrix(self, costs, path, max_moves=-1):
        n = costs.shape[0]
        path = np.array(path, dtype=np.int32, copy=True)
        cap = max_moves if max_moves >= 0 else 16 * n + 64
        la = np.zeros(cap, dtype=np.int32); lb = np.zeros(cap, dtype=np.int32); ld = np.zeros(cap, dtype=np.float64)
        cost = C.c_double(0)
        m = self.L.orc_two_opt_matrix(_d(costs), n, _i(path), C.byref(cost), cap, _i(la), _i(lb), _d(ld))
        return m, path, cost.value, la[:m], lb[:m], ld[:m]

    def extra_mileage(self, xy, node_a=-1, node_b=-1, wt=EUC_REF):
        """h_ExtraMileage (node_a < 0: EM_MAX initial edge) / h_extramileage_util: (path, cost)."""
        n = xy.shape[0]
        path = np.zeros(n, dtype=np.int32)
        c = C.c_double(0)
        e = self.L.orc_extra_mileage(_d(xy), n, wt, node_a, node_b, _i(path), C.byref(c))
        assert e == 0, e
        return path, c.value

    def nn_matrix(self, costs, start):
        """h_greedyutil on an explicit matrix: (code, path, cost)."""
        n = costs.shape[0]
        path = np.full(n, -7, dtype=np.int32)
        c = C.c_double(0)
        e = self.L.orc_nn_matrix(_d(costs), n, start, _i(path), C.byref(c))
        return e, path, c.value

    def mod_costs(self, costs):
        """h_Greedy_2opt_mod_costs (no time limit): (path, cost) = the last start's NN + 2-opt."""
        n = costs.shape[0]
        path = np.full(n, -7, dtype=np.int32)
        c = C.c_double(0)
        self.L.orc_mod_costs(_d(costs), n, _i(path), C.byref(c))
        return path, c.value

    def tabu_move(self, xy, path, cost, tabu_list, tenure, it, wt=EUC_REF):
        """In-place on path / tabu_list (int32 arrays). Returns (new cost, a, b, delta)."""
        c = C.c_double(cost)
        a, b, d = C.c_int(-1), C.c_int(-1), C.c_double(0)
        self.L.orc_tabu_move(_d(xy), xy.shape[0], wt, _i(path), C.byref(c), _i(tabu_list), tenure, it,
                             C.byref(a), C.byref(b), C.byref(d))
        return c.value, a.value, b.value, d.value

    def tabu_init(self, n):
        t, lo, hi = C.c_int(), C.c_int(), C.c_int()
        self.L.orc_tabu_init(n, C.byref(t), C.byref(lo), C.byref(hi))
        return t.value, lo.value, hi.value

    def tabu_linear_step(self, tenure, lo, hi, increment):
        t, inc = C.c_int(tenure), C.c_int(1 if increment else 0)
        self.L.orc_tabu_linear_step(C.byref(t), lo, hi, C.byref(inc))
        return t.value, bool(inc.value)

    def vns_kick(self, path):
        self.L.orc_vns_kick(path.shape[0], _i(path))

    def patching(self, xy, path, comp, ncomp, cost, wt=EUC_REF):
        """cx_patching, one patch in place on path / comp: (applied, ncomp, cost, a, b, delta)."""
        nc, c = C.c_int(ncomp), C.c_double(cost)
        a, b, d = C.c_int(-1), C.c_int(-1), C.c_double(0)
        r = self.L.orc_patching(_d(xy), xy.shape[0], wt, _i(path), _i(comp), C.byref(nc), C.byref(c), C.byref(a),
                                C.byref(b), C.byref(d))
        return bool(r), nc.value, c.value, a.value, b.value, d.value

    def audit_moves(self, xy, path, log_a, log_b, log_delta, full_at, wt=EUC_REF):
        path = np.array(path, dtype=np.int32, copy=True)
        la = np.ascontiguousarray(log_a, dtype=np.int32); lb = np.ascontiguousarray(log_b, dtype=np.int32)
        ld = np.ascontiguousarray(log_delta, dtype=np.float64)
        fa = np.ascontiguousarray(sorted(full_at), dtype=np.int64)
        r = self.L.orc_audit_moves(_d(xy), xy.shape[0], wt, _i(path), la.shape[0], _i(la), _i(lb), _d(ld),
                                   fa.ctypes.data_as(_lp), fa.shape[0])
        return r, path

    def num_threads(self):
        return self.L.orc_num_threads()

    def set_num_threads(self, t):
        self.L.orc_set_num_threads(t)


class Reference:
    """The unmodified reference objects (oracle/_ref/librefapi.so). Single global instance state."""

    SO = os.path.join(HERE, "_ref", "librefapi.so")

    @classmethod
    def available(cls):
        return os.path.exists(cls.SO)

    def __init__(self):
        if not self.available():
            raise RuntimeError("oracle/_ref/librefapi.so not built (run `make -C oracle ref` where /root/reference exists)")
        L = self.L = C.CDLL(self.SO)
        L.refapi_set_points.argtypes = [_dp, C.c_int]
        L.refapi_random_points.argtypes = [C.c_int, C.c_int]
        L.refapi_get_points.argtypes = [_dp]
        L.refapi_costs.restype = _dp
        L.refapi_set_timelimit.argtypes = [C.c_double]
        L.refapi_srand.argtypes = [C.c_uint]
        L.refapi_nn.argtypes = [C.c_int, _ip, _dp]
        L.refapi_two_opt_once.restype = C.c_double
        L.refapi_two_opt_once.argtypes = [_ip, _dp]
        L.refapi_two_opt.argtypes = [_ip, _dp, C.c_int, _ip, _ip, _dp]
        L.refapi_ref_2opt.argtypes = [_ip, _dp, C.c_int]
        L.refapi_tabu_move.argtypes = [_ip, _dp, _ip, C.c_int, C.c_int]
        L.refapi_extra_mileage_from.argtypes = [C.c_int, C.c_int, _ip, _dp]
        L.refapi_nn_costs.argtypes = [C.c_int, _dp, _ip, _dp]
        L.refapi_ref_2opt_costs.argtypes = [_dp, _ip, _dp]
        L.refapi_mod_costs.argtypes = [_dp, _ip, _dp]
        L.refapi_tabu_init.argtypes = [C.c_int, _ip, _ip, _ip]
        L.refapi_tabu_linear_step.argtypes = [_ip, C.c_int, C.c_int, _ip]
        L.refapi_vns_kick.argtypes = [_ip]
        L.refapi_patching.argtypes = [_ip, _ip, _ip, _dp]
        L.refapi_run.argtypes = [C.c_int, C.c_int, _ip, _dp, _ip]
        L.refapi_reverse_calls.restype = C.c_long
        L.refapi_last_move.argtypes = [_ip, _ip]
        self.n = 0

    def set_points(self, xy):
        xy = np.ascontiguousarray(xy, dtype=np.float64)
        e = self.L.refapi_set_points(_d(xy), xy.shape[0])
        assert e == 0, e
        self.n = xy.shape[0]

    def random_points(self, n, seed):
        assert self.L.refapi_random_points(n, seed) == 0
        self.n = n
        xy = np.empty((n, 2), dtype=np.float64)
        self.L.refapi_get_points(_d(xy))
        return xy

    def costs(self):
        p = self.L.refapi_costs()
        return np.ctypeslib.as_array(p, shape=(self.n, self.n))

    def nn(self, start):
        path = np.full(self.n, -7, dtype=np.int32)
        c = C.c_double(0)
        e = self.L.refapi_nn(start, _i(path), C.byref(c))
        return e, path, c.value

    def two_opt_once(self, path, cost):
        c = C.c_double(cost)
        d = self.L.refapi_two_opt_once(_i(path), C.byref(c))
        a, b = C.c_int(), C.c_int()
        self.L.refapi_last_move(C.byref(a), C.byref(b))
        return d, c.value, a.value, b.value

    def two_opt(self, path, max_moves=-1):
        path = np.array(path, dtype=np.int32, copy=True)
        cap = max_moves if max_moves >= 0 else 16 * self.n + 64
        la = np.zeros(cap, dtype=np.int32); lb = np.zeros(cap, dtype=np.int32); ld = np.zeros(cap, dtype=np.float64)
        c = C.c_double(0)
        m = self.L.refapi_two_opt(_i(path), C.byref(c), cap if max_moves < 0 else max_moves, _i(la), _i(lb), _d(ld))
        return m, path, c.value, la[:m], lb[:m], ld[:m]

    def extra_mileage_from(self, node_a, node_b):
        path = np.zeros(self.n, dtype=np.int32)
        c = C.c_double(0)
        e = self.L.refapi_extra_mileage_from(node_a, node_b, _i(path), C.byref(c))
        return e, path, c.value

    def nn_costs(self, costs, start):
        path = np.full(self.n, -7, dtype=np.int32)
        c = C.c_double(0)
        e = self.L.refapi_nn_costs(start, _d(costs), _i(path), C.byref(c))
        return e, path, c.value

    def ref_2opt_costs(self, costs, path):
        path = np.array(path, dtype=np.int32, copy=True)
        c = C.c_double(0)
        e = self.L.refapi_ref_2opt_costs(_d(costs), _i(path), C.byref(c))
        return e, path, c.value

    def mod_costs(self, costs):
        path = np.full(self.n, -7, dtype=np.int32)
        c = C.c_double(0)
        e = self.L.refapi_mod_costs(_d(costs), _i(path), C.byref(c))
        return e, path, c.value

    def tabu_move(self, path, cost, tabu_list, tenure, it):
        c = C.c_double(cost)
        e = self.L.refapi_tabu_move(_i(path), C.byref(c), _i(tabu_list), tenure, it)
        assert e == 0
        return c.value

    def tabu_init(self, n):
        t, lo, hi = C.c_int(), C.c_int(), C.c_int()
        self.L.refapi_tabu_init(n, C.byref(t), C.byref(lo), C.byref(hi))
        return t.value, lo.value, hi.value

    def tabu_linear_step(self, tenure, lo, hi, increment):
        t, inc = C.c_int(tenure), C.c_int(1 if increment else 0)
        self.L.refapi_tabu_linear_step(C.byref(t), lo, hi, C.byref(inc))
        return t.value, bool(inc.value)

    def srand(self, s):
        self.L.refapi_srand(s)

    def vns_kick(self, path):
        assert self.L.refapi_vns_kick(_i(path)) == 0

    def patching(self, path, comp, ncomp, cost):
        """The reference's cx_patching (cplex_model.c:693), one patch in place: (ncomp, cost)."""
        nc, c = C.c_int(ncomp), C.c_double(cost)
        self.L.refapi_patching(_i(path), _i(comp), C.byref(nc), C.byref(c))
        return nc.value, c.value

    def run(self, alg, k=2**31 - 1):
        """alg: 0 GREEDY, 1 GREEDY_ITER, 2 2OPT_GREEDY, 3 TABU_SEARCH, 4 VNS, 5 EXTRA_MILEAGE (EM_MAX). Needs ./results/ for 3/4."""
        path = np.zeros(self.n, dtype=np.int32)
        c, s = C.c_double(0), C.c_int(0)
        e = self.L.refapi_run(alg, k, _i(path), C.byref(c), C.byref(s))
        return e, path, c.value, s.value
