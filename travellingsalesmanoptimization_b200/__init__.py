"""travellingsalesmanoptimization_b200 -- host-side Python mirror of libtspcuda.so's C ABI (include/tspcuda.h).

The product is the shared library (hand-written sm_100a CUDA behind a plain C ABI); this module
is a thin ctypes binding used by tests/, bench.py and __graft_entry__.py.  It contains no
algorithm and no CPU fallback: if the library or a CUDA device is missing, calls raise.
Names mirror the reference helpers they replace (see include/tspcuda.h for file:line):
  Instance.cost_matrix()   tsp_compute_costs      Instance.nn(start)       h_greedyutil
  Instance.two_opt(path)   ref_2opt               Instance.two_opt_once()  ref_2opt_once
  Instance.tabu_move()     tabu_best_move         Instance.multistart_two_opt()  h_greedy_2opt loop
"""
import ctypes as C
import os

import numpy as np

from .build import LIB, build
from .tsplib import load_instance, read_tsplib  # noqa: F401  (data loading for bench.py / scripts)

EUC_REF, CEIL_2D, ATT, GEO = 0, 1, 2, 3
SCAN_AUTO, SCAN_FULL, SCAN_PRUNED = 0, 1, 2
WT_BY_NAME = {"EUC_2D": EUC_REF, "EUC_REF": EUC_REF, "CEIL_2D": CEIL_2D, "ATT": ATT, "GEO": GEO}
ERROR_NAMES = {0: "T_OK", 1: "CANCELLED", 3: "INVALID_ARGUMENT", 4: "DEADLINE_EXCEEDED", 8: "RESOURCE_EXHAUSTED", 9: "FAILED_PRECONDITION",
               11: "OUT_OF_RANGE", 12: "UNIMPLEMENTED", 13: "INTERNAL", 14: "UNAVAILABLE"}
OK, DEADLINE_EXCEEDED, INVALID_ARGUMENT, OUT_OF_RANGE, UNAVAILABLE = 0, 4, 3, 11, 14
FAILED_PRECONDITION = 9

EXPORTS = [
    "tspcuda_device_count", "tspcuda_last_error", "tspcuda_version", "tspcuda_create", "tspcuda_destroy",
    "tspcuda_nnodes", "tspcuda_get_stats", "tspcuda_reset_stats", "tspcuda_cost_matrix", "tspcuda_cost_matrix_device",
    "tspcuda_cost_matrix_checksum", "tspcuda_nn", "tspcuda_nn_multi", "tspcuda_two_opt", "tspcuda_two_opt_once",
    "tspcuda_tabu_move", "tspcuda_tour_upload", "tspcuda_tour_descend", "tspcuda_tour_download", "tspcuda_tour_movelog",
    "tspcuda_multistart_two_opt", "tspcuda_pipe_rate", "tspcuda_selfcheck_sqrt", "tspcuda_timer_start",
    "tspcuda_timer_stop", "tspcuda_flush_l2", "tspcuda_set_scan_mode", "tspcuda_tabu_search",
    "tspcuda_extra_mileage", "tspcuda_mx_create", "tspcuda_mx_destroy", "tspcuda_mx_get_stats", "tspcuda_mx_nn", "tspcuda_mx_two_opt",
    "tspcuda_mx_two_opt_once", "tspcuda_patching", "tspcuda_mgpu_create", "tspcuda_mgpu_destroy", "tspcuda_mgpu_ndev",
    "tspcuda_mgpu_multistart", "tspcuda_vns_search", "tspcuda_restarts", "tspcuda_mgpu_restarts", "tspcuda_mx_update",
]
RESTART_TABU, RESTART_VNS = 0, 1


class TspCudaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"tspcuda error {code} ({ERROR_NAMES.get(code, '?')}): {msg}")
        self.code = code


class Move(C.Structure):
    _fields_ = [("a", C.c_int32), ("b", C.c_int32), ("delta", C.c_int32), ("reserved", C.c_int32)]


class MoveF(C.Structure):
    _fields_ = [("a", C.c_int32), ("b", C.c_int32), ("delta", C.c_double)]


class Stats(C.Structure):
    _fields_ = [("scans", C.c_int64), ("moves", C.c_int64), ("pair_evals", C.c_int64), ("kernel_launches", C.c_int64),
                ("scan_kernel_ms", C.c_double), ("last_kernel_ms", C.c_double), ("h2d_bytes", C.c_int64),
                ("d2h_bytes", C.c_int64), ("pruned_scans", C.c_int64), ("candidate_evals", C.c_int64),
                ("pruned_kernel_ms", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_dp, _ip, _vp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p
_i64p = C.POINTER(C.c_int64)
_lib = None


def lib():
    """Load libtspcuda.so, (re)building it first when it is missing or older than its sources. Raises if impossible."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("TSPCUDA_LIB") or LIB      # development override: an alternative build of the SAME library
    if path == LIB:
        build()                                      # returns at once unless needs_build()
    L = C.CDLL(path)
    L.tspcuda_last_error.restype = C.c_char_p
    L.tspcuda_version.restype = C.c_char_p
    L.tspcuda_create.argtypes = [C.POINTER(_vp), _dp, C.c_int, C.c_int, C.c_int]
    L.tspcuda_destroy.argtypes = [_vp]
    L.tspcuda_destroy.restype = None
    L.tspcuda_nnodes.argtypes = [_vp]
    L.tspcuda_get_stats.argtypes = [_vp, C.POINTER(Stats)]
    L.tspcuda_reset_stats.argtypes = [_vp]
    L.tspcuda_cost_matrix.argtypes = [_vp, _dp]
    L.tspcuda_cost_matrix_device.argtypes = [_vp, _dp]
    L.tspcuda_cost_matrix_checksum.argtypes = [_vp, _i64p]
    L.tspcuda_nn.argtypes = [_vp, C.c_int, _ip, _dp]
    L.tspcuda_nn_multi.argtypes = [_vp, C.c_int, C.c_int, C.c_int, C.c_double, _dp, _ip, _ip, _dp, _ip]
    L.tspcuda_two_opt.argtypes = [_vp, _ip, _dp, C.c_int64, C.c_double, _i64p, C.POINTER(Move), C.c_int64]
    L.tspcuda_two_opt_once.argtypes = [_vp, _ip, _dp, _dp, C.POINTER(Move)]
    L.tspcuda_tabu_move.argtypes = [_vp, _ip, _dp, _ip, C.c_int, C.c_int, C.POINTER(Move)]
    L.tspcuda_tour_upload.argtypes = [_vp, _ip]
    L.tspcuda_tour_descend.argtypes = [_vp, C.c_int64, _i64p, _ip]
    L.tspcuda_tour_download.argtypes = [_vp, _ip, _dp]
    L.tspcuda_tour_movelog.argtypes = [_vp, C.POINTER(Move), C.c_int64, _i64p]
    L.tspcuda_multistart_two_opt.argtypes = [_vp, C.c_int, C.c_int, C.c_int, C.c_double, _dp, _dp, _i64p, _ip, _ip, _dp, _ip]
    L.tspcuda_restarts.argtypes = [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_double, _ip, _ip, _dp, _ip]
    L.tspcuda_mgpu_restarts.argtypes = [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_double, _ip, _ip, _dp, _ip, _dp, _dp]
    L.tspcuda_pipe_rate.argtypes = [_vp, C.c_int, _dp]
    L.tspcuda_selfcheck_sqrt.argtypes = [_vp, _i64p]
    L.tspcuda_timer_start.argtypes = [_vp]
    L.tspcuda_timer_stop.argtypes = [_vp, _dp]
    L.tspcuda_flush_l2.argtypes = [_vp]
    L.tspcuda_set_scan_mode.argtypes = [_vp, C.c_int]
    L.tspcuda_extra_mileage.argtypes = [_vp, C.c_int, C.c_int, _ip, _dp, _ip, _ip]
    L.tspcuda_patching.argtypes = [_vp, _ip, _ip, _ip, _dp, C.c_int, C.POINTER(Move), C.c_int64, _ip]
    L.tspcuda_vns_search.argtypes = [_vp, _ip, _dp, C.c_int, C.c_double, _dp, _ip, _dp, _ip, _ip]
    L.tspcuda_mgpu_create.argtypes = [C.POINTER(_vp), _dp, C.c_int, C.c_int, C.c_int, _ip]
    L.tspcuda_mgpu_destroy.argtypes = [_vp]
    L.tspcuda_mgpu_destroy.restype = None
    L.tspcuda_mgpu_ndev.argtypes = [_vp]
    L.tspcuda_mgpu_multistart.argtypes = [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, _ip, _ip, _dp, _ip, _dp, _dp]
    L.tspcuda_mx_create.argtypes = [C.POINTER(_vp), _dp, C.c_int, C.c_int]
    L.tspcuda_mx_destroy.argtypes = [_vp]
    L.tspcuda_mx_destroy.restype = None
    L.tspcuda_mx_update.argtypes = [_vp, _dp]
    L.tspcuda_mx_get_stats.argtypes = [_vp, C.POINTER(Stats)]
    L.tspcuda_mx_nn.argtypes = [_vp, C.c_int, _ip, _dp, _ip]
    L.tspcuda_mx_two_opt.argtypes = [_vp, _ip, _dp, C.c_int64, C.c_double, _i64p, C.POINTER(MoveF), C.c_int64]
    L.tspcuda_mx_two_opt_once.argtypes = [_vp, _ip, _dp, _dp]
    L.tspcuda_tabu_search.argtypes = [_vp, _ip, _dp, _ip, _ip, C.c_int, C.c_int, C.c_double, _dp, _ip, _dp, _ip]
    _lib = L
    return L


def device_count():
    return lib().tspcuda_device_count()


def _chk(code, allow=()):
    if code != 0 and code not in allow:
        raise TspCudaError(code, lib().tspcuda_last_error().decode())
    return code


def _ipc(a):
    return a.ctypes.data_as(_ip)


def _dpc(a):
    return a.ctypes.data_as(_dp)


class Instance:
    """One TSP instance resident on one GPU (tspcuda_ctx)."""

    def __init__(self, xy, weight_type=EUC_REF, device=0):
        self.xy = np.ascontiguousarray(xy, dtype=np.float64)
        assert self.xy.ndim == 2 and self.xy.shape[1] == 2
        self.n = self.xy.shape[0]
        self._h = _vp()
        _chk(lib().tspcuda_create(C.byref(self._h), _dpc(self.xy), self.n, weight_type, device))

    def close(self):
        if self._h:
            lib().tspcuda_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- stats -----------------------------------------------------------------------------------
    def stats(self):
        s = Stats()
        _chk(lib().tspcuda_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def reset_stats(self):
        _chk(lib().tspcuda_reset_stats(self._h))

    def set_scan_mode(self, mode):
        """SCAN_AUTO / SCAN_FULL (every pair, as the reference) / SCAN_PRUNED (exact candidate pruning)."""
        _chk(lib().tspcuda_set_scan_mode(self._h, mode))

    # -- K1 ----------------------------------------------------------------------------------------
    def cost_matrix(self):
        out = np.empty((self.n, self.n), dtype=np.float64)
        _chk(lib().tspcuda_cost_matrix(self._h, _dpc(out)))
        return out

    def cost_matrix_device(self):
        ms = C.c_double(0)
        _chk(lib().tspcuda_cost_matrix_device(self._h, C.byref(ms)))
        return ms.value

    def cost_matrix_checksum(self):
        s = C.c_int64(0)
        _chk(lib().tspcuda_cost_matrix_checksum(self._h, C.byref(s)))
        return s.value

    # -- NN ----------------------------------------------------------------------------------------
    def nn(self, start):
        path = np.full(self.n, -7, dtype=np.int32)
        cost = C.c_double(0)
        code = lib().tspcuda_nn(self._h, start, _ipc(path), C.byref(cost))
        return code, path, cost.value

    def nn_multi(self, first_start, nstarts, stride=1, time_limit=None):
        """h_Greedy_iterative's loop.  Without time_limit: (best_start, best_path, best_cost, costs); with one (seconds):
        (code, starts_done, best_start, best_path, best_cost, costs) -- code DEADLINE_EXCEEDED = best of the prefix taken."""
        costs = np.zeros(nstarts, dtype=np.float64)
        best_path = np.zeros(self.n, dtype=np.int32)
        bs, bc, done = C.c_int(-1), C.c_double(0), C.c_int(0)
        code = _chk(lib().tspcuda_nn_multi(self._h, first_start, nstarts, stride, -1.0 if time_limit is None else time_limit, _dpc(costs),
                                           C.byref(bs), _ipc(best_path), C.byref(bc), C.byref(done)), allow=(DEADLINE_EXCEEDED,))
        if time_limit is None:
            return bs.value, best_path, bc.value, costs
        return code, done.value, bs.value, best_path, bc.value, costs

    # -- 2-opt -------------------------------------------------------------------------------------
    def two_opt(self, path, max_moves=-1, time_limit=-1.0, log_cap=0):
        """ref_2opt: returns (code, moves, path, cost, log[(a,b,delta)...])."""
        path = np.array(path, dtype=np.int32, copy=True)
        cost, moves = C.c_double(0), C.c_int64(0)
        log = (Move * max(1, log_cap))()
        code = _chk(lib().tspcuda_two_opt(self._h, _ipc(path), C.byref(cost), max_moves, time_limit, C.byref(moves),
                                          log if log_cap else None, log_cap), allow=(DEADLINE_EXCEEDED,))
        k = min(moves.value, log_cap)
        return code, moves.value, path, cost.value, [(log[i].a, log[i].b, log[i].delta) for i in range(k)]

    def two_opt_once(self, path, cost):
        """ref_2opt_once in place on `path` (int32 array): returns (best_delta, new cost, (a,b))."""
        c, d, mv = C.c_double(cost), C.c_double(0), Move()
        _chk(lib().tspcuda_two_opt_once(self._h, _ipc(path), C.byref(c), C.byref(d), C.byref(mv)))
        return d.value, c.value, (mv.a, mv.b)

    def tabu_move(self, path, cost, tabu_list, tenure, it):
        """tabu_best_move in place on path / tabu_list: returns (new cost, (a,b,delta))."""
        c, mv = C.c_double(cost), Move()
        _chk(lib().tspcuda_tabu_move(self._h, _ipc(path), C.byref(c), _ipc(tabu_list), tenure, it, C.byref(mv)))
        return c.value, (mv.a, mv.b, mv.delta)

    def tabu_search(self, path, cost, tabu_list, tenures, first_iter=0, incumbent_cost=None):
        """Device-resident tabu loop (mh_TabuSearch's k-loop).  path / tabu_list are updated in place.
        Returns (cost, trace[k], best_iter, best_path or None, best_cost)."""
        tenures = np.ascontiguousarray(tenures, dtype=np.int32)
        k = tenures.shape[0]
        c = C.c_double(cost)
        trace = np.zeros(k, dtype=np.float64)
        best_path = np.zeros(self.n, dtype=np.int32)
        bc, bi = C.c_double(0), C.c_int(-1)
        inc = cost if incumbent_cost is None else incumbent_cost
        _chk(lib().tspcuda_tabu_search(self._h, _ipc(path), C.byref(c), _ipc(tabu_list), _ipc(tenures), first_iter, k,
                                       inc, _dpc(trace), _ipc(best_path), C.byref(bc), C.byref(bi)))
        return c.value, trace, bi.value, (best_path if bi.value >= 0 else None), bc.value

    def extra_mileage(self, node_a=-1, node_b=-1):
        """h_ExtraMileage (node_a < 0: EM_MAX initial edge) / h_extramileage_util: (path, cost, (A, B))."""
        path = np.zeros(self.n, dtype=np.int32)
        cost, a, b = C.c_double(0), C.c_int(-1), C.c_int(-1)
        _chk(lib().tspcuda_extra_mileage(self._h, node_a, node_b, _ipc(path), C.byref(cost), C.byref(a), C.byref(b)))
        return path, cost.value, (a.value, b.value)

    def vns_search(self, path, n_iters, best_cost, time_limit=-1.0):
        """The loop of mh_VNS (metaheuristic.c:277-321), kicks drawn from the process's rand() stream.
        Returns (code, path, cost, trace, best_iter, best_path or None, best_cost, iters_done)."""
        path = np.array(path, dtype=np.int32, copy=True)
        trace = np.zeros(max(1, n_iters), dtype=np.float64)
        best_path = np.zeros(self.n, dtype=np.int32)
        c, bc, bi, done = C.c_double(0), C.c_double(best_cost), C.c_int(-1), C.c_int(0)
        code = _chk(lib().tspcuda_vns_search(self._h, _ipc(path), C.byref(c), n_iters, time_limit, _dpc(trace), _ipc(best_path),
                                             C.byref(bc), C.byref(bi), C.byref(done)), allow=(DEADLINE_EXCEEDED,))
        return code, path, c.value, trace[:done.value], bi.value, (best_path if bi.value >= 0 else None), bc.value, done.value

    def patching(self, path, comp, ncomp, cost, max_patches=-1, allow=()):
        """cx_patching (cplex_model.c:693-755); max_patches < 0: the caller's while (ncomp > 1) loop (:205-212).
        Returns (code, path, comp, ncomp, cost, [(a, b, delta) per patch])."""
        path = np.ascontiguousarray(path, dtype=np.int32).copy()
        comp = np.ascontiguousarray(comp, dtype=np.int32).copy()
        nc, c, done = C.c_int(ncomp), C.c_double(cost), C.c_int(0)
        cap = max(1, ncomp)
        log = (Move * cap)()
        code = _chk(lib().tspcuda_patching(self._h, _ipc(path), _ipc(comp), C.byref(nc), C.byref(c), max_patches, log, cap,
                                           C.byref(done)), allow=allow)
        return code, path, comp, nc.value, c.value, [(log[k].a, log[k].b, log[k].delta) for k in range(done.value)]

    # -- device-resident tour ----------------------------------------------------------------------
    def tour_upload(self, path):
        path = np.ascontiguousarray(path, dtype=np.int32)
        _chk(lib().tspcuda_tour_upload(self._h, _ipc(path)))

    def tour_descend(self, max_moves=-1):
        m, opt = C.c_int64(0), C.c_int(0)
        _chk(lib().tspcuda_tour_descend(self._h, max_moves, C.byref(m), C.byref(opt)))
        return m.value, bool(opt.value)

    def tour_download(self):
        path = np.zeros(self.n, dtype=np.int32)
        cost = C.c_double(0)
        _chk(lib().tspcuda_tour_download(self._h, _ipc(path), C.byref(cost)))
        return path, cost.value

    def tour_movelog(self, cap):
        log = (Move * max(1, cap))()
        cnt = C.c_int64(0)
        _chk(lib().tspcuda_tour_movelog(self._h, log, cap, C.byref(cnt)))
        k = min(cnt.value, cap)
        return cnt.value, [(log[i].a, log[i].b, log[i].delta) for i in range(k)]

    # -- multi-start ---------------------------------------------------------------------------------
    def multistart_two_opt(self, first_start, nstarts, stride=1, time_limit=None):
        """h_greedy_2opt's loop.  Without time_limit: (best_start, best_path, best_cost, nn_costs, final_costs, moves); with one
        (seconds): (code, starts_done, ...the same six...) -- code DEADLINE_EXCEEDED = best over the prefix of starts taken."""
        nn_costs = np.zeros(nstarts, dtype=np.float64)
        final = np.zeros(nstarts, dtype=np.float64)
        moves = np.zeros(nstarts, dtype=np.int64)
        best_path = np.zeros(self.n, dtype=np.int32)
        bs, bc, done = C.c_int(-1), C.c_double(0), C.c_int(0)
        code = _chk(lib().tspcuda_multistart_two_opt(self._h, first_start, nstarts, stride, -1.0 if time_limit is None else time_limit,
                                                     _dpc(nn_costs), _dpc(final), moves.ctypes.data_as(_i64p), C.byref(bs), _ipc(best_path),
                                                     C.byref(bc), C.byref(done)), allow=(DEADLINE_EXCEEDED,))
        if time_limit is None:
            return bs.value, best_path, bc.value, nn_costs, final, moves
        return code, done.value, bs.value, best_path, bc.value, nn_costs, final, moves

    def restarts(self, kind, first_start, nstarts, stride=1, iters=100, seed=0, time_limit=-1.0):
        """Config 5: tabu (RESTART_TABU) / VNS (RESTART_VNS) restarts from start nodes first_start + r*stride.
        Returns (code, restarts_done, best_start, best_path, best_cost)."""
        best_path = np.zeros(self.n, dtype=np.int32)
        bs, bc, done = C.c_int(-1), C.c_double(0), C.c_int(0)
        code = _chk(lib().tspcuda_restarts(self._h, kind, first_start, nstarts, stride, iters, seed, time_limit, C.byref(bs),
                                           _ipc(best_path), C.byref(bc), C.byref(done)), allow=(DEADLINE_EXCEEDED,))
        return code, done.value, bs.value, best_path, bc.value

    # -- measurement ---------------------------------------------------------------------------------
    def pipe_rate(self, kind):
        r = C.c_double(0)
        _chk(lib().tspcuda_pipe_rate(self._h, kind, C.byref(r)))
        return r.value

    def selfcheck_sqrt(self):
        m = C.c_int64(-1)
        _chk(lib().tspcuda_selfcheck_sqrt(self._h, C.byref(m)))
        return m.value

    def timer_start(self):
        _chk(lib().tspcuda_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_double(0)
        _chk(lib().tspcuda_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def flush_l2(self):
        _chk(lib().tspcuda_flush_l2(self._h))


class MatrixInstance:
    """A caller-supplied dense cost matrix resident on one GPU (tspcuda_mx): the `costs`-argument paths of
    h_greedyutil / ref_2opt / ref_2opt_once and, through them, h_Greedy_2opt_mod_costs."""

    def __init__(self, costs, device=0):
        self.costs = np.ascontiguousarray(costs, dtype=np.float64)
        assert self.costs.ndim == 2 and self.costs.shape[0] == self.costs.shape[1]
        self.n = self.costs.shape[0]
        self._h = _vp()
        _chk(lib().tspcuda_mx_create(C.byref(self._h), _dpc(self.costs), self.n, device))

    def close(self):
        if self._h:
            lib().tspcuda_mx_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def stats(self):
        s = Stats()
        _chk(lib().tspcuda_mx_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def update(self, costs):
        """Replace the matrix (same n) in place: no device allocation."""
        costs = np.ascontiguousarray(costs, dtype=np.float64)
        assert costs.shape == (self.n, self.n)
        self.costs = costs
        _chk(lib().tspcuda_mx_update(self._h, _dpc(self.costs)))

    def nn(self, start):
        """h_greedyutil(start, ., costs): (code, path, cost, nodes reached)."""
        path = np.full(self.n, -7, dtype=np.int32)
        cost, ln = C.c_double(0), C.c_int(0)
        code = lib().tspcuda_mx_nn(self._h, start, _ipc(path), C.byref(cost), C.byref(ln))
        return code, path, cost.value, ln.value

    def two_opt(self, path, max_moves=-1, time_limit=-1.0, log_cap=0):
        """ref_2opt(., costs, false): (code, moves, path, cost, log[(a,b,delta)...])."""
        path = np.array(path, dtype=np.int32, copy=True)
        cost, moves = C.c_double(0), C.c_int64(0)
        log = (MoveF * max(1, log_cap))()
        code = _chk(lib().tspcuda_mx_two_opt(self._h, _ipc(path), C.byref(cost), max_moves, time_limit, C.byref(moves),
                                             log if log_cap else None, log_cap), allow=(DEADLINE_EXCEEDED,))
        k = min(moves.value, log_cap)
        return code, moves.value, path, cost.value, [(log[i].a, log[i].b, log[i].delta) for i in range(k)]

    def two_opt_once(self, path, cost):
        """ref_2opt_once(., costs) in place on `path`: (best_delta, new cost)."""
        c, d = C.c_double(cost), C.c_double(0)
        _chk(lib().tspcuda_mx_two_opt_once(self._h, _ipc(path), C.byref(c), C.byref(d)))
        return d.value, c.value

    def mod_costs(self):
        """h_Greedy_2opt_mod_costs without a time limit (heuristics.c:118-149): the loop keeps nothing between
        starts, so its result is NN(n-1) + 2-opt.  Returns (path, cost)."""
        code, p, c, ln = self.nn(self.n - 1)
        code, moves, p, c, _ = self.two_opt(p)
        return p, c


class MultiGpuInstance:
    """tspcuda_mgpu: one context per GPU of this box inside ONE process (what the drop-in CLI uses with
    TSPCUDA_GPUS > 1); multi-start work sharded block-cyclically, NCCL all-reduce(min) + broadcast of the best tour."""

    def __init__(self, xy, weight_type=EUC_REF, ndev=0, devices=None):
        self.xy = np.ascontiguousarray(xy, dtype=np.float64)
        self.n = self.xy.shape[0]
        self._h = _vp()
        dev = None if devices is None else np.ascontiguousarray(devices, dtype=np.int32)
        _chk(lib().tspcuda_mgpu_create(C.byref(self._h), _dpc(self.xy), self.n, weight_type, ndev,
                                       None if dev is None else _ipc(dev)))
        self.ndev = lib().tspcuda_mgpu_ndev(self._h)

    def close(self):
        if self._h:
            lib().tspcuda_mgpu_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def multistart(self, two_opt=True, first_start=0, nstarts=None, stride=1, time_limit=None):
        """(best_start, best_path, best_cost, compute_ms, reduce_ms) over starts first_start + k*stride; with a time_limit
        (seconds) the tuple is preceded by (code, starts_done)."""
        nstarts = self.n if nstarts is None else nstarts
        path = np.zeros(self.n, dtype=np.int32)
        bs, bc, cm, rm, done = C.c_int(-1), C.c_double(0), C.c_double(0), C.c_double(0), C.c_int(0)
        code = _chk(lib().tspcuda_mgpu_multistart(self._h, 1 if two_opt else 0, first_start, nstarts, stride,
                                                  -1.0 if time_limit is None else time_limit, C.byref(bs), _ipc(path), C.byref(bc),
                                                  C.byref(done), C.byref(cm), C.byref(rm)), allow=(DEADLINE_EXCEEDED,))
        if time_limit is None:
            return bs.value, path, bc.value, cm.value, rm.value
        return code, done.value, bs.value, path, bc.value, cm.value, rm.value

    def restarts(self, kind, first_start, nstarts, stride=1, iters=100, seed=0, time_limit=-1.0):
        """Config 5 restarts sharded over the GPUs: (code, restarts_done, best_start, best_path, best_cost, compute_ms, reduce_ms)."""
        path = np.zeros(self.n, dtype=np.int32)
        bs, bc, cm, rm, done = C.c_int(-1), C.c_double(0), C.c_double(0), C.c_double(0), C.c_int(0)
        code = _chk(lib().tspcuda_mgpu_restarts(self._h, kind, first_start, nstarts, stride, iters, seed, time_limit, C.byref(bs),
                                                _ipc(path), C.byref(bc), C.byref(done), C.byref(cm), C.byref(rm)), allow=(DEADLINE_EXCEEDED,))
        return code, done.value, bs.value, path, bc.value, cm.value, rm.value
