/*
 * tspcuda.h -- C ABI of libtspcuda.so: the B200-native (sm_100a) heuristic core of the C TSP solver
 * enricobolzonello/TravellingSalesmanOptimization.
 *
 * Plain C, plain pointers and sizes; no CUDA, torch or reference types in any signature.  Every
 * entry point returns one of the reference's ERROR_CODE values (src/utils/errors.h:33-51), so the
 * interposer (travellingsalesmanoptimization_b200/csrc/interpose.c) can hand them straight back to the unchanged driver.
 * There is NO CPU fallback anywhere behind this header: without a CUDA device every call that
 * needs one returns TSPCUDA_UNAVAILABLE and tspcuda_last_error() says why.
 *
 * Reference interface each entry point replaces (file:line under the reference's src/):
 *   tspcuda_cost_matrix           tsp_compute_costs            tsp.c:608-636
 *   tspcuda_nn                    h_greedyutil                 algorithms/heuristics.c:216-288
 *   tspcuda_nn_multi              h_Greedy_iterative loop      algorithms/heuristics.c:34-72
 *   tspcuda_two_opt               ref_2opt (sans incumbent)    algorithms/refinment.c:3-37
 *   tspcuda_two_opt_once          ref_2opt_once                algorithms/refinment.c:39-93
 *                                 (+ ref_reverse_path          algorithms/refinment.c:95-114)
 *   tspcuda_tabu_move             tabu_best_move               algorithms/metaheuristic.c:188-245
 *   tspcuda_tabu_search           mh_TabuSearch inner loop     algorithms/metaheuristic.c:128-164
 *   tspcuda_multistart_two_opt    h_greedy_2opt loop           algorithms/heuristics.c:74-116
 *   tspcuda_extra_mileage         h_ExtraMileage / h_extramileage_util   algorithms/heuristics.c:156-210, :290-367
 *   tspcuda_mx_nn / _two_opt      h_greedyutil / ref_2opt on a caller's matrix (h_Greedy_2opt_mod_costs, heuristics.c:118-149)
 *   tspcuda_vns_search            mh_VNS loop (ref_2opt + vns_kick)              algorithms/metaheuristic.c:277-321, :344-409
 *   tspcuda_patching              cx_patching (+ its while (ncomp > 1) loop)   algorithms/cplex_model.c:693-755, :205-212
 *   tspcuda_restarts / _mgpu_*    no counterpart: independent tabu / VNS runs from different start nodes (BASELINE.json configs[4])
 *
 * Tours cross the boundary exactly as the reference stores them: a successor array int path[n],
 * path[i] = node visited after node i (src/utils/utils.h:42-47).  Costs are doubles holding
 * integers (the reference's rounded edge costs), bit-identical to the reference's.
 */
#ifndef TSPCUDA_H_
#define TSPCUDA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ERROR_CODE values of the reference (src/utils/errors.h:33-51) used by this library. */
enum {
    TSPCUDA_OK = 0,                 /* T_OK */
    TSPCUDA_CANCELLED = 1,          /* CANCELLED (not an error, errors.c:31-37) */
    TSPCUDA_INVALID_ARGUMENT = 3,   /* bad pointer / not a single-cycle successor array */
    TSPCUDA_DEADLINE_EXCEEDED = 4,  /* time limit hit between scans (refinment.c:17-24); not an error */
    TSPCUDA_RESOURCE_EXHAUSTED = 8, /* device memory */
    TSPCUDA_FAILED_PRECONDITION = 9,/* patching: no admissible pair (see tspcuda_patching) */
    TSPCUDA_OUT_OF_RANGE = 11,      /* n or coordinate extent beyond what the packed key supports */
    TSPCUDA_UNIMPLEMENTED = 12,
    TSPCUDA_INTERNAL = 13,          /* CUDA failure */
    TSPCUDA_UNAVAILABLE = 14        /* no CUDA device, or start node out of range (heuristics.c:223-226) */
};

/* Edge-weight rule.  Only EUC_REF exists in the reference (tsp.c:629: float sqrt, round half up);
 * the others follow TSPLIB95 and have no reference behaviour ("parity unpinned"). */
enum { TSPCUDA_EUC_REF = 0, TSPCUDA_CEIL_2D = 1, TSPCUDA_ATT = 2, TSPCUDA_GEO = 3 };

#define TSPCUDA_MAX_NODES 131072      /* 17-bit node ids in the packed (delta,a,b) argmin key */

typedef struct tspcuda_ctx tspcuda_ctx;

typedef struct {
    int32_t a, b;        /* node ids, a < b: the reference's best_swap (refinment.c:65-66) */
    int32_t delta;       /* best_delta (integer-valued) */
    int32_t reserved;
} tspcuda_move;

typedef struct {
    int64_t scans;            /* 2-opt neighbourhood scans run (each n(n-1)/2 pair evaluations) */
    int64_t moves;            /* moves applied */
    int64_t pair_evals;       /* scans * n(n-1)/2 */
    int64_t kernel_launches;  /* kernels of this library launched since the context was created */
    double scan_kernel_ms;    /* CUDA-event time of the descent kernels (scan + apply), accumulated */
    double last_kernel_ms;    /* CUDA-event time of the most recent timed kernel */
    int64_t h2d_bytes;        /* bytes copied host->device by this context */
    int64_t d2h_bytes;        /* bytes copied device->host by this context */
    int64_t pruned_scans;     /* scans (of `scans`) answered by the exact candidate-pruned kernel; they
                                 add nothing to pair_evals / scan_kernel_ms */
    int64_t candidate_evals;  /* candidate nodes examined by those scans */
    double pruned_kernel_ms;  /* CUDA-event time of the pruned descent kernels, accumulated */
} tspcuda_stats;

/* ---- library -------------------------------------------------------------------------------- */
int tspcuda_device_count(void);                 /* 0 when no usable CUDA device */
const char *tspcuda_last_error(void);           /* thread-local description of the last failure */
const char *tspcuda_version(void);

/* ---- instance context ----------------------------------------------------------------------- */
/* A context owns one CUDA stream, the device copy of the instance and all scratch.  Calls on the SAME
 * context must not overlap (the reference's heuristic drivers are single-threaded); different contexts
 * (tspcuda_ctx or tspcuda_mx) may be used from different threads concurrently. */
/* xy: n points, AoS {x,y} doubles exactly as the reference's `point` array (utils.h:37-40).
 * The coordinates are copied to the device; the caller keeps ownership of xy. */
int tspcuda_create(tspcuda_ctx **out, const double *xy, int n, int weight_type, int device);
void tspcuda_destroy(tspcuda_ctx *ctx);
int tspcuda_nnodes(const tspcuda_ctx *ctx);
int tspcuda_get_stats(tspcuda_ctx *ctx, tspcuda_stats *out);
int tspcuda_reset_stats(tspcuda_ctx *ctx);

/* ---- K1: cost matrix (tsp.c:608-636) -------------------------------------------------------- */
/* Fills host_out[n*n] (row-major doubles, diagonal -1.0).  The matrix is built in HBM with
 * 128-bit stores and copied back; size_t indexing (the reference's int n*n overflows at 46341). */
int tspcuda_cost_matrix(tspcuda_ctx *ctx, double *host_out);
/* Device-only build (kept in HBM, nothing copied): kernel time in *ms_out.  For measurement. */
int tspcuda_cost_matrix_device(tspcuda_ctx *ctx, double *ms_out);
/* Device-side checksum (sum of all entries, exact in int64) of the HBM-resident matrix. */
int tspcuda_cost_matrix_checksum(tspcuda_ctx *ctx, int64_t *sum_out);

/* ---- nearest neighbour (heuristics.c:216-288) ----------------------------------------------- */
int tspcuda_nn(tspcuda_ctx *ctx, int start, int *path_out, double *cost_out);
/* Starts first_start + k*stride, k in [0,nstarts): per-start costs, and the best tour with the
 * reference's tie rule (strict '<' in ascending start order => lowest start wins, heuristics.c:58).
 * time_limit_s < 0: none.  Otherwise seconds from the call: the clock is looked at before starts are taken, as the
 * reference does before every start (heuristics.c:42-50); once it has passed no further start is taken and the call
 * returns TSPCUDA_DEADLINE_EXCEEDED with the best of the PREFIX of starts it got to (*starts_done_out of them, in
 * ascending order; *best_start_out = -1 if none).  costs_out / best_path_out / starts_done_out may be NULL. */
int tspcuda_nn_multi(tspcuda_ctx *ctx, int first_start, int nstarts, int stride, double time_limit_s, double *costs_out,
                     int *best_start_out, int *best_path_out, double *best_cost_out, int *starts_done_out);

/* ---- 2-opt (refinment.c) --------------------------------------------------------------------- */
/* Host tour in, host tour out.  Recomputes the cost (refinment.c:6-9), then repeats
 * best-improvement scan + segment reversal while best_delta < -1e-7 (:15-27).
 * max_moves < 0: until the local optimum.  time_limit_s < 0: none; otherwise checked between
 * batches of scans and TSPCUDA_DEADLINE_EXCEEDED is returned with the tour reached so far.
 * log (may be NULL): the first log_cap moves (a,b,delta). */
int tspcuda_two_opt(tspcuda_ctx *ctx, int *path, double *cost_out, int64_t max_moves, double time_limit_s,
                    int64_t *moves_out, tspcuda_move *log, int64_t log_cap);
/* One scan (+ move if best_delta < -1e-7): *cost_inout += best_delta; returns best_delta in
 * *best_delta_out (0 at a local optimum), the move in *move_out (a = -1 if none; may be NULL). */
int tspcuda_two_opt_once(tspcuda_ctx *ctx, int *path, double *cost_inout, double *best_delta_out,
                         tspcuda_move *move_out);
/* tabu_best_move (metaheuristic.c:188-245): best admissible move even if worsening; applies it,
 * *cost_inout += delta, stamps tabu_list[a,b,succ_a,succ_b] = iter.  move_out->a = -1 if every
 * pair was tabu.  tabu_list: int[n], -1 = never (metaheuristic.c:68-70). */
int tspcuda_tabu_move(tspcuda_ctx *ctx, int *path, double *cost_inout, int *tabu_list, int tenure, int iter,
                      tspcuda_move *move_out);

/* n_iters iterations of the tabu loop of mh_TabuSearch (metaheuristic.c:128-164) in one device-resident
 * run: iteration k uses tenure tenures[k] (computed ahead by the caller with the reference's own
 * policy functions, metaheuristic.c:7-59, which do not depend on the search state) and iteration
 * number first_iter + k, exactly as n_iters calls of tspcuda_tabu_move would.  path / *cost_inout /
 * tabu_list are updated in place; cost_trace_out[k] = tour cost after iteration k (the
 * results/TabuResults.dat column).  The incumbent rule of tsp_update_best_solution (strict '<',
 * tsp.c:671) is applied on the device against incumbent_cost: if some iteration improves on it,
 * *best_iter_out is the last such iteration (-1 otherwise) and best_path_out / *best_cost_out hold
 * that tour.  Any output pointer except path, cost_inout and tabu_list may be NULL.
 * incumbent_cost: any value >= 9e18, +infinity and the reference's own "no incumbent yet" DBL_MAX (tsp.c:37) mean that
 * every cost is an improvement; NaN (and a non-finite *cost_inout) is TSPCUDA_INVALID_ARGUMENT.
 * Scans: the exact candidate-pruned scan with a bound (K5c, n >= 4096) while it finds an admissible pair, the full scan
 * (K5b) otherwise and always under TSPCUDA_SCAN_FULL -- identical traces either way. */
int tspcuda_tabu_search(tspcuda_ctx *ctx, int *path, double *cost_inout, int *tabu_list, const int *tenures,
                        int first_iter, int n_iters, double incumbent_cost, double *cost_trace_out,
                        int *best_path_out, double *best_cost_out, int *best_iter_out);

/* n_iters iterations of the loop of mh_VNS (metaheuristic.c:277-321) with the tour resident on the device:
 * iteration i = ref_2opt to the local optimum (:288), incumbent test on a strictly smaller cost (:296-300),
 * cost_trace_out[i] = that cost (the results/VNSResults.dat column, :303), then rand() % 9 - 2 kicks (:306-316;
 * vns_kick :344-409 = three random positions with the reference's rejection test, segment exchange of tabu_make_move
 * case 7).  The kick runs on the host on the tour order and draws from the PROCESS's rand() stream, in the reference's
 * order -- the caller seeds it (or does not: the reference never seeds it on file instances); per iteration the tour
 * order travels device -> host and back (2 x 4n bytes), nothing else.  time_limit_s < 0: none; otherwise seconds from
 * the call, tested at the top of every iteration and inside the descents (TSPCUDA_DEADLINE_EXCEEDED).
 * path: in = start tour, out = the tour after the last kicks (what the reference's `solution` holds at the end);
 * *cost_inout: cost after the last descent.  *best_cost_inout: in = the cost to beat (mh_VNS: the start cost, :270),
 * out = the best cost seen; best_path_out / *best_iter_out: that tour and its iteration (-1: no iteration improved
 * on the input; best_path_out is then untouched).  *iters_done_out: iterations completed. */
int tspcuda_vns_search(tspcuda_ctx *ctx, int *path, double *cost_inout, int n_iters, double time_limit_s,
                       double *cost_trace_out, int *best_path_out, double *best_cost_inout, int *best_iter_out,
                       int *iters_done_out);

/* How 2-opt descents scan the neighbourhood.  FULL evaluates all n(n-1)/2 pairs per scan, as the
 * reference does (refinment.c:49-69).  PRUNED evaluates only the pairs that can improve the tour
 * (found by fixed-radius searches over a grid of the nodes, csrc/pruned.cuh) and provably returns
 * the same move, ties included.  AUTO (default) measures both and uses the faster one.  The per-call tabu
 * move (tspcuda_tabu_move: best admissible move, even worsening) always uses FULL; the device-resident tabu loop
 * uses a bounded pruned scan where it can (see tspcuda_tabu_search).  Environment: TSPCUDA_SCAN=full|pruned. */
enum { TSPCUDA_SCAN_AUTO = 0, TSPCUDA_SCAN_FULL = 1, TSPCUDA_SCAN_PRUNED = 2 };
int tspcuda_set_scan_mode(tspcuda_ctx *ctx, int mode);

/* Device-resident variants (inputs already in HBM; used for kernel-only measurement and by the
 * multi-start drivers): upload once, descend in place, download when wanted. */
int tspcuda_tour_upload(tspcuda_ctx *ctx, const int *path);
int tspcuda_tour_descend(tspcuda_ctx *ctx, int64_t max_moves, int64_t *moves_out, int *at_local_optimum);
int tspcuda_tour_download(tspcuda_ctx *ctx, int *path_out, double *cost_out);
int tspcuda_tour_movelog(tspcuda_ctx *ctx, tspcuda_move *log, int64_t cap, int64_t *count_out);

/* ---- multi-start NN + 2-opt (heuristics.c:74-116) ------------------------------------------- */
/* Starts first_start + k*stride.  Per-start NN costs, final costs and move counts (any may be
 * NULL); best tour under the reference's incumbent rule (strict '<', tsp.c:671, ascending start
 * order => lowest start wins ties).  The device scratch of a batch (28 n bytes per start of a batch of
 * at most 16 x #SMs starts, 192 n bytes per concurrent descent, at most #SMs) is kept in the
 * context for later calls and released by tspcuda_destroy.
 * time_limit_s < 0: none.  Otherwise seconds from the call, with the reference's semantics (heuristics.c:83-90,
 * refinment.c:17-24): the limit is looked at before a start is taken and before every scan of a descent; once it has
 * passed no further start is taken (starts are handed out in ascending order, so the ones taken are a prefix:
 * *starts_done_out of them), descents in flight stop where they are and their tours still compete for the incumbent
 * (as ref_2opt's update_incumbent does after DEADLINE_EXCEEDED), and the call returns TSPCUDA_DEADLINE_EXCEEDED.
 * Entries of final_costs_out / moves_out for starts never taken are -1; *best_start_out = -1 if none was. */
int tspcuda_multistart_two_opt(tspcuda_ctx *ctx, int first_start, int nstarts, int stride, double time_limit_s,
                               double *nn_costs_out, double *final_costs_out, int64_t *moves_out, int *best_start_out,
                               int *best_path_out, double *best_cost_out, int *starts_done_out);

/* ---- tabu / VNS restarts (BASELINE.json configs[4]; a framework addition: the reference runs one trajectory) ---- */
/* Restart r = start node first_start + r*stride, an independent unit:
 *   TSPCUDA_RESTART_TABU  NN(start) -> 2-opt local optimum (ref_2opt) -> `iters` iterations of the loop of mh_TabuSearch
 *                         (metaheuristic.c:128-164) with the default linear tenure schedule (tabu_init :65-84,
 *                         tabu_linear_policy :40-59); result = the best tour seen, local optimum included
 *   TSPCUDA_RESTART_VNS   NN(start) -> `iters` iterations of the loop of mh_VNS (metaheuristic.c:277-321); the kicks
 *                         draw from a generator private to the restart that yields the sequence srand(seed + start)
 *                         gives the process stream (glibc random_r), so restarts are reproducible whatever thread or
 *                         GPU runs them; result = the best local optimum seen
 * Best over the restarts under the incumbent rule (strict '<' in ascending start order).  time_limit_s as above:
 * TSPCUDA_DEADLINE_EXCEEDED with the best of the restarts it got to (*restarts_done_out, the last possibly cut short). */
enum { TSPCUDA_RESTART_TABU = 0, TSPCUDA_RESTART_VNS = 1 };
int tspcuda_restarts(tspcuda_ctx *ctx, int kind, int first_start, int nstarts, int stride, int iters, unsigned seed,
                     double time_limit_s, int *best_start_out, int *best_path_out, double *best_cost_out, int *restarts_done_out);

/* ---- Extra Mileage (heuristics.c:156-210 h_ExtraMileage, :290-367 h_extramileage_util) ---------- */
/* node_a < 0: the EM_MAX initial edge (first pair i<j with the strictly largest cost, :169-178), else the
 * given edge (EM_RANDOM draws it on the host with rand(), :180-183).  Then n-2 insertions with the
 * reference's tie rule (first strict minimum over unvisited i ascending, tour edges in edges[] order).
 * *cost_out = 2 c(A,B) + the mileage of every insertion (:193, :360); path_out = successor array. */
int tspcuda_extra_mileage(tspcuda_ctx *ctx, int node_a, int node_b, int *path_out, double *cost_out,
                          int *node_a_out, int *node_b_out);

/* ---- patching heuristic of the Benders loop (cplex_model.c:693-755 cx_patching) ------------------------- */
/* path: a successor array made of *ncomp_inout disjoint cycles, comp[v] = label of the cycle of v (one label
 * per cycle, as cx_build_sol leaves them; anything else is INVALID_ARGUMENT).  One patch = the first strict
 * minimum, over ordered node pairs (a,b) of different components in a-major order, of
 * delta = [c(a,succ b) + c(b,succ a)] - [c(a,succ a) + c(b,succ b)] among the pairs with delta > 0 (:699-722);
 * then path[a] = succ b, path[b] = succ a, *cost_inout += delta, the cycle of b takes the label of a and
 * *ncomp_inout decreases (:727-754).  max_patches < 0: until one component is left -- the caller's
 * `while (solution.ncomp > 1) cx_patching(&solution)` (:205-212) in one device-resident run; 1 = one call of
 * cx_patching.  log (may be NULL): (a, b, delta) of the first log_cap patches; *patches_out (may be NULL):
 * patches applied.  If some scan finds no pair with delta > 0 the state reached so far is returned with
 * TSPCUDA_FAILED_PRECONDITION (the reference has undefined behaviour there: it indexes path[-1], :730-735). */
int tspcuda_patching(tspcuda_ctx *ctx, int *path, int *comp, int *ncomp_inout, double *cost_inout, int max_patches,
                     tspcuda_move *log, int64_t log_cap, int *patches_out);

/* ---- the `costs`-argument paths: a caller-supplied dense matrix ------------------------------- */
/* h_greedyutil / ref_2opt / ref_2opt_once called with a matrix other than the instance's own
 * (heuristics.c:216, refinment.c:3,39), and through them h_Greedy_2opt_mod_costs (heuristics.c:118-149,
 * the x*-weighted matrices of the CPLEX relaxation callback, cplex_model.c:1187-1217).  The matrix
 * (row-major n*n doubles, any values; -1.0 = NOT_CONNECTED for the NN step, heuristics.c:257) is
 * copied to the device once per context.  Deltas are FP64 in the reference's expression order
 * (refinment.c:60-62); tours and costs are bit-identical to the reference's. */
typedef struct tspcuda_mx tspcuda_mx;
typedef struct {
    int32_t a, b;
    double delta;
} tspcuda_move_f;
int tspcuda_mx_create(tspcuda_mx **out, const double *costs, int n, int device);
void tspcuda_mx_destroy(tspcuda_mx *mx);
/* Replaces the matrix of an existing context (same n) without re-allocating anything: for callers that receive a
 * fresh matrix per call (the relaxation callback, cplex_model.c:1187-1217, thousands of calls per run). */
int tspcuda_mx_update(tspcuda_mx *mx, const double *costs);
int tspcuda_mx_get_stats(tspcuda_mx *mx, tspcuda_stats *out);
/* *len_out (may be NULL): nodes reached; < n when NOT_CONNECTED entries close the path early. */
int tspcuda_mx_nn(tspcuda_mx *mx, int start, int *path_out, double *cost_out, int *len_out);
int tspcuda_mx_two_opt(tspcuda_mx *mx, int *path, double *cost_out, int64_t max_moves, double time_limit_s,
                       int64_t *moves_out, tspcuda_move_f *log, int64_t log_cap);
int tspcuda_mx_two_opt_once(tspcuda_mx *mx, int *path, double *cost_inout, double *best_delta_out);

/* ---- multi-start work over the GPUs of one box, from inside the library ------------------ */
/* One worker thread and one context per GPU; starts are sharded block-cyclically (GPU g of G takes the caller's
 * starts g, g+G, ...), no data-path collective; the incumbent is merged by ONE ncclAllReduce(min) of the 64-bit key
 * cost << 20 | start node (= the reference's strict-'<' rule in ascending start order, tsp.c:671, heuristics.c:58/82:
 * the lowest start wins cost ties) and ONE ncclBroadcast of the winner's successor array.  NCCL is loaded at run time
 * (libnccl.so.2); without it, or with fewer visible GPUs than asked for, the calls return TSPCUDA_UNAVAILABLE.
 * ndev <= 0: every visible GPU; devices == NULL: 0..ndev-1.  With ndev == 1 no NCCL is needed or loaded.
 * This is what the interposer uses for h_Greedy_iterative / h_greedy_2opt when TSPCUDA_GPUS > 1. */
typedef struct tspcuda_mgpu tspcuda_mgpu;
int tspcuda_mgpu_create(tspcuda_mgpu **out, const double *xy, int n, int weight_type, int ndev, const int *devices);
void tspcuda_mgpu_destroy(tspcuda_mgpu *m);
int tspcuda_mgpu_ndev(const tspcuda_mgpu *m);
/* two_opt = 0: all-start nearest neighbour (h_Greedy_iterative, heuristics.c:34-72); 1: NN + 2-opt per start
 * (h_greedy_2opt, :74-116).  Starts first_start + k*stride, k in [0,nstarts).  Same results as the single-GPU calls
 * tspcuda_nn_multi / tspcuda_multistart_two_opt.  *compute_ms_out / *reduce_ms_out (may be NULL): wall time of the
 * slowest shard and of the incumbent reduction (max over GPUs).  time_limit_s: every GPU works against the same limit
 * on its own shard (see the single-GPU calls); TSPCUDA_DEADLINE_EXCEEDED then comes with the best over the prefixes of
 * the shards, *starts_done_out = starts taken over all GPUs, *best_start_out = -1 if none was. */
int tspcuda_mgpu_multistart(tspcuda_mgpu *m, int two_opt, int first_start, int nstarts, int stride, double time_limit_s,
                            int *best_start_out, int *best_path_out, double *best_cost_out, int *starts_done_out,
                            double *compute_ms_out, double *reduce_ms_out);
/* tspcuda_restarts sharded the same way (restart r on GPU r mod G), same reduction. */
int tspcuda_mgpu_restarts(tspcuda_mgpu *m, int kind, int first_start, int nstarts, int stride, int iters, unsigned seed,
                          double time_limit_s, int *best_start_out, int *best_path_out, double *best_cost_out,
                          int *restarts_done_out, double *compute_ms_out, double *reduce_ms_out);

/* ---- measurement helpers -------------------------------------------------------------------- */
/* CUDA-event bracket on the stream every kernel of this context is launched on (torch.cuda.Event
 * only sees torch's own stream): start records an event, stop records another, synchronises and
 * returns the elapsed device time in *ms_out. */
int tspcuda_timer_start(tspcuda_ctx *ctx);
int tspcuda_timer_stop(tspcuda_ctx *ctx, double *ms_out);
/* Overwrites a scratch buffer larger than the 126 MB L2 (between timed iterations). */
int tspcuda_flush_l2(tspcuda_ctx *ctx);
/* Register-only pipe-rate microbenchmarks run on the context's device; result in operations/s
 * (thread-level, not warp-level).  kind: 0 DADD, 1 DMUL, 2 DFMA, 3 F2F.F32.F64, 4 MUFU.RSQ,
 * 5 FFMA, 6 the full EUC_REF edge-cost sequence (general coordinates), 7 the same for integer coordinates
 * (edge costs/s: the arithmetic ceiling of a scan that needs one edge cost per pair), 8 F2F.F32.F64 alone,
 * 9 the integer-coordinate sequence two costs at a time with packed FP32x2 instructions (what the scan runs);
 * 10 the same two costs from FP32 instructions only (error-free transformation, no FP64, no F2F: measured slower, unused),
 * 101 its mismatch count against kind 9's sequence over 1.5e9 integer coordinate triples (must be 0);
 * 100 (after tspcuda_cost_matrix_device) bytes/s of the driver's own fill (cudaMemsetAsync) of the HBM-resident matrix
 * buffer: what a pure writer reaches on this device, next to which the matrix build is read. */
int tspcuda_pipe_rate(tspcuda_ctx *ctx, int kind, double *ops_per_s_out);
/* Exhaustive self-check of the in-kernel float sqrt against IEEE sqrt.rn.f32 over every float in
 * [2^-60, 2^64): number of mismatches in *mismatches_out (must be 0). */
int tspcuda_selfcheck_sqrt(tspcuda_ctx *ctx, int64_t *mismatches_out);

#ifdef __cplusplus
}
#endif
#endif /* TSPCUDA_H_ */
