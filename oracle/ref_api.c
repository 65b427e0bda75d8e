/*
 * oracle/ref_api.c -- TEST INFRASTRUCTURE.  Never linked into, imported by or called from the product
 * (libtspcuda.so / travellingsalesmanoptimization_b200); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load the library this file is built into.
 *
 * A flat, ctypes-friendly API over the UNMODIFIED reference objects.  This file is compiled against
 * the reference's own headers (-I/root/reference/src) and linked with oracle/_ref/libtspref.so, which
 * oracle/Makefile builds from the reference sources where they lie (they are never copied into this repo).
 * It contains no algorithm of its own: every function sets the reference's globals
 * (tsp_inst / tsp_env, src/tsp.c:3-4) and calls the reference's own helpers:
 *   tsp_compute_costs   src/tsp.c:608
 *   h_greedyutil        src/algorithms/heuristics.c:216
 *   ref_2opt_once       src/algorithms/refinment.c:39
 *   ref_2opt            src/algorithms/refinment.c:3
 *   tabu_best_move      src/algorithms/metaheuristic.c:188
 *   h_Greedy_iterative / h_greedy_2opt / mh_TabuSearch / mh_VNS  (whole-algorithm entry points)
 *   cx_patching         src/algorithms/cplex_model.c:693  (plain C; that file is compiled unchanged against
 *                       oracle/stubs/cx/cplex.h, every CPX* / CCcut_* entry point it mentions is a generated
 *                       stub that aborts -- cx_patching calls none of them)
 */
#define _GNU_SOURCE
#include <signal.h>
#include <stdlib.h>
#include <string.h>

#include "algorithms/metaheuristic.h"   /* -> heuristics.h -> refinment.h -> tsp.h */

static int g_inited = 0;

static void refapi_init_once(void) {
    if (!g_inited) {
        tsp_init();                       /* src/tsp.c:6 */
        err_setverbosity(QUIET);
        signal(SIGPIPE, SIG_IGN);         /* gnuplot may be absent: src/utils/plot.c:12 */
        tsp_inst.points = NULL;
        tsp_inst.costs = NULL;
        tsp_inst.best_solution.path = NULL;
        tsp_inst.best_solution.comp = NULL;
        tsp_inst.threads_seeds = NULL;
        tsp_env.inputfile = NULL;
        g_inited = 1;
    }
}

static void refapi_drop_instance(void) {
    free(tsp_inst.points); tsp_inst.points = NULL;
    free(tsp_inst.costs); tsp_inst.costs = NULL;
    free(tsp_inst.best_solution.path); tsp_inst.best_solution.path = NULL;
}

/* Install n points (AoS x,y doubles) as the current instance and build the cost matrix with the
 * reference's tsp_compute_costs().  Returns its ERROR_CODE, or -1 if n*n overflows int (tsp.c:614). */
int refapi_set_points(const double *xy, int n) {
    refapi_init_once();
    refapi_drop_instance();
    if (n <= 0 || n > 46340) return -1;
    tsp_inst.nnodes = n;
    tsp_inst.points = (point *)calloc((size_t)n, sizeof(point));
    for (int i = 0; i < n; i++) { tsp_inst.points[i].x = xy[2 * i]; tsp_inst.points[i].y = xy[2 * i + 1]; }
    tsp_inst.best_solution.path = (int *)calloc((size_t)n, sizeof(int));   /* as src/main.c:6 */
    tsp_inst.best_solution.cost = __DBL_MAX__;
    tsp_inst.starting_node = 0;
    tsp_env.timelimit = -1;
    utils_startclock(&tsp_inst.c);
    return (int)tsp_compute_costs();
}

/* Random instance exactly as the driver's "-n N -seed S" does (src/tsp.c:468-481). */
int refapi_random_points(int n, int seed) {
    refapi_init_once();
    refapi_drop_instance();
    if (n <= 0 || n > 46340) return -1;
    tsp_inst.nnodes = n;
    tsp_env.seed = seed;
    tsp_inst.best_solution.cost = __DBL_MAX__;
    tsp_env.timelimit = -1;
    utils_startclock(&tsp_inst.c);
    int e = (int)tsp_generate_randompoints();
    tsp_inst.best_solution.path = (int *)calloc((size_t)n, sizeof(int));
    return e;
}

int refapi_nnodes(void) { return tsp_inst.nnodes; }
void refapi_get_points(double *xy) {
    for (int i = 0; i < tsp_inst.nnodes; i++) { xy[2 * i] = tsp_inst.points[i].x; xy[2 * i + 1] = tsp_inst.points[i].y; }
}
const double *refapi_costs(void) { return tsp_inst.costs; }
void refapi_set_timelimit(double t) { tsp_env.timelimit = t; utils_startclock(&tsp_inst.c); }
void refapi_reset_incumbent(void) { tsp_inst.best_solution.cost = __DBL_MAX__; tsp_inst.starting_node = 0; }
void refapi_srand(unsigned s) { srand(s); }

int refapi_nn(int start, int *path, double *cost) {
    tsp_solution s; s.path = path; s.cost = __DBL_MAX__; s.ncomp = 0; s.comp = NULL;
    int e = (int)h_greedyutil(start, &s, tsp_inst.costs);
    *cost = s.cost;
    return e;
}

double refapi_two_opt_once(int *path, double *cost) {
    tsp_solution s; s.path = path; s.cost = *cost; s.ncomp = 0; s.comp = NULL;
    double d = ref_2opt_once(&s, tsp_inst.costs);
    *cost = s.cost;
    return d;
}

/* Move capture: libtspref.so is built -fPIC with gcc's default semantic interposition, so
 * ref_2opt_once / tabu_best_move reach ref_reverse_path through the PLT.
 * This library precedes libtspref.so in the lookup scope, so the definition below is the one they
 * call; it records (a,b) and forwards to the reference's own ref_reverse_path (refinment.c:95). */
#include <dlfcn.h>
static int g_last_a = -1, g_last_b = -1;
static long g_reverse_calls = 0;
typedef void (*reverse_fn)(int, int, int, int, int *, int *);
void ref_reverse_path(int a, int succ_a, int b, int succ_b, int *prev, int *solution_path) {
    static reverse_fn real = NULL;
    if (!real) real = (reverse_fn)dlsym(RTLD_NEXT, "ref_reverse_path");
    g_last_a = a; g_last_b = b; g_reverse_calls++;
    real(a, succ_a, b, succ_b, prev, solution_path);
}
long refapi_reverse_calls(void) { return g_reverse_calls; }
void refapi_last_move(int *a, int *b) { *a = g_last_a; *b = g_last_b; }

/* ref_2opt's loop (refinment.c:6-27) restated here ONLY to log each move and to allow a move cap;
 * the scan and the reversal are the reference's own ref_2opt_once.  max_moves < 0 = until the
 * local optimum.  Returns the number of moves applied. */
int refapi_two_opt(int *path, double *cost, int max_moves, int *log_a, int *log_b, double *log_delta) {
    int n = tsp_inst.nnodes;
    tsp_solution s; s.path = path; s.cost = 0; s.ncomp = 0; s.comp = NULL;
    for (int i = 0; i < n; i++) s.cost += tsp_inst.costs[i * n + path[i]];      /* refinment.c:6-9 */
    int moves = 0;
    while (max_moves < 0 || moves < max_moves) {
        double d = ref_2opt_once(&s, tsp_inst.costs);
        if (!(d < EPSILON)) break;                                               /* refinment.c:27 */
        if (log_a) { log_a[moves] = g_last_a; log_b[moves] = g_last_b; }
        if (log_delta) log_delta[moves] = d;
        moves++;
    }
    *cost = s.cost;
    return moves;
}

/* The real ref_2opt (refinment.c:3), incumbent update included. */
int refapi_ref_2opt(int *path, double *cost, int update_incumbent) {
    tsp_solution s; s.path = path; s.cost = *cost; s.ncomp = 0; s.comp = NULL;
    int e = (int)ref_2opt(&s, tsp_inst.costs, update_incumbent != 0);
    *cost = s.cost;
    return e;
}

/* The same helpers on a caller-supplied matrix (the `costs` argument of heuristics.c:216, refinment.c:3
 * and heuristics.c:118): nnodes must have been set through refapi_set_points / refapi_random_points. */
int refapi_nn_costs(int start, double *costs, int *path, double *cost) {
    tsp_solution s; s.path = path; s.cost = __DBL_MAX__; s.ncomp = 0; s.comp = NULL;
    int e = (int)h_greedyutil(start, &s, costs);
    *cost = s.cost;
    return e;
}
int refapi_ref_2opt_costs(double *costs, int *path, double *cost) {
    tsp_solution s; s.path = path; s.cost = *cost; s.ncomp = 0; s.comp = NULL;
    int e = (int)ref_2opt(&s, costs, 0);
    *cost = s.cost;
    return e;
}
int refapi_mod_costs(double *costs, int *path, double *cost) {
    tsp_solution s; s.path = path; s.cost = __DBL_MAX__; s.ncomp = 0; s.comp = NULL;
    int e = (int)h_Greedy_2opt_mod_costs(&s, costs);
    *cost = s.cost;
    return e;
}

/* h_extramileage_util (heuristics.c:290) from a given first edge, set up as h_ExtraMileage does (:189-196). */
int refapi_extra_mileage_from(int nodeA, int nodeB, int *path, double *cost) {
    int n = tsp_inst.nnodes;
    tsp_solution s; s.path = path; s.ncomp = 0; s.comp = NULL;
    memset(path, 0, (size_t)n * sizeof(int));
    s.cost = 2 * tsp_get_cost(nodeA, nodeB);
    path[nodeA] = nodeB;
    path[nodeB] = nodeA;
    int e = (int)h_extramileage_util(&s, nodeA, nodeB);
    *cost = s.cost;
    return e;
}

int refapi_tabu_move(int *path, double *cost, int *tabu_list, int tenure, int iter) {
    tabu_search ts; memset(&ts, 0, sizeof(ts));
    ts.tenure = tenure; ts.tabu_list = tabu_list;
    return (int)tabu_best_move(path, cost, &ts, iter);
}

/* tabu_init (metaheuristic.c:65) + one tabu_linear_policy step (:40), exposed so tests can follow
 * the reference's tenure schedule without re-deriving it. */
void refapi_tabu_init(int n, int *tenure, int *min_tenure, int *max_tenure) {
    tabu_search ts; tabu_init(&ts, n);
    *tenure = ts.tenure; *min_tenure = ts.min_tenure; *max_tenure = ts.max_tenure;
    free(ts.tabu_list);
}
int refapi_tabu_linear_step(int *tenure, int min_tenure, int max_tenure, int *increment) {
    tabu_search ts; memset(&ts, 0, sizeof(ts));
    ts.tenure = *tenure; ts.min_tenure = min_tenure; ts.max_tenure = max_tenure; ts.increment = (*increment != 0);
    tsp_env.policy = POL_LINEAR;
    int e = (int)tabu_linear_policy(&ts);
    *tenure = ts.tenure; *increment = ts.increment ? 1 : 0;
    return e;
}

int refapi_vns_kick(int *path) {
    tsp_solution s; s.path = path; s.cost = 0; s.ncomp = 0; s.comp = NULL;
    return (int)vns_kick(&s);
}

/* cx_patching (cplex_model.c:693-755): one patch.  comp / ncomp as cx_build_sol leaves them. */
void cx_patching(tsp_solution *solution);
int refapi_patching(int *path, int *comp, int *ncomp, double *cost) {
    tsp_solution s; s.path = path; s.cost = *cost; s.ncomp = *ncomp; s.comp = comp;
    cx_patching(&s);
    *cost = s.cost; *ncomp = s.ncomp;
    return 0;
}

double refapi_solution_cost(const int *path) { return tsp_solution_cost((int *)path); }   /* src/tsp.c:730 */

/* Whole-algorithm entry points, exactly as src/main.c:4-87 dispatches them (minus plotting).
 * alg: 0 GREEDY, 1 GREEDY_ITER, 2 2OPT_GREEDY, 3 TABU_SEARCH, 4 VNS, 5 EXTRA_MILEAGE with EM_MAX (tsp.h:45-50). */
int refapi_run(int alg, int k, int *best_path, double *best_cost, int *starting_node) {
    int n = tsp_inst.nnodes;
    refapi_reset_incumbent();
    tsp_env.k = k;
    tsp_inst.alg = (algorithms)alg;
    utils_startclock(&tsp_inst.c);
    int e;
    switch (alg) {
        case 0: e = (int)h_Greedy(); break;
        case 1: e = (int)h_Greedy_iterative(); break;
        case 2: e = (int)h_greedy_2opt(); break;
        case 3: e = (int)mh_TabuSearch(); break;
        case 4: e = (int)mh_VNS(); break;
        case 5: tsp_env.mileage_init = EM_MAX; e = (int)h_ExtraMileage(); break;
        default: return -1;
    }
    memcpy(best_path, tsp_inst.best_solution.path, (size_t)n * sizeof(int));
    *best_cost = tsp_inst.best_solution.cost;
    *starting_node = tsp_inst.starting_node;
    return e;
}
