/*
 * oracle/onfly.c -- TEST INFRASTRUCTURE: CPU restatement of the reference's heuristic hot path.
 *
 * This file is the parity ORACLE for libtspcuda.so.  It is never linked into, imported by or
 * called from the product; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it (as the checker or as the timed CPU baseline, never as the
 * thing shipped).  It restates, function by function, what the reference computes, but with
 * 64-bit indexing and distances evaluated on the fly from the coordinates, so that it also runs
 * where the reference cannot (n > 46 340: `int` n*n overflow at src/tsp.c:614; 59 GB matrix).
 *
 * Pinning: tests/test_oracle_vs_reference.py checks every function below move-for-move against
 * the UNMODIFIED reference objects (oracle/_ref/libtspref.so via oracle/ref_api.c) wherever the
 * reference can run, and tests/test_oracle_golden.py checks it against the golden values of
 * BASELINE.md section 3 (tests/golden/golden.json).  EUC_REF is therefore PINNED.
 * CEIL_2D / ATT / GEO have NO reference behaviour (src/tsp.c:576-583 rejects them): for those
 * weight types this oracle follows the TSPLIB95 definitions and says so: "parity unpinned".
 *
 * Build: gcc -O3 -fopenmp -ffp-contract=off -fno-math-errno (no -march=native, no -ffast-math):
 * the reference's effective flags are -O3 on baseline x86-64, i.e. no FMA contraction.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { ORC_EUC_REF = 0, ORC_CEIL_2D = 1, ORC_ATT = 2, ORC_GEO = 3 };

/* ERROR_CODE values used on this path, src/utils/errors.h:33-51 */
enum { ORC_T_OK = 0, ORC_INVALID_ARGUMENT = 3, ORC_INTERNAL = 13, ORC_UNAVAILABLE = 14 };

#define ORC_EPSILON (-1.0E-7)            /* src/tsp.h:19 */
#define ORC_NOT_CONNECTED (-1.0)         /* src/utils/utils.h:35 (-1.0f) */

/* ---------------------------------------------------------------------------------------------
 * Edge cost.  EUC_REF follows src/tsp.c:629 exactly:
 *     (double)((int)(sqrtf(pow(dx,2) + pow(dy,2)) + 0.5))
 * pow(.,2) is a plain multiply under gcc -O3; the sum is formed in double, narrowed to float by
 * the sqrtf() argument conversion, the float square root is widened back to double, 0.5 is added
 * in double, the result truncated to a 32-bit int and widened again.
 * ------------------------------------------------------------------------------------------- */
static inline double orc_geo_rad(double v) {
    /* TSPLIB95 section 2.4: PI = 3.141592; deg = nint-toward-zero part, min = remainder */
    const double PI = 3.141592;
    double deg = (double)(long)v;
    double min = v - deg;
    return PI * (deg + 5.0 * min / 3.0) / 180.0;
}

static inline double orc_edge(int wt, double xi, double yi, double xj, double yj) {
    double dx = xj - xi, dy = yj - yi;
    switch (wt) {
        case ORC_EUC_REF: {
            float s = (float)(dx * dx + dy * dy);
            return (double)((int)((double)sqrtf(s) + 0.5));
        }
        case ORC_CEIL_2D:   /* TSPLIB95 2.3: ceil of the double Euclidean distance */
            return (double)((int)ceil(sqrt(dx * dx + dy * dy)));
        case ORC_ATT: {     /* TSPLIB95 2.5 pseudo-Euclidean */
            double r = sqrt((dx * dx + dy * dy) / 10.0);
            double t = (double)(long)(r + 0.5);
            return (t < r) ? t + 1.0 : t;
        }
        case ORC_GEO: {     /* TSPLIB95 2.4; x = latitude, y = longitude in DDD.MM */
            const double RRR = 6378.388;
            double lati = orc_geo_rad(xi), loni = orc_geo_rad(yi);
            double latj = orc_geo_rad(xj), lonj = orc_geo_rad(yj);
            double q1 = cos(loni - lonj), q2 = cos(lati - latj), q3 = cos(lati + latj);
            return (double)((int)(RRR * acos(0.5 * ((1.0 + q1) * q2 - (1.0 - q1) * q3)) + 1.0));
        }
    }
    return ORC_NOT_CONNECTED;
}

double orc_edge_cost(int wt, double xi, double yi, double xj, double yj) { return orc_edge(wt, xi, yi, xj, yj); }

/* FNV-1a-64 over the successor array, one uint32 unit per entry. */
uint64_t orc_fnv1a(const int *path, int n) {
    uint64_t h = 0xcbf29ce484222325ULL;
    for (int i = 0; i < n; i++) { h ^= (uint64_t)(uint32_t)path[i]; h *= 0x100000001b3ULL; }
    return h;
}

/* src/tsp.c:468-481 tsp_generate_randompoints + TSP_RAND() (src/utils/utils.h:23-26): glibc
 * srand(seed), then x, y per point, each ((double)rand()/RAND_MAX)*(5000 - -5000) + -5000. */
void orc_random_points(int n, int seed, double *xy) {
    srand((unsigned)seed);
    for (int i = 0; i < n; i++) {
        xy[2 * i] = ((double)rand() / RAND_MAX) * (5000 - (-5000)) + (-5000);
        xy[2 * i + 1] = ((double)rand() / RAND_MAX) * (5000 - (-5000)) + (-5000);
    }
}

/* src/tsp.c:608-636 tsp_compute_costs: row-major n*n doubles, diagonal -1.0, symmetric. */
int orc_cost_matrix(const double *xy, int n, int wt, double *out) {
    if (n <= 0) return ORC_INVALID_ARGUMENT;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        double *row = out + (size_t)i * (size_t)n;
        for (int j = 0; j < n; j++)
            row[j] = (i == j) ? ORC_NOT_CONNECTED : orc_edge(wt, xy[2 * i], xy[2 * i + 1], xy[2 * j], xy[2 * j + 1]);
    }
    return ORC_T_OK;
}

/* src/tsp.c:687-728 tsp_is_tour + :642-667 tsp_validate_solution: permutation and single cycle. */
int orc_is_tour(const int *path, int n) {
    if (n <= 0) return 0;
    unsigned char *seen = (unsigned char *)calloc((size_t)n, 1);
    int cur = 0, ok = 1;
    for (int k = 0; k < n; k++) {
        if (cur < 0 || cur >= n || seen[cur]) { ok = 0; break; }
        seen[cur] = 1;
        cur = path[cur];
    }
    if (ok && cur != 0) ok = 0;
    free(seen);
    return ok;
}

/* src/tsp.c:730-736 tsp_solution_cost / refinment.c:6-9: sum over node ids of c[i, path[i]]. */
double orc_tour_cost(const double *xy, int n, int wt, const int *path) {
    double c = 0;
    for (int i = 0; i < n; i++) c += orc_edge(wt, xy[2 * i], xy[2 * i + 1], xy[2 * path[i]], xy[2 * path[i] + 1]);
    return c;
}

/* src/algorithms/heuristics.c:216-288 h_greedyutil: nearest-neighbour tour from `start`;
 * ascending-i scan with strict '<' (:253-263) => lowest node id wins ties; path[curr] = argmin;
 * the last node closes to start (:271); cost accumulated in visit order plus closing edge. */
int orc_nn(const double *xy, int n, int wt, int start, int *path, double *cost) {
    if (xy == NULL) return ORC_INTERNAL;                       /* :218-221 (costs == NULL) */
    if (start >= n || start < 0) return ORC_UNAVAILABLE;       /* :223-226 */
    unsigned char *visited = (unsigned char *)calloc((size_t)n, 1);
    int curr = start;
    visited[curr] = 1;
    double sol = 0;
    for (;;) {
        int min_idx = -1;
        double min_dist = __DBL_MAX__;
        double cx = xy[2 * curr], cy = xy[2 * curr + 1];
        for (int i = 0; i < n; i++) {
            if (i != curr && !visited[i]) {
                double t = orc_edge(wt, cx, cy, xy[2 * i], xy[2 * i + 1]);
                if (t != ORC_NOT_CONNECTED && t < min_dist) { min_dist = t; min_idx = i; }
            }
        }
        path[curr] = min_idx;
        if (min_idx == -1) { path[curr] = start; break; }
        visited[min_idx] = 1;
        curr = min_idx;
        sol += min_dist;
    }
    sol += (curr == start) ? ORC_NOT_CONNECTED
                           : orc_edge(wt, xy[2 * curr], xy[2 * curr + 1], xy[2 * start], xy[2 * start + 1]);
    /* n == 1: the reference adds costs[start*n+start] = -1.0 (:282); kept for fidelity */
    *cost = sol;
    free(visited);
    return ORC_T_OK;
}

/* heuristics.c:34-72 h_Greedy_iterative restricted to starts [s0, s0+count): ascending order,
 * strict '<' on cost (:58) => the lowest start wins ties.  Returns the best start. */
int orc_nn_all(const double *xy, int n, int wt, int s0, int count, int *best_path, double *best_cost, double *costs_out) {
    int best = -1;
    double bc = __DBL_MAX__;
#pragma omp parallel
    {
        int *p = (int *)malloc((size_t)n * sizeof(int));
#pragma omp for schedule(dynamic, 1)
        for (int k = 0; k < count; k++) {
            double c;
            orc_nn(xy, n, wt, s0 + k, p, &c);
            if (costs_out) costs_out[k] = c;
#pragma omp critical
            {
                if (c < bc || (c == bc && s0 + k < best)) { bc = c; best = s0 + k; if (best_path) memcpy(best_path, p, (size_t)n * sizeof(int)); }
            }
        }
        free(p);
    }
    *best_cost = bc;
    return best;
}

/* ---------------------------------------------------------------------------------------------
 * 2-opt scan.  Restates the pair loop of refinment.c:49-69 (tabu == NULL, init 0) and of
 * metaheuristic.c:198-222 (tabu != NULL, init DBL_MAX, 4-node tabu filter :208-210, :416-418).
 * Enumerates NODE-ID pairs a<b; skip rule :55; delta = (c[a,b] + c[sa,sb]) - (c[a,sa] + c[b,sb])
 * in that expression order (:60-62); strict '<' => the lexicographically smallest (a,b) wins
 * among equal deltas.  OpenMP over a with an ordered (delta,a,b) merge reproduces that.
 * ------------------------------------------------------------------------------------------- */
typedef struct { double delta; int a, b; } orc_move;

/* One row of the pair loop, written so gcc can vectorise it (sqrtps / cvtpd2ps / cvttpd2dq are
 * the packed forms of exactly the scalar IEEE operations, so results are bit-identical):
 * out[b] = (c[a,b] + c[sa,sb]) - (c[a,sa] + c[b,sb]) for b in (a, n).  The skip rule and the
 * argmin are applied afterwards by the caller. */
#define ORC_ROW_FN(NAME, WT)                                                                          \
    __attribute__((target_clones("default", "avx2", "avx512f"))) /* no "fma": IEEE-identical */       \
    static void NAME(int a, int n, const double *x, const double *y, const double *sx,                \
                     const double *sy, const double *dn, double *out) {                               \
        double ax = x[a], ay = y[a], sax = sx[a], say = sy[a], da = dn[a];                            \
        for (int b = a + 1; b < n; b++) {                                                             \
            double current = da + dn[b];                                                              \
            double swapped = orc_edge(WT, ax, ay, x[b], y[b]) + orc_edge(WT, sax, say, sx[b], sy[b]); \
            out[b] = swapped - current;                                                               \
        }                                                                                             \
    }
ORC_ROW_FN(orc_row_euc, ORC_EUC_REF)
ORC_ROW_FN(orc_row_ceil, ORC_CEIL_2D)
ORC_ROW_FN(orc_row_att, ORC_ATT)
ORC_ROW_FN(orc_row_geo, ORC_GEO)

static orc_move orc_scan_rows(const double *xy, int n, int wt, const int *path, double init,
                              const int *tabu_list, int tenure, int iter, int a_lo, int a_hi);

static orc_move orc_scan(const double *xy, int n, int wt, const int *path, double init,
                         const int *tabu_list, int tenure, int iter) {
    return orc_scan_rows(xy, n, wt, path, init, tabu_list, tenure, iter, 0, n - 1);
}

/* The pair loop restricted to rows a in [a_lo, a_hi) (all b > a): the whole scan when
 * [0, n-1); a bounded sample of it for CPU-baseline timing. */
static orc_move orc_scan_rows(const double *xy, int n, int wt, const int *path, double init,
                              const int *tabu_list, int tenure, int iter, int a_lo, int a_hi) {
    double *x = (double *)malloc((size_t)n * sizeof(double));
    double *y = (double *)malloc((size_t)n * sizeof(double));
    double *sx = (double *)malloc((size_t)n * sizeof(double));
    double *sy = (double *)malloc((size_t)n * sizeof(double));
    double *dn = (double *)malloc((size_t)n * sizeof(double));
    int *prev = (int *)malloc((size_t)n * sizeof(int));
    unsigned char *tb = (unsigned char *)calloc((size_t)n, 1);   /* node is tabu */
    unsigned char *tm = (unsigned char *)calloc((size_t)n, 1);   /* node or its successor is tabu */
    for (int i = 0; i < n; i++) {
        int s = path[i];
        x[i] = xy[2 * i]; y[i] = xy[2 * i + 1];
        sx[i] = xy[2 * s]; sy[i] = xy[2 * s + 1];
        prev[s] = i;                                             /* refinment.c:43-46 */
        dn[i] = (s == i) ? ORC_NOT_CONNECTED : orc_edge(wt, x[i], y[i], sx[i], sy[i]);
        if (tabu_list)   /* is_in_tabu_list, metaheuristic.c:416-418 */
            tb[i] = (iter - tabu_list[i] < tenure && tabu_list[i] != -1) ? 1 : 0;
    }
    if (tabu_list) for (int i = 0; i < n; i++) tm[i] = tb[i] | tb[path[i]];
    orc_move best = { init, -1, -1 };
#pragma omp parallel
    {
        orc_move loc = { init, -1, -1 };
        double *row = (double *)malloc((size_t)n * sizeof(double));
#pragma omp for schedule(dynamic, 8) nowait
        for (int a = a_lo; a < a_hi; a++) {
            if (tm[a]) continue;                                  /* metaheuristic.c:208 (a or succ_a tabu) */
            switch (wt) {
                case ORC_EUC_REF: orc_row_euc(a, n, x, y, sx, sy, dn, row); break;
                case ORC_CEIL_2D: orc_row_ceil(a, n, x, y, sx, sy, dn, row); break;
                case ORC_ATT: orc_row_att(a, n, x, y, sx, sy, dn, row); break;
                default: orc_row_geo(a, n, x, y, sx, sy, dn, row); break;
            }
            int sa = path[a], pa = prev[a];
            /* refinment.c:55 skip rule: succ_a == succ_b (never, a != b on a permutation),
             * a == succ_b  <=> b == prev[a],  b == succ_a. */
            if (sa > a) row[sa] = __DBL_MAX__;
            if (pa > a) row[pa] = __DBL_MAX__;
            double rmin = __DBL_MAX__;
            if (!tabu_list) {                                     /* cheap reject: row minimum cannot win */
                for (int b = a + 1; b < n; b++) rmin = row[b] < rmin ? row[b] : rmin;
                if (rmin > loc.delta || (rmin == loc.delta && loc.a >= 0)) continue;   /* a only grows per thread */
            }
            for (int b = a + 1; b < n; b++) {
                if (b == sa || b == pa || path[b] == sa) continue;
                if (tm[b]) continue;                              /* metaheuristic.c:208 (b or succ_b tabu) */
                double delta = row[b];
                if (delta < loc.delta || (delta == loc.delta && loc.a >= 0 && (a < loc.a || (a == loc.a && b < loc.b)))) {
                    loc.delta = delta; loc.a = a; loc.b = b;
                }
            }
        }
        free(row);
#pragma omp critical
        {
            if (loc.a >= 0 && (loc.delta < best.delta ||
                (loc.delta == best.delta && best.a >= 0 && (loc.a < best.a || (loc.a == best.a && loc.b < best.b))))) best = loc;
        }
    }
    free(x); free(y); free(sx); free(sy); free(dn); free(prev); free(tb); free(tm);
    return best;
}

/* Same scan over an explicit row-major cost matrix (any doubles, e.g. the fractional matrices of
 * heuristics.c:118-149 h_Greedy_2opt_mod_costs).  size_t indexing. */
static orc_move orc_scan_matrix(const double *c, int n, const int *path, double init,
                                const int *tabu_list, int tenure, int iter) {
    orc_move best = { init, -1, -1 };
    size_t N = (size_t)n;
#pragma omp parallel
    {
        orc_move loc = { init, -1, -1 };
#pragma omp for schedule(dynamic, 16) nowait
        for (int a = 0; a < n - 1; a++) {
            int sa = path[a];
            for (int b = a + 1; b < n; b++) {
                int sb = path[b];
                if (sa == sb || a == sb || b == sa) continue;
                if (tabu_list) {
                    int t = 0;
                    int v[4] = { a, b, sa, sb };
                    for (int k = 0; k < 4; k++) t |= (iter - tabu_list[v[k]] < tenure && tabu_list[v[k]] != -1);
                    if (t) continue;
                }
                double current = c[a * N + sa] + c[b * N + sb];
                double swapped = c[a * N + b] + c[sa * N + sb];
                double delta = swapped - current;
                if (delta < loc.delta || (delta == loc.delta && loc.a >= 0 && (a < loc.a || (a == loc.a && b < loc.b)))) {
                    loc.delta = delta; loc.a = a; loc.b = b;
                }
            }
        }
#pragma omp critical
        {
            if (loc.a >= 0 && (loc.delta < best.delta ||
                (loc.delta == best.delta && best.a >= 0 && (loc.a < best.a || (loc.a == best.a && loc.b < best.b))))) best = loc;
        }
    }
    return best;
}

/* refinment.c:95-114 ref_reverse_path: path[a]=b; path[sa]=sb; walk prev from b until sa. */
static void orc_reverse(int a, int sa, int b, int sb, const int *prev, int *path) {
    path[a] = b;
    path[sa] = sb;
    int cur = b;
    for (;;) {
        int nd = prev[cur];
        path[cur] = nd;
        cur = nd;
        if (nd == sa) break;
    }
}

static void orc_apply(int n, int *path, int a, int b) {
    int *prev = (int *)malloc((size_t)n * sizeof(int));
    for (int i = 0; i < n; i++) prev[path[i]] = i;             /* refinment.c:43-46 */
    orc_reverse(a, path[a], b, path[b], prev, path);
    free(prev);
}

/* Scan only (no move applied): the argmin a GPU scan must reproduce.  Returns delta; a,b = -1 if
 * no pair beat `init` (0 for 2-opt, DBL_MAX for tabu). */
double orc_two_opt_scan(const double *xy, int n, int wt, const int *path, int *a, int *b) {
    orc_move m = orc_scan(xy, n, wt, path, 0.0, NULL, 0, 0);
    *a = m.a; *b = m.b;
    return m.delta;
}

/* Bounded sample of one scan: rows a in [a_lo, a_hi).  Returns the number of (a,b) pair
 * iterations it covered (for pair-evals/s), the partial argmin in *a,*b,*delta. */
long long orc_two_opt_scan_rows(const double *xy, int n, int wt, const int *path, int a_lo, int a_hi,
                                int *a, int *b, double *delta) {
    if (a_lo < 0) a_lo = 0;
    if (a_hi > n - 1) a_hi = n - 1;
    orc_move m = orc_scan_rows(xy, n, wt, path, 0.0, NULL, 0, 0, a_lo, a_hi);
    if (a) *a = m.a;
    if (b) *b = m.b;
    if (delta) *delta = m.delta;
    long long pairs = 0;
    for (int r = a_lo; r < a_hi; r++) pairs += (long long)(n - 1 - r);
    return pairs;
}

/* refinment.c:39-93 ref_2opt_once: scan, and if best_delta < EPSILON apply + cost += delta. */
double orc_two_opt_once(const double *xy, int n, int wt, int *path, double *cost, int *a_out, int *b_out) {
    orc_move m = orc_scan(xy, n, wt, path, 0.0, NULL, 0, 0);
    if (m.delta < ORC_EPSILON) {                                /* :74 */
        orc_apply(n, path, m.a, m.b);
        *cost += m.delta;                                       /* :85 */
    }
    if (a_out) *a_out = m.a;
    if (b_out) *b_out = m.b;
    return m.delta;
}

/* refinment.c:3-37 ref_2opt without the time limit and incumbent update (host-side in the
 * reference): recompute cost (:6-9), then loop ref_2opt_once while delta < EPSILON (:15-27).
 * max_moves < 0 = to the local optimum.  Optional move log.  Returns the number of moves. */
long orc_two_opt(const double *xy, int n, int wt, int *path, double *cost, long max_moves,
                 int *log_a, int *log_b, double *log_delta) {
    *cost = orc_tour_cost(xy, n, wt, path);
    long moves = 0;
    while (max_moves < 0 || moves < max_moves) {
        int a, b;
        double d = orc_two_opt_once(xy, n, wt, path, cost, &a, &b);
        if (!(d < ORC_EPSILON)) break;
        if (log_a) { log_a[moves] = a; log_b[moves] = b; }
        if (log_delta) log_delta[moves] = d;
        moves++;
    }
    return moves;
}

/* heuristics.c:156-210 h_ExtraMileage (EM_MAX initialisation :169-178) + :290-367 h_extramileage_util.
 * nodeA < 0: EM_MAX -- the first pair (i<j, ascending) with the strictly largest cost, (0,1) if every
 * cost is 0.  Then n-2 insertions: over unvisited i ascending and tour edges j in edges[] order, the
 * first strict minimum of c(u,i) + c(i,v) - c(u,v) (:321-339); edges[j] becomes (u,i), (i,v) is
 * appended (:348-356); cost starts at 2 c(A,B) (:193) and grows by the mileage (:362).  O(n^3). */
int orc_extra_mileage(const double *xy, int n, int wt, int nodeA, int nodeB, int *path, double *cost) {
    if (n < 2) return ORC_INVALID_ARGUMENT;
#define ORC_C(i, j) orc_edge(wt, xy[2 * (i)], xy[2 * (i) + 1], xy[2 * (j)], xy[2 * (j) + 1])
    if (nodeA < 0) {
        double maxd = 0.0;
        nodeA = 0; nodeB = 1;
        for (int i = 0; i < n; i++)
            for (int j = i + 1; j < n; j++) {
                double d = ORC_C(i, j);
                if (d > maxd) { nodeA = i; nodeB = j; maxd = d; }
            }
    }
    unsigned char *visited = (unsigned char *)calloc((size_t)n, 1);
    int *eu = (int *)malloc(sizeof(int) * (size_t)n), *ev = (int *)malloc(sizeof(int) * (size_t)n);
    for (int i = 0; i < n; i++) path[i] = 0;              /* tsp_init_solution: calloc'd path */
    visited[nodeA] = visited[nodeB] = 1;
    int m = 0;
    eu[m] = nodeA; ev[m++] = nodeB;
    eu[m] = nodeB; ev[m++] = nodeA;
    double sol = 2 * ORC_C(nodeA, nodeB);
    path[nodeA] = nodeB; path[nodeB] = nodeA;
    while (m < n) {
        double mileage = __DBL_MAX__;
        int bj = -1, bi = -1;
#pragma omp parallel
        {
            double lm = __DBL_MAX__;
            int lj = -1, li = -1;
#pragma omp for schedule(static) nowait
            for (int i = 0; i < n; i++) {
                if (visited[i]) continue;
                for (int j = 0; j < m; j++) {
                    int u = eu[j], v = ev[j];
                    double d = ORC_C(u, i) + ORC_C(i, v) - ORC_C(u, v);
                    if (d < lm) { lm = d; lj = j; li = i; }
                }
            }
#pragma omp critical
            {   /* static schedule: threads own ascending i ranges; keep the (delta, i) lexicographic minimum */
                if (li >= 0 && (lm < mileage || (lm == mileage && li < bi))) { mileage = lm; bj = lj; bi = li; }
            }
        }
        if (bj == -1) break;
        int u = eu[bj], v = ev[bj];
        path[u] = bi; ev[bj] = bi;
        path[bi] = v; eu[m] = bi; ev[m] = v; m++;
        visited[bi] = 1;
        sol += mileage;
    }
#undef ORC_C
    *cost = sol;
    free(visited); free(eu); free(ev);
    return ORC_T_OK;
}

/* heuristics.c:216-288 h_greedyutil over an explicit row-major matrix (the `costs` argument: the
 * instance matrix or the fractional matrices of h_Greedy_2opt_mod_costs, heuristics.c:118-149).
 * Entries equal to NOT_CONNECTED (-1.0) are skipped (:257); if no candidate is left the tour is
 * closed to `start` (:268-271) -- early when every remaining entry is -1.0, as in the reference.
 * The cost is accumulated in visit order, closing edge last (:277, :282). */
int orc_nn_matrix(const double *c, int n, int start, int *path, double *cost) {
    if (c == NULL) return ORC_INTERNAL;
    if (start >= n || start < 0) return ORC_UNAVAILABLE;
    size_t N = (size_t)n;
    unsigned char *visited = (unsigned char *)calloc(N, 1);
    int curr = start;
    visited[curr] = 1;
    double sol = 0;
    for (;;) {
        int min_idx = -1;
        double min_dist = __DBL_MAX__;
        const double *row = c + (size_t)curr * N;
        for (int i = 0; i < n; i++) {
            if (i != curr && !visited[i]) {
                double t = row[i];
                if (t != ORC_NOT_CONNECTED && t < min_dist) { min_dist = t; min_idx = i; }
            }
        }
        path[curr] = min_idx;
        if (min_idx == -1) { path[curr] = start; break; }
        visited[min_idx] = 1;
        curr = min_idx;
        sol += min_dist;
    }
    sol += c[(size_t)curr * N + start];
    *cost = sol;
    free(visited);
    return ORC_T_OK;
}

/* heuristics.c:118-149 h_Greedy_2opt_mod_costs without a time limit: every start in ascending order,
 * NN + ref_2opt(update_incumbent = false) into the SAME solution struct, nothing kept between
 * starts -- so what the caller receives is the result of the LAST start, n-1. */
long orc_two_opt_matrix(const double *c, int n, int *path, double *cost, long max_moves, int *log_a, int *log_b, double *log_delta);
int orc_mod_costs(const double *c, int n, int *path, double *cost) {
    int e = ORC_T_OK;
    for (int s = 0; s < n; s++) {
        e = orc_nn_matrix(c, n, s, path, cost);
        if (e != ORC_T_OK) continue;
        orc_two_opt_matrix(c, n, path, cost, -1, NULL, NULL, NULL);
    }
    return ORC_T_OK;
}

long orc_two_opt_matrix(const double *c, int n, int *path, double *cost, long max_moves,
                        int *log_a, int *log_b, double *log_delta) {
    size_t N = (size_t)n;
    double s = 0;
    for (int i = 0; i < n; i++) s += c[i * N + path[i]];
    long moves = 0;
    while (max_moves < 0 || moves < max_moves) {
        orc_move m = orc_scan_matrix(c, n, path, 0.0, NULL, 0, 0);
        if (!(m.delta < ORC_EPSILON)) break;
        orc_apply(n, path, m.a, m.b);
        s += m.delta;
        if (log_a) { log_a[moves] = m.a; log_b[moves] = m.b; }
        if (log_delta) log_delta[moves] = m.delta;
        moves++;
    }
    *cost = s;
    return moves;
}

/* metaheuristic.c:188-245 tabu_best_move: best admissible (possibly worsening) move; apply it,
 * cost += delta, stamp tabu_list[a,b,sa,sb] = iter (:231-239).  Always T_OK.  a_out = -1 if every
 * pair was tabu. */
int orc_tabu_move(const double *xy, int n, int wt, int *path, double *cost, int *tabu_list,
                  int tenure, int iter, int *a_out, int *b_out, double *delta_out) {
    orc_move m = orc_scan(xy, n, wt, path, __DBL_MAX__, tabu_list, tenure, iter);
    if (m.delta < __DBL_MAX__ && m.a >= 0) {
        int a = m.a, b = m.b, sa = path[a], sb = path[b];
        orc_apply(n, path, a, b);
        *cost += m.delta;
        tabu_list[a] = iter; tabu_list[b] = iter; tabu_list[sa] = iter; tabu_list[sb] = iter;
    }
    if (a_out) *a_out = m.a;
    if (b_out) *b_out = m.b;
    if (delta_out) *delta_out = m.delta;
    return ORC_T_OK;
}

/* metaheuristic.c:65-84 tabu_init tenure bounds and :40-59 tabu_linear_policy. */
void orc_tabu_init(int n, int *tenure, int *min_tenure, int *max_tenure) {
    *tenure = (int)(0.125 * n + 1);       /* MIN_FRACTION * nnodes + 1, metaheuristic.h:12 */
    *max_tenure = (int)(0.25 * n);        /* MAX_FRACTION, metaheuristic.h:13 */
    *min_tenure = (int)(0.125 * n);
}
void orc_tabu_linear_step(int *tenure, int min_tenure, int max_tenure, int *increment) {
    if (max_tenure == *tenure || min_tenure == *tenure) *increment = !*increment;
    if (*increment) (*tenure)++; else (*tenure)--;
}

/* metaheuristic.c:344-409 vns_kick + :490-500 tabu_make_move case 7, driven by glibc rand().
 * Restated including the position/node-id mix-up of the rejection test (:372).  The reference reads
 * tour[-1] and tour[n] out of bounds there (indexes[j] = 0 resp. n-1).  Under glibc both are 0 in
 * practice -- tour[-1] is the upper half of the chunk's size field, tour[n] lies in the zero-filled
 * slack of the calloc'd chunk unless 4n fills it exactly (n = 2 mod 4: the next chunk's size field,
 * heap dependent) -- so 0 is what is restated: after index n-1 was accepted, a draw of 0 is rejected.
 * Checked against the reference objects on small n, where the case is frequent
 * (tests/test_oracle_vs_reference.py::test_vns_kick_small_n). */
void orc_vns_kick(int n, int *path) {
    int *tour = (int *)malloc((size_t)n * sizeof(int));
    int node = 0;
    for (int idx = 0; idx < n; idx++) { tour[idx] = node; node = path[node]; }
    int idxs[3];
    for (int i = 0; i < 3; i++) {
        int r;
        do {
            r = rand() % n;
            for (int j = 0; j < i; j++) {
                int lo = idxs[j] - 1, hi = idxs[j] + 1;
                int tlo = (lo >= 0) ? tour[lo] : 0, thi = (hi < n) ? tour[hi] : 0;
                if (r == idxs[j] || r == tlo || r == thi) { r = -1; break; }
            }
        } while (r == -1);
        idxs[i] = r;
        for (int j = i; j > 0; j--)
            if (idxs[j] < idxs[j - 1]) { int t = idxs[j]; idxs[j] = idxs[j - 1]; idxs[j - 1] = t; }
    }
    int A = tour[idxs[0]], sA = tour[(idxs[0] + 1) % n];
    int B = tour[idxs[1]], sB = tour[(idxs[1] + 1) % n];
    int C = tour[idxs[2]], sC = tour[(idxs[2] + 1) % n];
    path[A] = sB; path[C] = sA; path[B] = sC;                 /* case 7, :490-496 */
    free(tour);
}

/* cplex_model.c:693-755 cx_patching: ONE patch of a successor array made of several cycles (comp[v] =
 * label of the cycle of v, *ncomp of them).  Ordered pairs (a,b), a-major / b-minor, of different
 * components (:699-706); delta = [c(a,succ b) + c(b,succ a)] - [c(a,succ a) + c(b,succ b)] (:712-714);
 * kept on strict '<' AND delta > 0 (:716).  Then path[a] = succ b, path[b] = succ a, cost += delta, the
 * cycle of b is relabelled with comp[a] by walking from succ b until b, ncomp-- (:727-754).
 * Returns 1 and the pair in *a_out / *b_out / *delta_out (any may be NULL) if a patch was applied; 0 and
 * no change when no pair has delta > 0 (the reference then indexes path[-1], :730-735: undefined).
 * Costs on the fly with 64-bit-safe indexing (the reference reads tsp_inst.costs through tsp_get_cost). */
int orc_patching(const double *xy, int n, int wt, int *path, int *comp, int *ncomp, double *cost, int *a_out,
                 int *b_out, double *delta_out) {
    double best_delta = __DBL_MAX__;
    int ba = -1, bb = -1;
    for (int a = 0; a < n; a++) {
        int sa = path[a];
        double casa = orc_edge(wt, xy[2 * a], xy[2 * a + 1], xy[2 * sa], xy[2 * sa + 1]);
        for (int b = 0; b < n; b++) {
            if (comp[a] == comp[b]) continue;
            int sb = path[b];
            double current = casa + orc_edge(wt, xy[2 * b], xy[2 * b + 1], xy[2 * sb], xy[2 * sb + 1]);
            double swapped = orc_edge(wt, xy[2 * a], xy[2 * a + 1], xy[2 * sb], xy[2 * sb + 1]) +
                             orc_edge(wt, xy[2 * b], xy[2 * b + 1], xy[2 * sa], xy[2 * sa + 1]);
            double delta = swapped - current;
            if (delta < best_delta && delta > 0) { best_delta = delta; ba = a; bb = b; }
        }
    }
    if (ba < 0) return 0;
    int sa = path[ba], sb = path[bb];
    path[ba] = sb;
    path[bb] = sa;
    *cost += best_delta;
    int label = comp[ba];
    for (int i = sb;; i = path[i]) {          /* path[bb] already points at sa, as in the reference's walk */
        comp[i] = label;
        if (i == bb) break;
    }
    *ncomp -= 1;
    if (a_out) *a_out = ba;
    if (b_out) *b_out = bb;
    if (delta_out) *delta_out = best_delta;
    return 1;
}

/* ---------------------------------------------------------------------------------------------
 * Audit of a GPU move log on instances too big for a full CPU descent:
 * replays `nmoves` logged moves (a,b,delta) from `path`; every logged delta is recomputed in
 * O(1); at the iterations listed in full_at[] (ascending) a full scan confirms the logged move
 * is the argmin.  Returns 0 if everything matches, else 1 + index of the first mismatch.
 * path is left at the state after the last replayed move.
 * ------------------------------------------------------------------------------------------- */
long orc_audit_moves(const double *xy, int n, int wt, int *path, long nmoves, const int *log_a, const int *log_b,
                     const double *log_delta, const long *full_at, long nfull) {
    long fi = 0;
    for (long m = 0; m < nmoves; m++) {
        int a = log_a[m], b = log_b[m];
        if (a < 0 || b <= a || b >= n) return 1 + m;
        int sa = path[a], sb = path[b];
        double cur = orc_edge(wt, xy[2 * a], xy[2 * a + 1], xy[2 * sa], xy[2 * sa + 1]) +
                     orc_edge(wt, xy[2 * b], xy[2 * b + 1], xy[2 * sb], xy[2 * sb + 1]);
        double sw = orc_edge(wt, xy[2 * a], xy[2 * a + 1], xy[2 * b], xy[2 * b + 1]) +
                    orc_edge(wt, xy[2 * sa], xy[2 * sa + 1], xy[2 * sb], xy[2 * sb + 1]);
        if (sw - cur != log_delta[m]) return 1 + m;
        if (fi < nfull && full_at[fi] == m) {
            orc_move best = orc_scan(xy, n, wt, path, 0.0, NULL, 0, 0);
            if (best.a != a || best.b != b || best.delta != log_delta[m]) return 1 + m;
            fi++;
        }
        orc_apply(n, path, a, b);
    }
    return 0;
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int t) {
#ifdef _OPENMP
    omp_set_num_threads(t);
#else
    (void)t;
#endif
}
