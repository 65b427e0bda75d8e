// matrix.cuh -- the `costs`-argument paths: nearest neighbour and 2-opt on a caller-supplied dense
// matrix of arbitrary doubles (h_greedyutil / ref_2opt with costs != tsp_inst.costs, and
// h_Greedy_2opt_mod_costs, heuristics.c:118-149, whose x*-weighted matrices come from the CPLEX
// relaxation callback, cplex_model.c:1187-1217).
//
// Nothing can be recomputed from coordinates here, and deltas are genuine FP64 values whose rounding
// depends on the expression order, so the kernels evaluate exactly the reference's expressions
//     current = c[a,sa] + c[b,sb];  swapped = c[a,b] + c[sa,sb];  delta = swapped - current
// (refinment.c:60-62, a < b node ids; the two sums are commutative in IEEE arithmetic, the matrix
// entries are read at the reference's own indices so that asymmetric matrices behave identically)
// and reduce (delta, a, b) lexicographically -- the reference's first strict minimum in ascending
// (a,b) order -- in two steps across CTAs: 64-bit atomicMin on the order-preserving image of delta,
// then on (a << 32 | b) among the pairs that attain it.  The scan is bound by the two gathers per
// pair into the n x n matrix (L2 for the sizes a branch-and-cut callback sees).
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace tspcuda {

struct MxState {
    double cost;               // running tour cost (recomputed in node order, then += best_delta per move)
    double last_delta;
    long long moves, log_count;
    int status, last_a, last_b, pad;
};
struct MxMove { int a, b; double delta; };

__device__ __forceinline__ unsigned long long ordered_bits(double d) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(d);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

struct MxBest {
    double delta;
    int a, b;
};
// reference order: strict '<' on delta while scanning (a,b) ascending => smallest (a,b) among equal deltas
__device__ __forceinline__ bool mx_better(double d, int a, int b, const MxBest &o) {
    if (a < 0) return false;
    if (o.a < 0) return true;
    return d < o.delta || (d == o.delta && (a < o.a || (a == o.a && b < o.b)));
}

// ---- nearest neighbour on a matrix (heuristics.c:216-288), one CTA per start ---------------------
constexpr int MX_NN_BLOCK = 512;

__global__ void __launch_bounds__(MX_NN_BLOCK)
nn_matrix_kernel(const double *__restrict__ C, int n, int first_start, int stride, unsigned char *visited_all,
                 int *path_all, int *order_all, double *cost_out, int *len_out) {
    __shared__ double s_val[MX_NN_BLOCK / 32];
    __shared__ int s_idx[MX_NN_BLOCK / 32];
    __shared__ int s_curr;
    const int sidx = blockIdx.x;
    const int start = first_start + sidx * stride;
    unsigned char *visited = visited_all + (size_t)sidx * n;
    int *path = path_all + (size_t)sidx * n;
    int *order = order_all + (size_t)sidx * n;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int i = t; i < n; i += MX_NN_BLOCK) { visited[i] = (i == start); path[i] = -1; }
    if (t == 0) { s_curr = start; order[0] = start; }
    __syncthreads();
    double cost = 0;                                   // meaningful in thread 0 only (sequential FP64 order)
    int len = 1;
    for (;;) {
        const int curr = s_curr;
        const double *row = C + (size_t)curr * (size_t)n;
        double bv = 0;
        int bi = -1;
        for (int i = t; i < n; i += MX_NN_BLOCK) {
            if (i != curr && !visited[i]) {
                const double v = row[i];
                if (v != -1.0 && (bi < 0 || v < bv)) { bv = v; bi = i; }     // NOT_CONNECTED skipped (:257); ascending i per thread
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double ov = __shfl_down_sync(FULL, bv, off);
            const int oi = __shfl_down_sync(FULL, bi, off);
            if (oi >= 0 && (bi < 0 || ov < bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
        }
        if (lane == 0) { s_val[warp] = bv; s_idx[warp] = bi; }
        __syncthreads();
        if (warp == 0) {
            bv = (lane < MX_NN_BLOCK / 32) ? s_val[lane] : 0.0;
            bi = (lane < MX_NN_BLOCK / 32) ? s_idx[lane] : -1;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double ov = __shfl_down_sync(FULL, bv, off);
                const int oi = __shfl_down_sync(FULL, bi, off);
                if (oi >= 0 && (bi < 0 || ov < bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
            }
            if (lane == 0) {
                if (bi < 0) {
                    path[curr] = start;                 // close the path (:268-271), possibly early
                    s_curr = -1;
                } else {
                    path[curr] = bi;
                    visited[bi] = 1;
                    order[len] = bi;
                    cost += bv;                         // :277
                    s_curr = bi;
                }
            }
        }
        __syncthreads();
        if (s_curr < 0) {
            if (t == 0) {
                cost += row[start];                     // :282, curr is the last node reached
                cost_out[sidx] = cost;
                len_out[sidx] = len;
            }
            break;
        }
        ++len;
    }
}

// ---- 2-opt on a matrix ---------------------------------------------------------------------------
struct MxTour {
    int *tid;        // [n+1] node at position p, tid[n] = tid[0]
    int *pos;        // [n]
    double *dnp;     // [n+1] c[tid[p], tid[p+1]], dnp[n] = dnp[0]
    int n;
};

__global__ void mx_tour_init_kernel(MxTour T, const int *order) {
    const int n = T.n;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p <= n; p += gridDim.x * blockDim.x) {
        const int node = order[p == n ? 0 : p];
        T.tid[p] = node;
        if (p < n) T.pos[node] = p;
    }
}

__device__ __forceinline__ void mx_edges(const MxTour &T, const double *C, int t, int nthreads) {
    const int n = T.n;
    for (int p = t; p <= n; p += nthreads) {
        const int q = p == n ? 0 : p;
        T.dnp[p] = C[(size_t)T.tid[q] * (size_t)n + T.tid[q + 1]];
    }
}

// refinment.c:95-114 in position space on (tid, pos) only; same deferred end-node rule as apply_move
__device__ __forceinline__ void mx_apply(const MxTour &T, int a, int b, int t, int nthreads, int &s_out, int &e_out) {
    const int n = T.n;
    const int i = T.pos[a], j = T.pos[b];
    const int s = (i + 1 == n) ? 0 : i + 1;
    const int e = j;
    int L = j - i; if (L < 0) L += n;
    const int half = L >> 1;
    s_out = s; e_out = e;
    for (int k = t; k < half; k += nthreads) {
        int p = s + k; if (p >= n) p -= n;
        int q = e - k; if (q < 0) q += n;
        const int tp = T.tid[p], tq = T.tid[q];
        T.tid[p] = tq; T.tid[q] = tp;
        if (k != 0) { T.pos[tq] = p; T.pos[tp] = q; }
        if (p == 0) T.tid[n] = tq;
        if (q == 0) T.tid[n] = tp;
    }
}

constexpr int MX_BLOCK = 256;

__global__ void __launch_bounds__(MX_BLOCK)
descend_matrix_kernel(const double *__restrict__ C, MxTour T, unsigned long long *slots, MxState *st, MxMove *log,
                      long long log_cap, long long max_moves, int recompute_cost) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    constexpr int WARPS = MX_BLOCK / 32;
    __shared__ MxBest s_best[WARPS];
    __shared__ MxBest s_cta;
    __shared__ unsigned long long s_word;               // one reader per CTA of the words the whole grid needs (see K7, pruned.cuh)
    const int n = T.n;
    const size_t N = (size_t)n;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gtid = blockIdx.x * MX_BLOCK + threadIdx.x;
    const int gsize = gridDim.x * MX_BLOCK;
    const int gw = blockIdx.x * WARPS + warp;
    const int W = gridDim.x * WARPS;

    mx_edges(T, C, gtid, gsize);
    if (recompute_cost && gtid == 0) {                  // refinment.c:6-9: node order, sequential FP64 sum
        double s = 0;
        for (int i = 0; i < n; i++) s += C[(size_t)i * N + T.tid[T.pos[i] + 1]];
        st->cost = s;
    }
    grid.sync();

    for (long long it = 0; it < max_moves; ++it) {
        unsigned long long *slot_d = slots + 2 * (it & 1), *slot_ab = slot_d + 1;
        MxBest best = {0.0, -1, -1};                    // best_delta starts at 0, strict '<' (refinment.c:41,63)
        for (int i = gw; i < n - 1; i += W) {           // pair rows by tour position, one warp per row
            const int u = T.tid[i], su = T.tid[i + 1];
            const double du = T.dnp[i];
            const double *row_u = C + (size_t)u * N, *row_su = C + (size_t)su * N;
            for (int j = i + 2 + lane; j < n; j += 32) {        // j = i+1 is tour-adjacent (refinment.c:55)
                if (i == 0 && j == n - 1) continue;             // so is (0, n-1)
                const int v = T.tid[j], sv = T.tid[j + 1];
                const double current = (u < v) ? du + T.dnp[j] : T.dnp[j] + du;
                const double swapped = (u < v) ? row_u[v] + row_su[sv] : C[(size_t)v * N + u] + C[(size_t)sv * N + su];
                const double delta = swapped - current;
                const int a = min(u, v), b = max(u, v);
                if (delta < 0.0 && mx_better(delta, a, b, best)) { best.delta = delta; best.a = a; best.b = b; }
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            MxBest o;
            o.delta = __shfl_down_sync(FULL, best.delta, off);
            o.a = __shfl_down_sync(FULL, best.a, off);
            o.b = __shfl_down_sync(FULL, best.b, off);
            if (mx_better(o.delta, o.a, o.b, best)) best = o;
        }
        if (lane == 0) s_best[warp] = best;
        __syncthreads();
        if (threadIdx.x == 0) {
            MxBest cb = s_best[0];
            for (int w = 1; w < WARPS; ++w)
                if (mx_better(s_best[w].delta, s_best[w].a, s_best[w].b, cb)) cb = s_best[w];
            s_cta = cb;
            if (cb.a >= 0) atomicMin(slot_d, ordered_bits(cb.delta));
        }
        grid.sync();
        if (threadIdx.x == 0) s_word = *(volatile unsigned long long *)slot_d;
        __syncthreads();
        const unsigned long long dmin = s_word;
        if (threadIdx.x == 0 && s_cta.a >= 0 && ordered_bits(s_cta.delta) == dmin)
            atomicMin(slot_ab, ((unsigned long long)(unsigned)s_cta.a << 32) | (unsigned)s_cta.b);
        grid.sync();
        if (dmin == ~0ull) {                             // no pair with delta < 0
            if (gtid == 0) st->status = 1;
            break;
        }
        if (threadIdx.x == 0) s_word = *(volatile unsigned long long *)slot_ab;      // (everybody has read dmin: a grid barrier lies between)
        __syncthreads();
        const unsigned long long ab = s_word;
        const int a = (int)(ab >> 32), b = (int)(ab & 0xFFFFFFFFu);
        // delta of the winner: recompute from its key bits (exact inverse of ordered_bits)
        const unsigned long long db = (dmin >> 63) ? (dmin & 0x7FFFFFFFFFFFFFFFull) : ~dmin;
        const double delta = __longlong_as_double((long long)db);
        if (!(delta < -1.0E-7)) {                        // EPSILON, tsp.h:19: scan found no applicable move
            if (gtid == 0) { st->status = 1; st->last_delta = delta; }
            break;
        }
        int seg_s, seg_e;
        mx_apply(T, a, b, gtid, gsize, seg_s, seg_e);
        if (gtid == 0) {
            slots[2 * ((it + 1) & 1)] = ~0ull;
            slots[2 * ((it + 1) & 1) + 1] = ~0ull;
            const long long lc = st->log_count;
            if (lc < log_cap) { log[lc].a = a; log[lc].b = b; log[lc].delta = delta; }
            st->log_count = lc + 1;
            st->moves += 1;
            st->cost += delta;                           // refinment.c:85
            st->last_a = a; st->last_b = b; st->last_delta = delta;
        }
        grid.sync();
        if (gtid == 0) { T.pos[b] = seg_s; T.pos[T.tid[seg_e]] = seg_e; }
        mx_edges(T, C, gtid, gsize);                     // directed edge costs: reversed edges change entry
        grid.sync();
    }
}

}  // namespace tspcuda
