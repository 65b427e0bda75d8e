// pruned.cuh -- K7: exact 2-opt best-improvement scan over CANDIDATE pairs only.
//
// The reference's scan (refinment.c:49-69) looks at all n(n-1)/2 pairs, but only an improving pair
// can win (best_delta starts at 0, strict '<', :63), and for the pair of tour positions {p,q}
//     delta = D(p,q) + D(p+1,q+1) - dn[p] - dn[q] < 0   ==>   D(p,q) < dn[p]  or  D(p+1,q+1) < dn[q].
// So every improving pair is found by two fixed-radius searches per node u (at position p):
//     nodes w with cost(u,w) < cost(u, succ u)   -> pair {p, pos w}
//     nodes w with cost(u,w) < cost(pred u, u)   -> pair {p-1, pos w - 1}
// Nodes never move, so the searches run over a static uniform grid of the node coordinates (built
// once per instance).  Every candidate gets the SAME exact integer delta and the SAME packed
// (delta, min id, max id) key as the full scan, and the set of candidates contains every pair with
// delta < 0: the argmin -- ties included -- is identical to the full scan's by construction, and
// "no candidate improves" proves local optimality.  tests/test_gpu_parity.py checks move logs of the
// two kernels against each other and against the CPU restatement of the reference.
//
// Round 2: the 2-opt kernels (K7, K8) no longer search and evaluate in every scan: they keep the records of the searches
// with their move deltas and STREAM them (the delta cache, second half of this file); the search-and-evaluate phases
// described next remain the scan of the tabu loop (K5c), whose bound changes from iteration to iteration.
//
// Work distribution.  Phase 1: one thread per tour position; search boxes holding at most
// PR_SMALL_NODES nodes are walked by that thread, larger ones (the long edges of a nearest-neighbour
// tour) are queued.  Phase 2 (after a grid barrier): the queued boxes are cut into (box, cell row)
// units -- a cell row of a box is one contiguous range of the cell-sorted node array -- and the units
// are split evenly over all warps of the grid, lanes striding over the nodes of a row.
#pragma once
#include <type_traits>

#include "kernels.cuh"

namespace tspcuda {

struct SpatialGrid {
    double gx0, gy0, inv_cell;
    int W, H;
    const int *cell_start;      // [W*H + 1] exclusive prefix of nodes per cell (row-major cells)
    const int *cnode;           // [n] node ids sorted by cell
    const double2 *cxy;         // [n] their coordinates, same order
};

struct BigBox {
    int p, cy0, cx0, cx1, nrows, pad;
    long long base;             // first unit (row) index of this box in the flattened unit space
};

#ifndef TSP_PRUNED_BLOCK
#define TSP_PRUNED_BLOCK 512
#endif
#ifndef TSP_PR_SMALL_NODES
#define TSP_PR_SMALL_NODES 32
#endif
constexpr int PRUNED_BLOCK = TSP_PRUNED_BLOCK;
#ifndef TSP_PR_NC
#define TSP_PR_NC 4
#endif
#ifndef TSP_PR_NC_BIG
#define TSP_PR_NC_BIG 4
#endif
constexpr int PR_NC_BIG = TSP_PR_NC_BIG;    // the same for 1024-thread CTAs (64 registers per thread)
constexpr int PR_NC = TSP_PR_NC;            // candidates evaluated together (latency overlap), phase 1
#ifndef TSP_PR_NC2
#define TSP_PR_NC2 2
#endif
constexpr int PR_NC2 = TSP_PR_NC2;          // the same per lane in phase 2
constexpr int PR_SMALL_ROWS = 12;
constexpr int PR_SMALL_NODES = TSP_PR_SMALL_NODES;
constexpr unsigned long long PR_UNIT_MASK = (1ull << 40) - 1;

// ---- grid construction ------------------------------------------------------------------------
__device__ __forceinline__ int grid_cell_of(const SpatialGrid &G, double x, double y) {
    int cx = __double2int_rd((x - G.gx0) * G.inv_cell);
    int cy = __double2int_rd((y - G.gy0) * G.inv_cell);
    cx = max(0, min(G.W - 1, cx));
    cy = max(0, min(G.H - 1, cy));
    return cy * G.W + cx;
}

__global__ void grid_count_kernel(SpatialGrid G, const double2 *xy, int n, int *count) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        atomicAdd(&count[grid_cell_of(G, xy[i].x, xy[i].y)], 1);
}

// single-CTA exclusive scan of count[0..m) into start[0..m], start[m] = total; count is zeroed so
// that it can serve as the fill cursor afterwards
__global__ void __launch_bounds__(1024) grid_scan_kernel(int *count, int *start, int m) {
    __shared__ int s_part[1024];
    const int t = threadIdx.x;
    const int per = (m + 1023) / 1024;
    const int lo = min(m, t * per), hi = min(m, lo + per);
    int sum = 0;
    for (int k = lo; k < hi; ++k) sum += count[k];
    s_part[t] = sum;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        int v = (t >= off) ? s_part[t - off] : 0;
        __syncthreads();
        s_part[t] += v;
        __syncthreads();
    }
    int run = s_part[t] - sum;
    for (int k = lo; k < hi; ++k) { const int c = count[k]; start[k] = run; run += c; count[k] = 0; }
    if (t == 1023) start[m] = s_part[1023];
}

__global__ void grid_fill_kernel(SpatialGrid G, const double2 *xy, int n, int *cursor, int *cnode, double2 *cxy) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double2 p = xy[i];
        const int c = grid_cell_of(G, p.x, p.y);
        const int slot = G.cell_start[c] + atomicAdd(&cursor[c], 1);
        cnode[slot] = i;
        cxy[slot] = p;
    }
}

// ---- K4b: nearest neighbour with the grid (heuristics.c:216-288), one CTA per start ---------------
// The plain kernel (K4, kernels.cuh) looks at every unvisited node in every step.  Here a step looks at
// the square of cells of half-width r = 1, 2, 4, ... around the current node: that square contains every
// node closer than r cells, so once the best unvisited node found in it costs less than any node
// farther than that can cost (Metric::cost_floor) it is the reference's choice -- same packed
// (cost, node id) argmin, i.e. lowest cost, then lowest id (heuristics.c:253-263).  When the square would
// hold more nodes than are left, the step scans the compacted list of remaining nodes instead (late in
// the construction the neighbourhood is exhausted and the next node is far away).
constexpr int NNG_BLOCK = 512;

template <class M>
__global__ void __launch_bounds__(NNG_BLOCK)
nn_grid_kernel(const double2 *__restrict__ xy, int n, SpatialGrid G, int first_start, int stride, double2 *cand_xy_all,
               int *cand_id_all, int *slot_all, int *order_all, long long *cost_out) {
    extern __shared__ unsigned s_vis[];              // visited bitmap, (n + 31) / 32 words
    __shared__ unsigned long long s_wkey[NNG_BLOCK / 32];
    __shared__ unsigned long long s_key;
    const int sidx = blockIdx.x;
    const int start = first_start + sidx * stride;
    double2 *cxy = cand_xy_all + (size_t)sidx * n;
    int *cid = cand_id_all + (size_t)sidx * n;
    int *slot = slot_all + (size_t)sidx * n;
    int *order = order_all + (size_t)sidx * n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double cell = 1.0 / G.inv_cell;

    for (int k = tid; k < (n + 31) / 32; k += NNG_BLOCK) s_vis[k] = 0;
    for (int k = tid; k < n - 1; k += NNG_BLOCK) {
        const int node = k + (k >= start ? 1 : 0);
        cxy[k] = xy[node];
        cid[k] = node;
        slot[node] = k;
    }
    if (tid == 0) order[0] = start;
    __syncthreads();
    if (tid == 0) s_vis[start >> 5] |= 1u << (start & 31);
    __syncthreads();
    int cur_node = start;
    long long cost = 0;
    int m = n - 1;                                   // unvisited nodes = live entries of the compacted list
    for (int step = 1; step < n; ++step, --m) {
        const double2 cur = xy[cur_node];
        const int ccx = max(0, min(G.W - 1, __double2int_rd((cur.x - G.gx0) * G.inv_cell)));
        const int ccy = max(0, min(G.H - 1, __double2int_rd((cur.y - G.gy0) * G.inv_cell)));
        unsigned long long found = KEY_EMPTY;
        for (int r = 1;; r *= 2) {
            const int x0 = max(0, ccx - r), x1 = min(G.W - 1, ccx + r), y0 = max(0, ccy - r), y1 = min(G.H - 1, ccy + r);
            const bool whole = x0 == 0 && y0 == 0 && x1 == G.W - 1 && y1 == G.H - 1;
            const long long box_nodes_est = 2LL * (x1 - x0 + 1) * (y1 - y0 + 1);
            unsigned long long bk = KEY_EMPTY;
            if (whole || box_nodes_est > m) {
                // remaining nodes are fewer than the square holds: scan the compacted list (complete by itself)
                for (int k = tid; k < m; k += NNG_BLOCK) {
                    const double2 p = cxy[k];
                    const int d = M::K(cur.x, cur.y, p.x, p.y) - KOFF;
                    bk = min(bk, ((unsigned long long)(unsigned)d << 32) | (unsigned)cid[k]);
                }
            } else {
                const int rows = y1 - y0 + 1;
                for (int ry = warp; ry < rows; ry += NNG_BLOCK / 32) {       // one warp per cell row of the square
                    const int cy = y0 + ry;
                    const int lo = G.cell_start[cy * G.W + x0], hi = G.cell_start[cy * G.W + x1 + 1];
                    for (int k = lo + lane; k < hi; k += 32) {
                        const int w = G.cnode[k];
                        const double2 p = G.cxy[k];          // requested with the id, not behind the visited test: one round trip
                        if ((s_vis[w >> 5] >> (w & 31)) & 1u) continue;
                        const int d = M::K(cur.x, cur.y, p.x, p.y) - KOFF;
                        bk = min(bk, ((unsigned long long)(unsigned)d << 32) | (unsigned)w);
                    }
                }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) bk = min(bk, __shfl_down_sync(FULL, bk, off));
            if (lane == 0) s_wkey[warp] = bk;
            __syncthreads();
            if (warp == 0) {
                bk = (lane < NNG_BLOCK / 32) ? s_wkey[lane] : KEY_EMPTY;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) bk = min(bk, __shfl_down_sync(FULL, bk, off));
                if (lane == 0) s_key = bk;
            }
            __syncthreads();
            found = s_key;
            __syncthreads();                         // s_key / s_wkey are rewritten by the next attempt
            if (whole || box_nodes_est > m) break;   // complete scans
            // nodes outside the square are at least r cells away (the current node lies inside its own cell)
            if (found != KEY_EMPTY && (double)(found >> 32) < M::cost_floor((double)r * cell)) break;
        }
        const int w = (int)(found & 0xFFFFFFFFu);
        if (tid == 0) {
            order[step] = w;
            s_vis[w >> 5] |= 1u << (w & 31);
            const int s = slot[w], last = m - 1;     // swap-remove from the compacted list
            const int lid = cid[last];
            cxy[s] = cxy[last]; cid[s] = lid; slot[lid] = s;
        }
        cost += (long long)(found >> 32);
        cur_node = w;
        __syncthreads();
    }
    if (tid == 0) {
        const double2 cur = xy[cur_node], s0 = xy[start];
        const long long closing = (n > 1) ? (long long)(M::K(cur.x, cur.y, s0.x, s0.y) - KOFF) : -1;   // heuristics.c:282
        cost_out[sidx] = cost + closing;
    }
}

// ---- candidate evaluation ---------------------------------------------------------------------
// Everything a search around the node at position p needs, loaded once.
struct Probe {
    double ux, uy, sx, sy, px, py;     // node at p, its successor (p+1), its predecessor (p-1)
    int p, pm, ida, idm, dnp, dnm;     // pm = p-1 (cyclic); node ids at p and pm; K-offset edge costs = search thresholds
};

// TABU: the search is for the best ADMISSIBLE pair with delta < bound: both thresholds grow by
// h = ceil(bound / 2) (delta < bound implies D(p,q) - dn[p] < h or D(p+1,q+1) - dn[q] < h), and an edge
// that touches a tabu node (dns flagged: dn - BIGPEN, i.e. below KOFF) starts no search at all.
template <bool TABU>
__device__ __forceinline__ void probe_load(const TourView &V, int p, Probe &P, int h) {
    const int n = V.n;
    P.p = p;
    P.pm = p == 0 ? n - 1 : p - 1;
    const double2 u = V.pxy[p], s = V.pxy[p + 1], q = V.pxy[P.pm];
    P.ux = u.x; P.uy = u.y; P.sx = s.x; P.sy = s.y; P.px = q.x; P.py = q.y;
    P.ida = V.tid[p]; P.idm = V.tid[P.pm];
    P.dnp = V.dn[p]; P.dnm = V.dn[P.pm];
    if (TABU) {                                       // dnp / dnm become the search thresholds
        P.dnp = (V.dns[p] < KOFF) ? 0 : P.dnp + h;
        P.dnm = (V.dns[P.pm] < KOFF) ? 0 : P.dnm + h;
    }
}

// pos[] of the two end nodes of the segment reversed by the PREVIOUS move is written by one thread
// right after the barrier that ends that move (every thread reads pos[a], pos[b] on entry of
// apply_move, so it cannot be written earlier); scans that start behind the same barrier take those
// two positions from registers instead of racing with that write.
// node2 / node3: the other two nodes whose tour edges the previous move changed (a and succ b); together with node0 (b)
// and node1 (succ a) they are the only nodes whose cached neighbour list (NbCache below) is out of date.
struct PosPatch { int node0, pos0, node1, pos1, node2, node3; };
#define POSPATCH_NONE {-1, 0, -1, 0, -1, -1}

// Cell box that contains every node w with cost(u,w) < max(dn[p], dn[p-1]).  false: empty search.
template <class M>
__device__ __forceinline__ bool probe_box(const SpatialGrid &G, const Probe &P, int &cx0, int &cx1, int &cy0, int &cy1) {
    const int rr = max(P.dnp, P.dnm) - KOFF;
    if (rr <= 0) return false;
    const double Rad = M::radius((double)rr);
    cx0 = max(0, min(G.W - 1, __double2int_rd((P.ux - Rad - G.gx0) * G.inv_cell)));
    cx1 = max(0, min(G.W - 1, __double2int_rd((P.ux + Rad - G.gx0) * G.inv_cell)));
    cy0 = max(0, min(G.H - 1, __double2int_rd((P.uy - Rad - G.gy0) * G.inv_cell)));
    cy1 = max(0, min(G.H - 1, __double2int_rd((P.uy + Rad - G.gy0) * G.inv_cell)));
    return true;
}

// ---- collect / evaluate ---------------------------------------------------------------------------
// A search only needs the candidate's coordinates to decide whether it can matter (cost(u,w) below one
// of the two tour-edge costs at u); everything after that -- pos[w], the neighbours' coordinates, the
// edge costs at q, the second new edge -- is a chain of dependent gathers.  The searches therefore only
// COLLECT the passing (p, w, cost, which-edge) records into a buffer private to the warp (ballot
// compaction, no atomics), and the records are EVALUATED afterwards one per lane, several in flight
// per lane: every lane busy whatever the box sizes were, and no gather chain inside the search loops.
#ifndef TSP_PR_WBUF
#define TSP_PR_WBUF 512
#endif
constexpr int PR_WBUF = TSP_PR_WBUF;        // records per warp buffer (16 bytes each); flushed when nearly full

template <class M, int NE, bool TABU>
__device__ __forceinline__ void pairs_eval(const TourView &V, const PosPatch &F, const uint4 *wbuf, int count, int &best,
                                           unsigned long long &bestkey) {
    const int n = V.n;
    const int lane = threadIdx.x & 31;
    if (lane == 0) NBSTAT(5, count);
    for (int base = 0; base < count; base += 32 * NE) {
        uint4 e[NE];
        int q[NE];
        bool ok[NE];
#pragma unroll
        for (int j = 0; j < NE; ++j) {
            const int idx = base + lane + 32 * j;
            e[j] = wbuf[min(idx, count - 1)];
            if (idx >= count) e[j].w = 0;
        }
#pragma unroll
        for (int j = 0; j < NE; ++j) q[j] = V.pos[e[j].y];
#pragma unroll
        for (int j = 0; j < NE; ++j) {
            if ((int)e[j].y == F.node0) q[j] = F.pos0;
            if ((int)e[j].y == F.node1) q[j] = F.pos1;
            const int diff = abs((int)e[j].x - q[j]);
            ok[j] = diff > 1 && diff != n - 1;                   // same node / tour neighbours: refinment.c:55
        }
        {   // pairs {p, q}: edges (p,p+1), (q,q+1) out, (p,q), (p+1,q+1) in
            double2 s1[NE], w1[NE];
            int dp[NE], dq[NE];
#pragma unroll
            for (int j = 0; j < NE; ++j) {
                const bool c1 = ok[j] && (e[j].w & 1u);
                const int p = (int)e[j].x;
                s1[j] = V.pxy[c1 ? p + 1 : 0];
                w1[j] = V.pxy[c1 ? q[j] + 1 : 0];
                dp[j] = V.dn[c1 ? p : 0];
                dq[j] = (TABU ? V.dns : V.dn)[c1 ? q[j] : 0];          // TABU: a flagged edge (below KOFF) makes the pair inadmissible
            }
#pragma unroll
            for (int j = 0; j < NE; ++j) {
                if (ok[j] && (e[j].w & 1u) && (!TABU || dq[j] >= KOFF)) {
                    const int K2 = M::K(s1[j].x, s1[j].y, w1[j].x, w1[j].y);
                    const int d = (int)(e[j].z + (unsigned)K2 - (unsigned)dp[j] - (unsigned)dq[j]);
                    if (d <= best) {
                        best = d;
                        const unsigned long long key = pack_key(d, V.tid[e[j].x], (int)e[j].y);
                        if (key < bestkey) bestkey = key;
                    }
                }
            }
        }
        {   // pairs {p-1, q-1}: (p,q) is their second new edge
            double2 s2[NE], w2[NE];
            int dpm[NE], dqm[NE], pm[NE], qm[NE];
#pragma unroll
            for (int j = 0; j < NE; ++j) {
                const bool c2 = ok[j] && (e[j].w & 2u);
                const int p = (int)e[j].x;
                pm[j] = p == 0 ? n - 1 : p - 1;
                qm[j] = q[j] == 0 ? n - 1 : q[j] - 1;
                s2[j] = V.pxy[c2 ? pm[j] : 0];
                w2[j] = V.pxy[c2 ? qm[j] : 0];
                dpm[j] = V.dn[c2 ? pm[j] : 0];
                dqm[j] = (TABU ? V.dns : V.dn)[c2 ? qm[j] : 0];
            }
#pragma unroll
            for (int j = 0; j < NE; ++j) {
                if (ok[j] && (e[j].w & 2u) && (!TABU || dqm[j] >= KOFF)) {
                    const int K2 = M::K(s2[j].x, s2[j].y, w2[j].x, w2[j].y);
                    const int d = (int)((unsigned)K2 + e[j].z - (unsigned)dpm[j] - (unsigned)dqm[j]);
                    if (d <= best) {
                        best = d;
                        const unsigned long long key = pack_key(d, V.tid[pm[j]], V.tid[qm[j]]);
                        if (key < bestkey) bestkey = key;
                    }
                }
            }
        }
    }
}

// One trip of a search: NCC candidates per lane (slot kk[j] of the cell-sorted arrays, valid[j]) of the
// search around position p (node coordinates ux/uy, edge costs dnp = dn[p], dnm = dn[p-1]).
template <class M, int NCC, int NE, bool TABU>
__device__ __forceinline__ void collect_trip(const TourView &V, const SpatialGrid &G, const PosPatch &F, int p, double ux,
                                             double uy, int dnp, int dnm, const int (&kk)[NCC], const bool (&valid)[NCC],
                                             uint4 *wbuf, int &wcount, int &best, unsigned long long &bestkey) {
    const int lane = threadIdx.x & 31;
    if (wcount > PR_WBUF - 32 * NCC) {                 // warp-uniform: flush
        __syncwarp();
        pairs_eval<M, NE, TABU>(V, F, wbuf, wcount, best, bestkey);
        __syncwarp();
        wcount = 0;
    }
    double2 c[NCC];
    int w[NCC];
#pragma unroll
    for (int j = 0; j < NCC; ++j) { c[j] = G.cxy[kk[j]]; w[j] = G.cnode[kk[j]]; }
#pragma unroll
    for (int j = 0; j < NCC; ++j) {
        const int K = M::K(ux, uy, c[j].x, c[j].y);
        const unsigned type = valid[j] ? ((K < dnp ? 1u : 0u) | (K < dnm ? 2u : 0u)) : 0u;
        const unsigned mask = __ballot_sync(FULL, type != 0);
        if (type) wbuf[wcount + __popc(mask & ((1u << lane) - 1))] = make_uint4((unsigned)p, (unsigned)w[j], (unsigned)K, type);
        wcount += __popc(mask);
    }
}

// ---- the two scan phases, shared by the whole-GPU kernel (K7) and the CTA-per-descent kernel (K8) --
// Phase 1: thread `t` of `nthreads` cooperating threads (a grid or one CTA; t's warp is complete and
// nthreads is a multiple of 32) takes tour positions t, t + nthreads, ...  Search boxes with at most
// PR_SMALL_NODES nodes are walked inline, larger ones are appended to boxes[] through the packed
// counter *qctr (boxes in the high 24 bits, cell rows in the low 40; one atomic per warp).
template <class M, int NC1, int NE, int L, bool TABU>
__device__ __forceinline__ void pruned_phase1(const TourView &V, const SpatialGrid &G, const PosPatch &F, int t, int nthreads,
                                              unsigned long long *qctr, BigBox *boxes, uint4 *wbuf, int &best,
                                              unsigned long long &bestkey, long long &ncand, int h) {
    int wcount = 0;
    const int n = V.n;
    const int lane = threadIdx.x & 31;
    // L lanes share a position (L = 4 when there are at least four times as many threads as positions, chosen by
    // the host): lane `sub` of the group walks the cell rows sub, sub + L, ... of the position's box, so the longest
    // per-lane stream of a warp -- what the whole warp waits for -- shrinks accordingly
    const int sub = lane & (L - 1);
    const int per_pass = nthreads / L;                             // positions per pass; nthreads, L: powers-of-two friendly
    for (int p0 = (t - lane) / L; p0 < n; p0 += per_pass) {        // warp-uniform trip count
        const int p = p0 + lane / L;
        Probe P;
        int cx0 = 0, cx1 = 0, cy0 = 0, cy1 = -1;
        bool has = false;
        if (p < n) {
            probe_load<TABU>(V, p, P, h);
            has = probe_box<M>(G, P, cx0, cx1, cy0, cy1);
        }
        const int nrows = has ? cy1 - cy0 + 1 : 0;
        int total = 0, mine = 0;                                   // nodes in the box / in this lane's rows
        if (has) {
            if (nrows <= PR_SMALL_ROWS) {
#pragma unroll
                for (int r = 0; r < PR_SMALL_ROWS; ++r) {          // unrolled: all row bounds in flight at once
                    const int cy = min(cy0 + r, cy1);
                    const int cnt = G.cell_start[cy * G.W + cx1 + 1] - G.cell_start[cy * G.W + cx0];
                    total += (cy0 + r <= cy1) ? cnt : 0;
                    mine += (cy0 + r <= cy1 && (r & (L - 1)) == sub) ? cnt : 0;
                }
            } else {
                total = PR_SMALL_NODES + 1;
            }
        }
        if (sub != 0) has = has && total <= PR_SMALL_NODES;        // a large box is queued once, by the group's first lane
        const bool big = total > PR_SMALL_NODES && sub == 0;
        const unsigned bigmask = __ballot_sync(FULL, big);
        if (bigmask) {                                             // one queue atomic per warp
            const int myrows = big ? nrows : 0;
            int incl = myrows;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int v = __shfl_up_sync(FULL, incl, off);
                if (lane >= off) incl += v;
            }
            const int totrows = __shfl_sync(FULL, incl, 31);
            unsigned long long old = 0;
            if (lane == 0)
                old = atomicAdd(qctr, ((unsigned long long)__popc(bigmask) << 40) | (unsigned long long)totrows);
            old = __shfl_sync(FULL, old, 0);
            if (big) {
                BigBox bx;
                bx.p = p; bx.cy0 = cy0; bx.cx0 = cx0; bx.cx1 = cx1; bx.nrows = nrows; bx.pad = 0;
                bx.base = (long long)(old & PR_UNIT_MASK) + (incl - myrows);
                boxes[(old >> 40) + __popc(bigmask & ((1u << lane) - 1))] = bx;
            }
        }
        // the small boxes of the warp's 32 positions, in lockstep: each lane streams ITS box across its cell
        // rows (row bounds are L1 hits by now), NC1 candidates per trip, into the warp's record buffer
        const int mytotal = (has && !big) ? mine : 0;
        ncand += mytotal;
        const int maxtotal = __reduce_max_sync(FULL, mytotal);
        int cy = cy0 + sub, k = 0, hi = 0, last = 0;
        for (int done = 0; done < maxtotal; done += NC1) {
            int kk[NC1];
            bool valid[NC1];
#pragma unroll
            for (int j = 0; j < NC1; ++j) {
                valid[j] = done + j < mytotal;
                if (valid[j]) {
                    while (k >= hi) {
                        k = G.cell_start[cy * G.W + cx0];
                        hi = G.cell_start[cy * G.W + cx1 + 1];
                        cy += L;
                    }
                    last = k++;
                }
                kk[j] = last;
            }
            collect_trip<M, NC1, NE, TABU>(V, G, F, p, P.ux, P.uy, P.dnp, P.dnm, kk, valid, wbuf, wcount, best, bestkey);
        }
    }
    __syncwarp();
    pairs_eval<M, NE, TABU>(V, F, wbuf, wcount, best, bestkey);
}

// ---- per-node DELTA cache: a scan that streams instead of searching ---------------------------------------
// A search around node u returns the nodes w with cost(u,w) below the larger of u's two tour-edge costs; each such
// RECORD (u,w) stands for the 2-opt moves that create the edge (u,w).  Write x, y for u's two tour neighbours and
// z, z' for w's.  Whatever the direction the tour is walked in, the move pairs x with z and y with z', or x with z'
// and y with z -- which of the two depends only on whether u and w are traversed in the same relative direction as
// when the record was made -- and each of the four deltas
//     cost(u,w) + cost(x,z) - cost(u,x) - cost(w,z),  ...
// depends on the identities of x, y, z, z' only.  A 2-opt move changes the tour edges of exactly FOUR nodes (a, succ a,
// b, succ b); every node strictly inside the reversed segment keeps both edges and merely swaps successor and
// predecessor.  So the cache keeps every record as {w, d0, d1}:
//     d[r] = the smaller of the two deltas that are valid 2-opt moves when  (o8[u] ^ o8[w]) & 1 == r,
// where bit 0 of o8[] is an orientation bit per node that apply_move flips for every node of a reversed segment
// (kernels.cuh).  A scan then is a STREAM over the records: load one, look up the byte o8[w], select d0 or d1, compare
// with the running best.  No position, coordinate or edge-cost gather, and no arithmetic, for a record that cannot win.
// Only records at or below the running best are re-evaluated exactly from the position arrays (pairs_eval: both
// deltas, adjacency test, packed keys), so the argmin -- ties included -- is taken over exact keys as before; the
// record set is the one the searches produce, which contains every pair with delta < 0 (header of this file).
// What is recomputed after a move:
//   * the lists of the four nodes whose edges changed: one warp-wide search each, by four designated warps;
//   * records (u,w) in OTHER lists whose w is one of those four (w's neighbours changed): apply_move stamps the four
//     nodes' o8 bytes (bits 1..7) with the move's epoch, the scan compares the byte it loads anyway with the current
//     epoch and recomputes such a record in place (a stamp that is 126 moves old may alias: a needless but harmless
//     recomputation); a record whose d0 is NB_DIRTY is treated the same way (marked from the host after a VNS kick);
//   * nothing else.
// Layout.  Most nodes have few records -- on a nearest-neighbour tour of pla85900 nine nodes in ten have none, a few
// hundred nodes with a long tour edge have thousands.  So: one 64-byte LINE per node = header word + NB_CAP records of
// three words, and one count byte per node (cnt8) that says whether the line has to be looked at at all; a list longer
// than NB_CAP lives as a contiguous run of {u, w, d0, d1} records in a HEAP (bump allocation; a rebuilt list abandons
// its old run after overwriting the u fields with -1), which the scan streams linearly, evenly split over all
// cooperating threads whatever the list sizes -- no queue, no second phase, no barrier.  When the heap is three
// quarters full the whole cache is declared stale and the heap reset (between two barriers); a list that finds no
// room stays stale and is searched again in every scan: slower, still exact.
constexpr int NB_WORDS = 16;                   // ints per node line
constexpr int NB_CAP = 5;                      // records held in the line itself
constexpr int NB_CF_STALE = 1023, NB_CF_HEAP = 1022;          // header: never searched / list in the heap (words 1, 2: start, count)
constexpr int NB_HDR_STALE = NB_CF_STALE;      // (bits 10..: how often a stale node's giant box has been evaluated without caching, see giant_age)
constexpr int NB_DIRTY = (int)0x80000000;      // d0 of a record that must be recomputed before use
constexpr int NB_GIANT = 512;                  // a search box with more nodes than this is not walked by one warp: giant_phase()
#ifndef TSP_NB_GIANT_KEEP
#define TSP_NB_GIANT_KEEP 1
#endif
constexpr int NB_FLAT = 64;                    // a search box with at most this many nodes is walked in one step (rebuild_list)
constexpr int NB_EPOCHS = 126;                 // epochs 0..125 cycle; 126: "no move yet"; 127: a node no move has touched
constexpr int NB_EP_NONE = 126, NB_EP_INIT = 127;

struct NbCache {
    int *lines;             // [n][NB_WORDS]
    int4 *heap;             // [heap_cap] records {u, w, d0, d1}; u < 0: abandoned
    int *heap_next;         // [0], [1]: bump counters of the heap, [2]: which of the two is in use (they swap at a reset), [3]: live
                            // records, [4]: nodes whose list lives in the heap (8 ints)
    unsigned char *cnt8;    // [n] non-zero: the node's line must be looked at (records, stale); zero: nothing there
    int heap_cap;
    int restart_pct;        // heap restart policy (heap_checkpoint): abandoned records tolerated, in percent of the default rule
    int giant_age;          // a GIANT search box is only evaluated (as in round 1), not cached, the first giant_age times the node is
                            // searched with the same tour edges: the three instance-spanning edges of a VNS kick are gone after
                            // two or three moves -- caching their ~n records each only fills the heap.  0: cache at once (descents).
};
__host__ __device__ inline int nb_cnt8_stride(int n) { return (n + 127) & ~127; }      // read four bytes at a time
// records of heap per node of the instance (default NB_HEAP_PER_NODE; TSPCUDA_NB_HEAP for tests that want the "no room" and
// restart paths taken: tests/test_gpu_northstar.py)
constexpr int NB_HEAP_PER_NODE = 24;
inline int nb_heap_cap(int n, int per_node) {
    const long long cap = (long long)per_node * n + 4096;
    return (int)(cap < 0x3FFFFFF0LL ? cap : 0x3FFFFFF0LL);
}

__global__ void nb_invalidate_kernel(NbCache C, unsigned char *o8, int n) {
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < n; u += gridDim.x * blockDim.x) {
        C.lines[(size_t)u * NB_WORDS] = NB_HDR_STALE;
        C.cnt8[u] = 1;
        o8[u] = NB_EP_INIT << 1;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { C.heap_next[0] = 0; C.heap_next[1] = 0; C.heap_next[2] = 0; C.heap_next[3] = 0; C.heap_next[4] = 0; }
}
// after the tour was changed from outside WITHOUT any reversal (a VNS kick exchanges blocks): the listed nodes got new
// neighbours.  Their own lists are marked stale (first kernel, which also raises flag[] and abandons their heap runs),
// every record pointing at one of them is marked dirty (second kernel), the flags are lowered again (third).
__global__ void nb_mark_nodes_kernel(NbCache C, const int *nodes, int count, unsigned char *flag, int raise) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) {
        const int u = nodes[k];
        flag[u] = (unsigned char)raise;
        if (raise) {
            int *h = C.lines + (size_t)u * NB_WORDS;
            if ((h[0] & 0x3FF) == NB_CF_HEAP) {
                for (int j = 0; j < h[2]; ++j) C.heap[h[1] + j].x = -1;
                atomicSub(C.heap_next + 3, h[2]);
                atomicSub(C.heap_next + 4, 1);
            }
            h[0] = NB_HDR_STALE;
            C.cnt8[u] = 1;
        }
    }
}
__global__ void nb_mark_records_kernel(NbCache C, const unsigned char *flag, int n) {
    const int stride = gridDim.x * blockDim.x, t = blockIdx.x * blockDim.x + threadIdx.x;
    for (int u = t; u < n; u += stride) {
        if (!C.cnt8[u]) continue;
        int *line = C.lines + (size_t)u * NB_WORDS;
        const int cf = line[0] & 0x3FF;
        if (cf > NB_CAP) continue;
        for (int j = 0; j < cf; ++j)
            if (flag[line[1 + 3 * j]]) line[2 + 3 * j] = NB_DIRTY;
    }
    const int hn = min(C.heap_next[C.heap_next[2]], C.heap_cap);
    for (int k = t; k < hn; k += stride) {
        const int4 r = C.heap[k];
        if (r.x >= 0 && flag[r.y]) C.heap[k].z = NB_DIRTY;
    }
}

__device__ __forceinline__ bool in_patch(const PosPatch &F, int node) {
    return node == F.node0 || node == F.node1 || node == F.node2 || node == F.node3;
}
__device__ __forceinline__ int pos_of(const TourView &V, const PosPatch &F, int node) {
    return node == F.node0 ? F.pos0 : node == F.node1 ? F.pos1 : V.pos[node];
}

// what the records of node u need from the tour: its position, both neighbours (ids, coordinates) and both edge costs
struct NodeCtx {
    double2 u, s, m;       // coordinates of the node, its successor, its predecessor
    int p, dnp, dnm, su, pu;
};
__device__ __forceinline__ void node_ctx(const TourView &V, const PosPatch &F, int node, NodeCtx &c) {
    const int n = V.n;
    c.p = pos_of(V, F, node);
    const int pm = c.p == 0 ? n - 1 : c.p - 1;
    c.u = V.pxy[c.p]; c.s = V.pxy[c.p + 1]; c.m = V.pxy[pm];
    c.dnp = V.dn[c.p]; c.dnm = V.dn[pm];
    c.su = V.tid[c.p + 1]; c.pu = V.tid[pm];
}

// the four deltas of record (u,w), K = K-offset cost(u,w): dsame = the moves valid in the CURRENT relative direction
// ({p,q}: successors paired, {p-1,q-1}: predecessors paired), dflip = those valid once exactly one of u, w has turned
template <class M>
__device__ __forceinline__ void record_deltas(const TourView &V, const PosPatch &F, const NodeCtx &c, int w, int K, int &dsame, int &dflip) {
    const int n = V.n;
    const int q = pos_of(V, F, w), qm = q == 0 ? n - 1 : q - 1;
    const double2 sw = V.pxy[q + 1], pw = V.pxy[qm];
    const unsigned dnq = (unsigned)V.dn[q], dnqm = (unsigned)V.dn[qm];
    const unsigned kss = (unsigned)M::K(c.s.x, c.s.y, sw.x, sw.y), kpp = (unsigned)M::K(c.m.x, c.m.y, pw.x, pw.y);
    const unsigned ksp = (unsigned)M::K(c.s.x, c.s.y, pw.x, pw.y), kps = (unsigned)M::K(c.m.x, c.m.y, sw.x, sw.y);
    const unsigned k = (unsigned)K;
    const int dss = (int)(k + kss - (unsigned)c.dnp - dnq), dpp = (int)(k + kpp - (unsigned)c.dnm - dnqm);
    const int dsp = (int)(k + ksp - (unsigned)c.dnp - dnqm), dps = (int)(k + kps - (unsigned)c.dnm - dnq);
    dsame = min(dss, dpp);
    dflip = min(dsp, dps);
}

// The streaming loops contain no call and no gather other than o8[]: whatever needs the position arrays -- a dirty
// record, a stale node -- is appended to a warp-private ATTENTION buffer (the first half of the warp's record buffer) and
// dealt with by one separate (noinline) function after the loops, or when the buffer runs full.  Records that tie or
// beat the running best go to the second half as node-id records for the exact evaluation.  The running best is shared:
// reduced over the warp after every step and exchanged between the warps of a CTA through one shared-memory word.
constexpr int AB_CAP = PR_WBUF / 2;            // attention entries per warp
constexpr int EB_CAP = PR_WBUF - AB_CAP;       // evaluation records per warp
constexpr unsigned AT_DIRTY = 1, AT_NODE = 2;
constexpr unsigned EV_NODEIDS = 4;             // evaluation record {u, w, -, 3 | EV_NODEIDS}: node ids instead of {p, w, K, type}
constexpr int NB_RING = 256;                   // per-warp ring of pending node ids (shared memory): 31 left over + 128 new at most
#ifndef TSP_NB_HU
#define TSP_NB_HU 4
#endif
constexpr int NB_HU = TSP_NB_HU;               // heap records in flight per lane
#ifndef TSP_NB_SPLIT_WARPS
#define TSP_NB_SPLIT_WARPS 256
#endif
constexpr int NB_SPLIT_WARPS = TSP_NB_SPLIT_WARPS;   // cooperating warps from which on the four searching warps take no share of the stream

// ---- the heap stream through shared memory (K8): bulk asynchronous copies completing on mbarriers -----------------
// MEASUREMENT ONLY (-DTSP_NB_STAGE=1; profiles/r02_pruned.md): built to test the reading "K8's scan waits for its heap records"
// (ncu: 14 % of the stall samples at the first use of a record).  Same tours, and SLOWER than the direct loads the library
// ships -- pla85900 148 starts 354 against 315 us per scan with 4 stages per warp, 347 with 3, 388 with 5: the more shared memory
// the ring takes from L1, the slower the o8[] gathers and the line reads that the scan does next to the stream.  The stream's
// latency is not what limits K8.
#ifndef TSP_NB_STAGE
#define TSP_NB_STAGE 0
#endif
#ifndef TSP_NB_STAGES
#define TSP_NB_STAGES 4
#endif
constexpr int NB_STAGES = TSP_NB_STAGES;       // stages per warp
constexpr int NB_STAGE_RECS = 64;              // records per stage (1 KB)
struct StageRing {
    int4 *buf;                    // [NB_STAGES][NB_STAGE_RECS], shared memory, private to the warp
    unsigned long long *bar;      // [NB_STAGES] mbarriers (shared memory), one arrival (the issuing lane's expect_tx) per use
    unsigned phase;               // bit st: the parity the next wait on stage st expects
};
__host__ __device__ constexpr size_t stage_smem_bytes(int warps) { return (size_t)warps * NB_STAGES * (NB_STAGE_RECS * sizeof(int4) + sizeof(unsigned long long)); }
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
// global -> shared, size a multiple of 16 bytes, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// heap records are written by ordinary stores (generic proxy) and read by the bulk copies of LATER scans (async proxy):
// every writer fences across the proxies before the barrier that ends its phase
__device__ __forceinline__ void heap_writes_to_async_proxy() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

#ifdef TSP_PHASE_TIMING
// sections of cached_scan as warp 1 of CTA 0 sees them (a warp that does not search): ns in the heap stream, the lines, the
// attention / evaluation tail; [3] = scans
__device__ unsigned long long g_scan_ns[4];
#define SCAN_MARK(slot) if (blockIdx.x == 0 && threadIdx.x == 32) { unsigned long long tn_; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tn_)); atomicAdd(&g_scan_ns[slot], tn_ - scan_t_); scan_t_ = tn_; }
#else
#define SCAN_MARK(slot)
#endif

struct ScanAcc {
    unsigned long long bestkey;
    long long ncand;
    int best, ecount;
};

// evaluate the warp's evaluation half: node-id records are completed to {p, w, K, type} first
template <class M, int NE>
__device__ __noinline__ ScanAcc flush_records(TourView V, PosPatch F, uint4 *ebuf, ScanAcc a) {
    __syncwarp();
    for (int idx = threadIdx.x & 31; idx < a.ecount; idx += 32) {
        uint4 e = ebuf[idx];
        if (e.w & EV_NODEIDS) {
            const int p = pos_of(V, F, (int)e.x);
            const double2 x = V.pxy[p], y = V.pxy[pos_of(V, F, (int)e.y)];
            ebuf[idx] = make_uint4((unsigned)p, e.y, (unsigned)M::K(x.x, x.y, y.x, y.y), 3u);
        }
    }
    __syncwarp();
    pairs_eval<M, NE, false>(V, F, ebuf, a.ecount, a.best, a.bestkey);
    __syncwarp();
    a.ecount = 0;
    return a;
}

// append the warp's records (hit lanes) to the evaluation half, evaluating it first when it is nearly full
template <class M, int NE>
__device__ __forceinline__ void eval_push(const TourView &V, const PosPatch &F, uint4 *ebuf, ScanAcc &a, bool hit, int p, int w, int K) {
    const unsigned mask = __ballot_sync(FULL, hit);
    if (mask == 0) return;                                         // warp-uniform
    if (a.ecount > EB_CAP - 32) a = flush_records<M, NE>(V, F, ebuf, a);
    if (hit) ebuf[a.ecount + __popc(mask & ((1u << (threadIdx.x & 31)) - 1))] = make_uint4((unsigned)p, (unsigned)w, (unsigned)K, 3u);
    a.ecount += __popc(mask);
}

template <class M>
__device__ __forceinline__ bool node_box(const SpatialGrid &G, const NodeCtx &c, int &cx0, int &cx1, int &cy0, int &cy1) {
    const int rr = max(c.dnp, c.dnm) - KOFF;
    if (rr <= 0) return false;                                     // both tour edges cost 0: nothing can be cheaper
    const double Rad = M::radius((double)rr);
    cx0 = max(0, min(G.W - 1, __double2int_rd((c.u.x - Rad - G.gx0) * G.inv_cell)));
    cx1 = max(0, min(G.W - 1, __double2int_rd((c.u.x + Rad - G.gx0) * G.inv_cell)));
    cy0 = max(0, min(G.H - 1, __double2int_rd((c.u.y - Rad - G.gy0) * G.inv_cell)));
    cy1 = max(0, min(G.H - 1, __double2int_rd((c.u.y + Rad - G.gy0) * G.inv_cell)));
    return true;
}

// Warp-wide search around node u: rewrites u's list and hands the records that can win NOW to the exact evaluation.
// Two passes over the search box: the first counts the records (costs only), so that a long list gets a heap run of
// its size class; the second computes the deltas and writes.  A node keeps its run while its list fits the run's
// capacity (header words 1..3: start, count, capacity), so the many nodes that are searched again and again with lists of
// six to a few dozen records do not leave dead records behind.
// SETTLE: only bring the cache up to date (no candidates): used when a launch ends right after a move.
template <class M, int NE, bool SETTLE>
__device__ __forceinline__ void rebuild_list(const TourView &V, const SpatialGrid &G, const PosPatch &F, const NbCache &C, int hp, int u, bool fresh,
                                             unsigned long long *qctr, BigBox *boxes, uint4 *ebuf, ScanAcc &a) {
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1;
    int *line = C.lines + (size_t)u * NB_WORDS;
    const int h0 = line[0], ostart = line[1], ocount = line[2], ocap = line[3];
    const bool had = (h0 & 0x3FF) == NB_CF_HEAP;
    __syncwarp();                                                  // every lane has the old header before lane 0 rewrites it
    NodeCtx c;
    node_ctx(V, F, u, c);
    int cx0 = 0, cx1 = 0, cy0 = 0, cy1 = -1;
    const bool any = node_box<M>(G, c, cx0, cx1, cy0, cy1);
    const int nrows = any ? cy1 - cy0 + 1 : 0;
    const int thr = max(c.dnp, c.dnm);
    int count = 0, tot = 0;
    int rlo = 0, rcnt = 0;                                         // lane r: the cell-sorted range of box row r
    if (lane < nrows) { rlo = G.cell_start[(cy0 + lane) * G.W + cx0]; rcnt = G.cell_start[(cy0 + lane) * G.W + cx1 + 1] - rlo; }
    tot = rcnt;
    for (int r = lane + 32; r < nrows; r += 32) tot += G.cell_start[(cy0 + r) * G.W + cx1 + 1] - G.cell_start[(cy0 + r) * G.W + cx0];
    tot = __reduce_add_sync(FULL, tot);
    if (tot > NB_GIANT) {
        // A long tour edge (a nearest-neighbour tour has a few hundred, a VNS kick makes three): thousands of candidates.
        // The box is not walked by this warp alone: it gets a heap run with one slot per node of the box and is queued;
        // behind the scan's barrier all cooperating warps fill the run row by row (giant_phase).  When a launch is
        // settling there is no such barrier: the node stays stale for the next launch.
        // A node that had a run of at least this size KEEPS it (a long edge usually survives the change of the node's other
        // edge: same box, same size): every slot below tot is rewritten in giant_phase, and until then nobody uses the old
        // records -- a node with a run is only searched again as one of the four nodes of the last move, whose records the
        // streaming warps skip.  (Rounds 1-2 abandoned the run and took a new one: thousands of dead slots per giant search,
        // the main source of the abandoned records a scan streams.)
#if TSP_NB_GIANT_KEEP
        const bool keep = !SETTLE && had && tot <= ocap;
#else
        const bool keep = false;
#endif
        if (had)
            for (int j = (keep ? tot : 0) + lane; j < ocount; j += 32) C.heap[ostart + j].x = -1;
        // fresh: the node's tour edges have just changed (one of the four nodes of the last move); otherwise the header says
        // how often this box has been evaluated without caching since
        const int age = (!fresh && (h0 & 0x3FF) == NB_CF_STALE) ? (h0 >> 10) : 0;
        if (!SETTLE && age < C.giant_age) {                        // evaluate only: queued without a heap run, the node stays stale
            if (keep)
                for (int j = lane; j < tot; j += 32) C.heap[ostart + j].x = -1;
            if (lane == 0) {
                NBSTAT(0, 1); NBSTAT(7, tot); NBSTAT(1, 1);
                if (had) { atomicAdd(C.heap_next + 3, -ocount); atomicAdd(C.heap_next + 4, -1); }
                line[0] = NB_CF_STALE | ((age + 1) << 10); C.cnt8[u] = 1;
                const unsigned long long old = atomicAdd(qctr, (1ull << 40) | (unsigned long long)nrows);
                BigBox bx;
                bx.p = c.p; bx.cy0 = cy0; bx.cx0 = cx0; bx.cx1 = cx1; bx.nrows = nrows; bx.pad = -1;
                bx.base = (long long)(old & PR_UNIT_MASK);
                boxes[old >> 40] = bx;
            }
            return;
        }
        int start = -1;
        if (keep) start = ostart;
        else {
            if (!SETTLE && lane == 0) start = *(volatile int *)(C.heap_next + hp) > C.heap_cap ? -1 : atomicAdd(C.heap_next + hp, tot);
            start = __shfl_sync(FULL, start, 0);
            if (start > C.heap_cap - tot) {                        // no room: what the run covers is void
                for (int j = start + lane; j < C.heap_cap; j += 32) C.heap[j].x = -1;
                start = -1;
            }
        }
        if (lane == 0) {
            NBSTAT(0, 1); NBSTAT(7, tot); NBSTAT(1, 1);
            atomicAdd(C.heap_next + 3, (start >= 0 ? tot : 0) - (had ? ocount : 0));
            if ((start >= 0) != had) atomicAdd(C.heap_next + 4, had ? -1 : 1);
            if (start >= 0) {
                line[0] = NB_CF_HEAP; line[1] = start; line[2] = tot; line[3] = keep ? ocap : tot; C.cnt8[u] = 0;
                const unsigned long long old = atomicAdd(qctr, (1ull << 40) | (unsigned long long)nrows);
                BigBox bx;
                bx.p = c.p; bx.cy0 = cy0; bx.cx0 = cx0; bx.cx1 = cx1; bx.nrows = nrows; bx.pad = start;
                bx.base = (long long)(old & PR_UNIT_MASK);
                boxes[old >> 40] = bx;
            } else { line[0] = NB_HDR_STALE; C.cnt8[u] = 1; }
        }
        if (start >= 0 || SETTLE) return;
        // no room in the heap: this warp walks the box after all, evaluating only (the node stays stale)
        for (int r = 0; r < nrows; ++r) {
            const int lo = G.cell_start[(cy0 + r) * G.W + cx0], hi = G.cell_start[(cy0 + r) * G.W + cx1 + 1];
            if (lane == 0) a.ncand += hi - lo;
            for (int k0 = lo; k0 < hi; k0 += 32) {
                const int k = min(k0 + lane, hi - 1);
                const double2 cw = G.cxy[k];
                const int w = G.cnode[k];
                const int K = M::K(c.u.x, c.u.y, cw.x, cw.y);
                eval_push<M, NE>(V, F, ebuf, a, k0 + lane < hi && K < thr && w != u && w != c.su && w != c.pu, c.p, w, K);
            }
        }
        return;
    }
    // The usual box (a node with two short tour edges) holds a few dozen nodes in two to five cell rows.  Walked row by row
    // it costs one round trip to L2 per row and pass; FLAT: the (at most NB_FLAT) candidates are numbered across the rows
    // (prefix sums of the row sizes, in registers), every lane loads its one or two candidates at once, and what the first
    // pass found stays in registers for the second: the search of one of the four nodes of a move -- the longest dependent
    // chain of a whole-GPU scan -- shrinks from 3 + 3 x rows round trips to 6.
#ifdef TSP_NB_NOFLAT
    const bool flat = false;                                       // (measurements: the row-by-row walk for every box)
#else
    const bool flat = nrows <= 32 && tot <= NB_FLAT;
#endif
    int fw[NB_FLAT / 32], fK[NB_FLAT / 32];
    unsigned fm[NB_FLAT / 32];
    if (flat) {
        int pre = rcnt;                                            // inclusive prefix of the row sizes
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const int v = __shfl_up_sync(FULL, pre, off); if (lane >= off) pre += v; }
        int row[NB_FLAT / 32];
#pragma unroll
        for (int j = 0; j < NB_FLAT / 32; ++j) row[j] = 0;
        for (int r = 0; r + 1 < nrows; ++r) {                      // row of candidate idx: the first whose prefix exceeds it (empty rows skipped)
            const int pr = __shfl_sync(FULL, pre, r);
#pragma unroll
            for (int j = 0; j < NB_FLAT / 32; ++j) row[j] += (lane + 32 * j >= pr) ? 1 : 0;
        }
#pragma unroll
        for (int j = 0; j < NB_FLAT / 32; ++j) {
            const int idx = lane + 32 * j;
            const int k = __shfl_sync(FULL, rlo, row[j]) + idx - (__shfl_sync(FULL, pre, row[j]) - __shfl_sync(FULL, rcnt, row[j]));
            bool pass = false;
            fw[j] = 0; fK[j] = 0;
            if (idx < tot) {
                const double2 cw = G.cxy[k];
                fw[j] = G.cnode[k];
                fK[j] = M::K(c.u.x, c.u.y, cw.x, cw.y);
                pass = fK[j] < thr && fw[j] != u && fw[j] != c.su && fw[j] != c.pu;                  // tour neighbours: refinment.c:55
            }
            fm[j] = __ballot_sync(FULL, pass);
            count += __popc(fm[j]);
        }
    } else {
        tot = 0;
        for (int r = 0; r < nrows; ++r) {
            const int lo = G.cell_start[(cy0 + r) * G.W + cx0], hi = G.cell_start[(cy0 + r) * G.W + cx1 + 1];
            tot += hi - lo;
            for (int k0 = lo; k0 < hi; k0 += 32) {
                const int k = min(k0 + lane, hi - 1);
                const double2 cw = G.cxy[k];
                const int w = G.cnode[k];
                const int K = M::K(c.u.x, c.u.y, cw.x, cw.y);
                const bool pass = k0 + lane < hi && K < thr && w != u && w != c.su && w != c.pu;          // tour neighbours: refinment.c:55
                count += __popc(__ballot_sync(FULL, pass));
            }
        }
    }
    if (lane == 0) { NBSTAT(0, 1); NBSTAT(7, tot); a.ncand += tot; }
    int start = -1, cap = 0;
    if (count > NB_CAP) {
        if (had && count <= ocap) { start = ostart; cap = ocap; }  // the node's own run, again
        else {
            cap = count <= 16 ? 16 : count <= 64 ? 64 : count + (count >> 2);
            if (lane == 0)                                         // a counter beyond the capacity stays where it is (no overflow)
                start = *(volatile int *)(C.heap_next + hp) > C.heap_cap ? -1 : atomicAdd(C.heap_next + hp, cap);
            start = __shfl_sync(FULL, start, 0);
            if (start < 0 || start > C.heap_cap - cap) {           // no room: the node stays stale; what the run covers is void
                if (start >= 0)
                    for (int j = start + lane; j < C.heap_cap; j += 32) C.heap[j].x = -1;
                start = -1;
            }
        }
    }
    if (had && start != ostart)                                    // the old run is abandoned
        for (int j = lane; j < ocount; j += 32) C.heap[ostart + j].x = -1;
    if (lane == 0 && (had || start >= 0)) {
        atomicAdd(C.heap_next + 3, (start >= 0 ? count : 0) - (had ? ocount : 0));                   // live records
        if ((start >= 0) != had) atomicAdd(C.heap_next + 4, had ? -1 : 1);                              // nodes with a run
    }
    if (count == 0) {
        if (lane == 0) { line[0] = 0; C.cnt8[u] = 0; }
        return;
    }
    const int ou = V.o8[u];
    int stored = 0;
    if (flat) {
        int ds[NB_FLAT / 32], df[NB_FLAT / 32], ow[NB_FLAT / 32];
#pragma unroll
        for (int j = 0; j < NB_FLAT / 32; ++j) { ds[j] = 0; df[j] = 0; ow[j] = ((fm[j] >> lane) & 1u) ? V.o8[fw[j]] : 0; }
#pragma unroll
        for (int j = 0; j < NB_FLAT / 32; ++j)
            if ((fm[j] >> lane) & 1u) record_deltas<M>(V, F, c, fw[j], fK[j], ds[j], df[j]);
#pragma unroll
        for (int j = 0; j < NB_FLAT / 32; ++j) {
            if (fm[j] == 0) continue;                              // warp-uniform
            const bool pass = (fm[j] >> lane) & 1u;
            if (pass) {
                const int o = stored + __popc(fm[j] & lt);
                const bool rel = ((ou ^ ow[j]) & 1) != 0;
                const int d0 = rel ? df[j] : ds[j], d1 = rel ? ds[j] : df[j];
                if (count <= NB_CAP) { line[1 + 3 * o] = fw[j]; line[2 + 3 * o] = d0; line[3 + 3 * o] = d1; }
                else if (start >= 0) C.heap[start + o] = make_int4(u, fw[j], d0, d1);
            }
            stored += __popc(fm[j]);
            if (!SETTLE) {
                const bool hit = pass && ds[j] <= a.best;
                if (hit) a.best = ds[j];
                eval_push<M, NE>(V, F, ebuf, a, hit, c.p, fw[j], fK[j]);
            }
        }
    } else
    for (int r = 0; r < nrows; ++r) {
        const int lo = G.cell_start[(cy0 + r) * G.W + cx0], hi = G.cell_start[(cy0 + r) * G.W + cx1 + 1];
        for (int k0 = lo; k0 < hi; k0 += 32) {
            const int k = min(k0 + lane, hi - 1);
            const double2 cw = G.cxy[k];
            const int w = G.cnode[k];
            const int K = M::K(c.u.x, c.u.y, cw.x, cw.y);
            const bool pass = k0 + lane < hi && K < thr && w != u && w != c.su && w != c.pu;
            const unsigned mask = __ballot_sync(FULL, pass);
            if (mask == 0) continue;                               // warp-uniform
            int dsame = 0, dflip = 0;
            if (pass) {
                record_deltas<M>(V, F, c, w, K, dsame, dflip);
                const int o = stored + __popc(mask & lt);
                const bool rel = ((ou ^ V.o8[w]) & 1) != 0;
                const int d0 = rel ? dflip : dsame, d1 = rel ? dsame : dflip;
                if (count <= NB_CAP) { line[1 + 3 * o] = w; line[2 + 3 * o] = d0; line[3 + 3 * o] = d1; }
                else if (start >= 0) C.heap[start + o] = make_int4(u, w, d0, d1);
            }
            stored += __popc(mask);
            if (!SETTLE) {
                const bool hit = pass && dsame <= a.best;
                if (hit) a.best = dsame;
                eval_push<M, NE>(V, F, ebuf, a, hit, c.p, w, K);
            }
        }
    }
    if (start >= 0) {                                              // the rest of the run: what an earlier, longer list left there, or nothing yet
        const int end = start == ostart && had ? max(ocount, count) : cap;
        for (int j = start + count + lane; j < start + end; j += 32) C.heap[j].x = -1;
    }
    if (lane == 0) {
        if (count <= NB_CAP) { line[0] = count; C.cnt8[u] = 1; }
        else if (start >= 0) { line[0] = NB_CF_HEAP; line[1] = start; line[2] = count; line[3] = cap; C.cnt8[u] = 0; }
        else { line[0] = NB_HDR_STALE; C.cnt8[u] = 1; NBSTAT(1, 1); }
    }
}

// the attention entries of a warp: dirty records {u, w, address of the record's w field | AT_DIRTY} -- recomputed in
// place, every lane its own -- and stale nodes {u, -, AT_NODE}: searched again, one at a time, by the whole warp
template <class M, int NE, bool SETTLE>
__device__ __noinline__ ScanAcc attend(TourView V, const SpatialGrid *Gp, PosPatch F, NbCache C, int hp, unsigned long long *qctr, BigBox *boxes,
                                       uint4 *abuf, int acount, ScanAcc a) {
    const SpatialGrid &G = *Gp;
    const int lane = threadIdx.x & 31;
    uint4 *ebuf = abuf + AB_CAP;
    if (lane == 0) NBSTAT(6, 1);
    __syncwarp();
    for (int base = 0; base < acount; base += 32) {
        const bool have = base + lane < acount;
        uint4 e = make_uint4(0, 0, 0, 0);
        if (have) e = abuf[base + lane];
        const unsigned kind = e.z & 3u;
        const bool isrec = have && kind == AT_DIRTY;
        if (__any_sync(FULL, isrec)) {
            bool hit = false;
            int p = 0, K = 0;
            const int u = (int)e.x, w = (int)e.y;
            if (isrec) {
                NBSTAT(3, 1);
                int *rec = reinterpret_cast<int *>(((unsigned long long)e.w << 32) | (unsigned long long)(e.z & ~3u));
                NodeCtx c;
                node_ctx(V, F, u, c);
                p = c.p;
                const double2 wxy = V.pxy[pos_of(V, F, w)];
                K = M::K(c.u.x, c.u.y, wxy.x, wxy.y);
                int dsame, dflip;
                record_deltas<M>(V, F, c, w, K, dsame, dflip);
                const bool rel = ((V.o8[u] ^ V.o8[w]) & 1) != 0;
                rec[1] = rel ? dflip : dsame;
                rec[2] = rel ? dsame : dflip;
                hit = dsame <= a.best;
                if (hit) a.best = dsame;
            }
            if (!SETTLE) eval_push<M, NE>(V, F, ebuf, a, hit, p, w, K);
        }
        unsigned nm = __ballot_sync(FULL, have && kind == AT_NODE);
        while (nm) {
            const int src = __ffs(nm) - 1;
            nm &= nm - 1;
            rebuild_list<M, NE, SETTLE>(V, G, F, C, hp, (int)__shfl_sync(FULL, e.x, src), __shfl_sync(FULL, e.y, src) != 0u, qctr, boxes, ebuf, a);
        }
    }
    __syncwarp();
    return a;
}

// The cached scan: warp `gw` of `nwarps` cooperating warps (a grid, a cluster or one CTA).
//   ep     epoch of the last move (NB_EP_NONE: none): nodes whose o8 byte carries it had their neighbours changed
//   hsnap  number of heap records to stream (the heap counter as it stood between the barriers before this scan)
//   hp     which of the two heap counters is in use
//   qctr, boxes   queue of the giant search boxes met in this scan (giant_phase, behind the caller's barrier)
//   ring   NB_RING ints of shared memory private to the warp
//   s_best the CTA's shared running best (armed to -1 by the caller behind a barrier)
//   sr     STAGED only: the warp's shared-memory stages for the heap stream (below)
template <class M, int NE, bool SETTLE, bool STAGED = false>
__device__ __forceinline__ void cached_scan(const TourView &V, const SpatialGrid &G, const PosPatch &F, int ep, int gw, int nwarps, int hsnap,
                                            int hp, unsigned long long *qctr, BigBox *boxes, uint4 *wbuf, const NbCache &C, int *ring, int *s_best, int &best,
                                            unsigned long long &bestkey, long long &ncand, StageRing *sr = nullptr) {
    ScanAcc a;
    a.bestkey = bestkey; a.ncand = ncand; a.best = best; a.ecount = 0;
    int acount = 0;
    const int n = V.n;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1;
    uint4 *ebuf = wbuf + AB_CAP;
    // the four nodes of the last move: one designated warp each (their searches are the longest dependent chains of a scan).
    // With many cooperating warps (the whole GPU on one descent) a warp's share of the stream is one or two records -- a few
    // dependent loads -- and the search is the longer chain: the four warps then ONLY search, at once, and the stream is
    // dealt out over the others (sgw of snw); with few warps (one CTA per descent) the stream dominates and all warps take part.
#ifdef TSP_PHASE_TIMING
    unsigned long long scan_t_ = 0;
    if (blockIdx.x == 0 && threadIdx.x == 32) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(scan_t_));
#endif
    const bool split = nwarps >= NB_SPLIT_WARPS && F.node0 >= 0;                             // the same for every cooperating warp
    const int sgw = split ? gw - 4 : gw, snw = split ? nwarps - 4 : nwarps;
    if (gw < 4) {
        const int fu = gw == 0 ? F.node0 : gw == 1 ? F.node1 : gw == 2 ? F.node2 : F.node3;
        if (fu >= 0) {
            if (lane == 0) wbuf[0] = make_uint4((unsigned)fu, 1u, AT_NODE, 0u);              // .y = 1: fresh edges
            acount = 1;
        }
    }
    if (!(split && gw < 4)) {
    // ---- the heap: every cooperating thread a share of the records.  What is done with the records of a step (HU per lane,
    // record j of this lane = heap[kidx[j]]) is the same for both ways of fetching them:
    auto heap_step = [&](auto hu_tag, const int4 *r, const int *kidx) {
        constexpr int HU = decltype(hu_tag)::value;
        int ou[HU], ow[HU];
#pragma unroll
        for (int j = 0; j < HU; ++j) {
            ou[j] = r[j].x >= 0 ? V.o8[r[j].x] : 0;
            ow[j] = r[j].x >= 0 ? V.o8[r[j].y] : 0;
        }
        int d[HU];
        unsigned dirty = 0;
        int mine = 0x7FFFFFFF;
#pragma unroll
        for (int j = 0; j < HU; ++j) {
            d[j] = 0x7FFFFFFF;
            bool live = r[j].x >= 0;
            if (live && (ou[j] >> 1) == ep && in_patch(F, r[j].x)) live = false;      // u's own list is being rebuilt by a designated warp
            if (live) {
                if (r[j].z == NB_DIRTY || (ow[j] >> 1) == ep) dirty |= 1u << j;
                else { d[j] = ((ou[j] ^ ow[j]) & 1) ? r[j].w : r[j].z; mine = min(mine, d[j]); }
            }
            a.ncand += live ? 1 : 0;
            NBSTAT(2, 1);
            if (live) NBSTAT(4, 1);
        }
        if (__any_sync(FULL, dirty != 0)) {                        // rare
#pragma unroll
            for (int j = 0; j < HU; ++j) {
                const bool f = (dirty >> j) & 1u;
                const unsigned m = __ballot_sync(FULL, f);
                if (f) {
                    const unsigned long long ra = (unsigned long long)(&C.heap[kidx[j]].y);
                    wbuf[acount + __popc(m & lt)] = make_uint4((unsigned)r[j].x, (unsigned)r[j].y, (unsigned)(ra & 0xFFFFFFFFu) | AT_DIRTY, (unsigned)(ra >> 32));
                }
                acount += __popc(m);
            }
        }
        if (!SETTLE) {
            const int seen = *(volatile int *)s_best;
            const int wm = __reduce_min_sync(FULL, min(mine, min(a.best, seen)));
            a.best = wm;
            if (lane == 0 && wm < seen) atomicMin(s_best, wm);
            if (__any_sync(FULL, mine <= wm)) {
#pragma unroll
                for (int j = 0; j < HU; ++j) {
                    const bool hit = d[j] <= wm;
                    const unsigned m = __ballot_sync(FULL, hit);
                    if (hit) ebuf[a.ecount + __popc(m & lt)] = make_uint4((unsigned)r[j].x, (unsigned)r[j].y, 0u, 3u | EV_NODEIDS);
                    a.ecount += __popc(m);
                }
            }
        }
    };
    if (STAGED) {
        // One CTA (or a small cluster) per descent: the stream IS the scan (pla85900: 5.8 MB of records per scan and descent,
        // 148 descents resident -- HBM traffic).  A warp that loads its records into registers has them in flight only
        // between the loads and their first use, about half the time.  Here the records travel by bulk asynchronous copies
        // (TMA, cp.async.bulk) into a ring of NB_STAGES shared-memory stages per warp, NB_STAGE_RECS records (1 KB) each,
        // completion on one mbarrier per stage: the copy of a stage is re-issued the moment its records are in registers,
        // so NB_STAGES - 1 stages per warp are in flight ALL the time, whatever the warp does with the records.  The heap
        // is dealt out in NB_STAGE_RECS-record blocks, block b to warp b mod snw.
        constexpr int HU = NB_STAGE_RECS / 32;
        const int nblk = (hsnap + NB_STAGE_RECS - 1) / NB_STAGE_RECS;
        const int nmine = sgw < nblk ? (nblk - sgw + snw - 1) / snw : 0;
        auto issue = [&](int i) {                                  // lane 0: block i of this warp into stage i mod NB_STAGES
            const int b = sgw + i * snw, st = i % NB_STAGES;
            const unsigned bytes = 16u * (unsigned)min(NB_STAGE_RECS, hsnap - b * NB_STAGE_RECS);
            mbar_expect_tx(&sr->bar[st], bytes);
            bulk_g2s(sr->buf + st * NB_STAGE_RECS, C.heap + (size_t)b * NB_STAGE_RECS, bytes, &sr->bar[st]);
        };
        if (lane == 0)
            for (int i = 0; i < min(NB_STAGES, nmine); ++i) issue(i);
        for (int i = 0; i < nmine; ++i) {
            if (acount > AB_CAP - 32 * HU) { a = attend<M, NE, SETTLE>(V, &G, F, C, hp, qctr, boxes, wbuf, acount, a); acount = 0; }      // warp-uniform, rare
            if (!SETTLE && a.ecount > EB_CAP - 32 * HU) a = flush_records<M, NE>(V, F, ebuf, a);
            const int st = i % NB_STAGES, k0 = (sgw + i * snw) * NB_STAGE_RECS;
            mbar_wait(&sr->bar[st], (sr->phase >> st) & 1u);
            sr->phase ^= 1u << st;
            int4 r[HU];
            int kidx[HU];
#pragma unroll
            for (int j = 0; j < HU; ++j) {
                kidx[j] = k0 + 32 * j + lane;
                r[j] = sr->buf[st * NB_STAGE_RECS + 32 * j + lane];
                if (kidx[j] >= hsnap) r[j] = make_int4(-1, 0, 0, 0);           // beyond the bytes the copy brought
            }
            __syncwarp();                                          // the stage is in registers: its next block may land
            if (lane == 0 && i + NB_STAGES < nmine) issue(i + NB_STAGES);
            heap_step(std::integral_constant<int, HU>{}, r, kidx);
        }
    } else {
        // NB_HU loads in flight per lane (a warp that waits for one 16-byte record at a time is bound by the memory latency)
        for (int k0 = sgw * 32; k0 < hsnap; k0 += snw * 32 * NB_HU) {
            if (acount > AB_CAP - 32 * NB_HU) { a = attend<M, NE, SETTLE>(V, &G, F, C, hp, qctr, boxes, wbuf, acount, a); acount = 0; }      // warp-uniform, rare
            if (!SETTLE && a.ecount > EB_CAP - 32 * NB_HU) a = flush_records<M, NE>(V, F, ebuf, a);
            int4 r[NB_HU];
            int kidx[NB_HU];
#pragma unroll
            for (int j = 0; j < NB_HU; ++j) {
                kidx[j] = k0 + j * snw * 32 + lane;
                r[j] = make_int4(-1, 0, 0, 0);
                if (kidx[j] < hsnap) r[j] = C.heap[kidx[j]];
            }
            heap_step(std::integral_constant<int, NB_HU>{}, r, kidx);
        }
    }
    SCAN_MARK(0)
    // ---- the lines: count bytes first, non-empty lines 32 per step
    // many warps, few nodes each (the whole GPU on one descent): one 32-node range per warp, so that every warp gets some;
    // otherwise blocks of 128 nodes, whose count bytes are read four per lane, dealt out round-robin (TSP_NB_LINE_RANGES: one
    // contiguous range per warp, as before -- node ids follow the geography in most files and the nodes with records are
    // not spread evenly, so contiguous ranges leave the warps of a CTA with unequal work before the scan's barrier)
    const int per = (n + snw - 1) / snw;
    const bool wide = per > 64;
    const int chunk = wide ? (per + 127) & ~127 : (per + 31) & ~31;
#ifdef TSP_NB_LINE_RANGES
    const int lo = min(n, sgw * chunk), hi = min(n, lo + chunk), wstep = 128;
#else
    const int lo = min(n, wide ? sgw * 128 : sgw * chunk), hi = wide ? n : min(n, lo + chunk), wstep = snw * 128;
#endif
    int base = lo, head = 0, tail = 0;
    for (;;) {
        while (base < hi && tail - head < 32) {                    // refill
            if (wide) {                                            // (cnt8 is allocated in whole words; base is a multiple of 128)
                const int u4 = base + 4 * lane;
                const unsigned c4 = u4 < hi ? *reinterpret_cast<const unsigned *>(C.cnt8 + u4) : 0u;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const bool want = ((c4 >> (8 * j)) & 0xFFu) != 0 && u4 + j < hi;
                    const unsigned m = __ballot_sync(FULL, want);
                    if (want) ring[(tail + __popc(m & lt)) & (NB_RING - 1)] = u4 + j;
                    tail += __popc(m);
                }
                base += wstep;
            } else {
                const int u1 = base + lane;
                const bool want = u1 < hi && C.cnt8[u1] != 0;
                const unsigned m = __ballot_sync(FULL, want);
                if (want) ring[(tail + __popc(m & lt)) & (NB_RING - 1)] = u1;
                tail += __popc(m);
                base += 32;
            }
        }
        const int step = min(32, tail - head);
        if (step == 0) break;
        __syncwarp();
        bool have = lane < step;
        const int u = have ? ring[(head + lane) & (NB_RING - 1)] : 0;
        head += step;
        __syncwarp();
        if (acount > AB_CAP - 32 * (NB_CAP + 1)) { a = attend<M, NE, SETTLE>(V, &G, F, C, hp, qctr, boxes, wbuf, acount, a); acount = 0; }
        if (!SETTLE && a.ecount > EB_CAP - 32 * NB_CAP) a = flush_records<M, NE>(V, F, ebuf, a);
        int4 q0 = make_int4(0, 0, 0, 0), q1 = q0, q2 = q0, q3 = q0;
        int ou = 0;
        int *line = C.lines + (size_t)u * NB_WORDS;
        if (have) {
            const int4 *L = reinterpret_cast<const int4 *>(line);
            q0 = L[0]; q1 = L[1]; q2 = L[2]; q3 = L[3];
            ou = V.o8[u];
            if ((ou >> 1) == ep && in_patch(F, u)) have = false;   // one of the four: a designated warp searches it
        }
        const int cf = q0.x & 0x3FF;
        const bool stale = have && cf == NB_CF_STALE;
        const int inl = (have && cf <= NB_CAP) ? cf : 0;
        a.ncand += inl;
        {
            const unsigned m = __ballot_sync(FULL, stale);
            if (stale) wbuf[acount + __popc(m & lt)] = make_uint4((unsigned)u, 0u, AT_NODE, 0u);
            acount += __popc(m);
        }
        const int rw[NB_CAP] = {q0.y, q1.x, q1.w, q2.z, q3.y};
        const int r0[NB_CAP] = {q0.z, q1.y, q2.x, q2.w, q3.z};
        const int r1[NB_CAP] = {q0.w, q1.z, q2.y, q3.x, q3.w};
        int d[NB_CAP];
        unsigned dirty = 0;
        int mine = 0x7FFFFFFF;
#pragma unroll
        for (int j = 0; j < NB_CAP; ++j) {
            d[j] = 0x7FFFFFFF;
            if (j < inl) {
                const int ow = V.o8[rw[j]];
                if (r0[j] == NB_DIRTY || (ow >> 1) == ep) dirty |= 1u << j;
                else { d[j] = ((ou ^ ow) & 1) ? r1[j] : r0[j]; mine = min(mine, d[j]); }
            }
        }
        if (__any_sync(FULL, dirty != 0)) {                        // rare
#pragma unroll
            for (int j = 0; j < NB_CAP; ++j) {
                const bool f = (dirty >> j) & 1u;
                const unsigned m = __ballot_sync(FULL, f);
                if (f) {
                    const unsigned long long ra = (unsigned long long)(line + 1 + 3 * j);
                    wbuf[acount + __popc(m & lt)] = make_uint4((unsigned)u, (unsigned)rw[j], (unsigned)(ra & 0xFFFFFFFFu) | AT_DIRTY, (unsigned)(ra >> 32));
                }
                acount += __popc(m);
            }
        }
        if (!SETTLE) {
            // the best anywhere in the CTA so far; only records that tie or beat it are worth an exact evaluation
            const int seen = *(volatile int *)s_best;
            const int wm = __reduce_min_sync(FULL, min(mine, min(a.best, seen)));
            a.best = wm;
            if (lane == 0 && wm < seen) atomicMin(s_best, wm);
            if (__any_sync(FULL, mine <= wm)) {
#pragma unroll
                for (int j = 0; j < NB_CAP; ++j) {
                    const bool hit = d[j] <= wm;
                    const unsigned m = __ballot_sync(FULL, hit);
                    if (hit) ebuf[a.ecount + __popc(m & lt)] = make_uint4((unsigned)u, (unsigned)rw[j], 0u, 3u | EV_NODEIDS);
                    a.ecount += __popc(m);
                }
            }
        }
    }
    SCAN_MARK(1)
    }   // !(split && gw < 4)
    if (acount > 0) a = attend<M, NE, SETTLE>(V, &G, F, C, hp, qctr, boxes, wbuf, acount, a);
    if (!SETTLE && a.ecount > 0) a = flush_records<M, NE>(V, F, ebuf, a);
    SCAN_MARK(2)
#ifdef TSP_PHASE_TIMING
    if (blockIdx.x == 0 && threadIdx.x == 32) atomicAdd(&g_scan_ns[3], 1ull);
#endif
    if (STAGED) heap_writes_to_async_proxy();
    best = a.best; bestkey = a.bestkey; ncand = a.ncand;
}

// Behind the scan's barrier, when giant boxes were queued (q = the packed counter): their cell rows are split evenly
// over the `nwarps` cooperating warps (this one is `gw`), as in pruned_phase2.  Candidate number i of a box (row by
// row) owns slot i of the box's heap run: a record where the candidate passes, an abandoned slot where it does not --
// no counting, no atomics; the records that can win now go to the exact evaluation.
template <class M, int NE>
__device__ __noinline__ void giant_phase(TourView V, const SpatialGrid *Gp, PosPatch F, NbCache C, long long gw, long long nwarps,
                                         unsigned long long q, const BigBox *boxes, uint4 *wbuf, int *best_io, unsigned long long *key_io,
                                         long long *ncand_io) {
    const SpatialGrid &G = *Gp;
    const int lane = threadIdx.x & 31;
    const long long nunits = (long long)(q & PR_UNIT_MASK);
    const int nboxes = (int)(q >> 40);
    const long long per = (nunits + nwarps - 1) / nwarps;
    const long long g0 = gw * per, g1 = min(nunits, g0 + per);
    if (g0 >= g1) return;
    ScanAcc a;
    a.bestkey = *key_io; a.ncand = *ncand_io; a.best = *best_io; a.ecount = 0;
    uint4 *ebuf = wbuf + AB_CAP;
    int lo_b = 0, hi_b = nboxes;                                  // last box with base <= g0: 33-ary search
    while (hi_b - lo_b > 1) {
        const int step = (hi_b - lo_b + 32) / 33;
        const int idx = lo_b + (lane + 1) * step;
        const bool ok = idx < hi_b && boxes[idx].base <= g0;
        const int cnt = __popc(__ballot_sync(FULL, ok));
        const int nlo = lo_b + cnt * step;
        hi_b = min(hi_b, lo_b + (cnt + 1) * step);
        lo_b = nlo;
    }
    int bi = lo_b;
    BigBox b = boxes[bi];
    NodeCtx c;
    int u = 0, ou = 0, thr = 0, rowoff = 0;
    bool fresh = true;
    for (long long g = g0; g < g1; ++g) {
        while (g >= b.base + b.nrows) { b = boxes[++bi]; fresh = true; }
        const int row = (int)(g - b.base);
        if (fresh) {                                               // a new box: its node, and the slots the rows before this one own
            u = V.tid[b.p];
            node_ctx(V, F, u, c);
            ou = V.o8[u];
            thr = max(c.dnp, c.dnm);
            int off = 0;
            for (int r = lane; r < row; r += 32) off += G.cell_start[(b.cy0 + r) * G.W + b.cx1 + 1] - G.cell_start[(b.cy0 + r) * G.W + b.cx0];
            rowoff = __reduce_add_sync(FULL, off);
            fresh = false;
        }
        const int cy = b.cy0 + row;
        const int lo = G.cell_start[cy * G.W + b.cx0], hi = G.cell_start[cy * G.W + b.cx1 + 1];
        if (lane == 0) a.ncand += hi - lo;
        for (int k0 = lo; k0 < hi; k0 += 32) {
            const int k = k0 + lane;
            bool pass = false;
            int w = 0, K = 0, dsame = 0, dflip = 0;
            if (k < hi) {
                const double2 cw = G.cxy[k];
                w = G.cnode[k];
                K = M::K(c.u.x, c.u.y, cw.x, cw.y);
                pass = K < thr && w != u && w != c.su && w != c.pu;
                if (b.pad >= 0) {
                    int4 rec = make_int4(-1, 0, 0, 0);
                    if (pass) {
                        record_deltas<M>(V, F, c, w, K, dsame, dflip);
                        const bool rel = ((ou ^ V.o8[w]) & 1) != 0;
                        rec = make_int4(u, w, rel ? dflip : dsame, rel ? dsame : dflip);
                    }
                    C.heap[b.pad + rowoff + (k - lo)] = rec;
                }
            }
            bool hit = pass;                                       // a box without a run (giant_age): every passing candidate is evaluated exactly
            if (b.pad >= 0) {
                hit = pass && dsame <= a.best;
                if (hit) a.best = dsame;
            }
            eval_push<M, NE>(V, F, ebuf, a, hit, c.p, w, K);
        }
        rowoff += hi - lo;
    }
    if (a.ecount > 0) a = flush_records<M, NE>(V, F, ebuf, a);
    heap_writes_to_async_proxy();
    *best_io = a.best; *key_io = a.bestkey; *ncand_io = a.ncand;
}

// between the last two barriers of an iteration, every cooperating thread (t of nthreads): the heap fill the next scan
// streams -- or a fresh heap.  Abandoned records are streamed like live ones.  Reading them costs the less the more
// threads share a descent, searching a node again does not get cheaper: once the abandoned records outnumber the live
// ones (one CTA per descent; proportionally more for a cluster or the whole GPU), or the heap is three quarters full, the
// nodes whose lists live in the heap are declared stale and the heap starts over at the other, zeroed, counter.  Lists
// held in the lines are not affected.
// hn, live: the counters heap_next[hp] and heap_next[3] as they stand between those barriers (read by the caller: one thread
// per CTA and a shared-memory broadcast where a whole grid cooperates)
__device__ __forceinline__ int heap_checkpoint(const NbCache &C, int n, int t, int nthreads, int &hp, int hn, int live) {
    const long long teams = max(1, nthreads >> 10);
    if (hn <= C.heap_cap / 4 * 3 && ((long long)hn - live) * 100 <= (teams * live + (n >> 1)) * C.restart_pct) return min(hn, C.heap_cap);
    for (int u = t; u < n; u += nthreads) {
        int *h = C.lines + (size_t)u * NB_WORDS;
        if ((*h & 0x3FF) == NB_CF_HEAP) { *h = NB_HDR_STALE; C.cnt8[u] = 1; }
    }
    hp ^= 1;
    if (t == 0) { C.heap_next[hp] = 0; C.heap_next[2] = hp; C.heap_next[3] = 0; C.heap_next[4] = 0; }      // nobody looks at these before the next barrier
    return 0;
}

// Phase 2 (behind a barrier): the queued boxes, cut into (box, cell row) units -- one contiguous
// range of the cell-sorted node array each -- split evenly over the `nwarps` cooperating warps;
// this warp is number `gw`.  q = the packed counter after phase 1.
template <class M, int NE, bool TABU>
__device__ __forceinline__ void pruned_phase2(const TourView &V, const SpatialGrid &G, const PosPatch &F, long long gw,
                                              long long nwarps, unsigned long long q, const BigBox *boxes, uint4 *wbuf,
                                              int &best, unsigned long long &bestkey, long long &ncand, int h) {
    const int lane = threadIdx.x & 31;
    const long long nunits = (long long)(q & PR_UNIT_MASK);
    const int nboxes = (int)(q >> 40);
    if (nunits <= 0) return;
    const long long per = (nunits + nwarps - 1) / nwarps;
    const long long g0 = gw * per, g1 = min(nunits, g0 + per);
    if (g0 >= g1) return;
    int wcount = 0;
    // last box with base <= g0: 33-ary search, one probe per lane and step
    int lo_b = 0, hi_b = nboxes;
    while (hi_b - lo_b > 1) {
        const int step = (hi_b - lo_b + 32) / 33;
        const int idx = lo_b + (lane + 1) * step;
        const bool ok = idx < hi_b && boxes[idx].base <= g0;
        const int cnt = __popc(__ballot_sync(FULL, ok));           // bases ascend: the ok lanes are a prefix
        const int nlo = lo_b + cnt * step;
        hi_b = min(hi_b, lo_b + (cnt + 1) * step);
        lo_b = nlo;
    }
    int bi = lo_b;
    BigBox b = boxes[bi];
    Probe P;
    probe_load<TABU>(V, b.p, P, h);
    for (long long g = g0; g < g1; ++g) {
        while (g >= b.base + b.nrows) { b = boxes[++bi]; probe_load<TABU>(V, b.p, P, h); }
        const int cy = b.cy0 + (int)(g - b.base);
        const int lo = G.cell_start[cy * G.W + b.cx0], hi = G.cell_start[cy * G.W + b.cx1 + 1];
        if (lane == 0) ncand += hi - lo;
        for (int k0 = lo; k0 < hi; k0 += 32 * PR_NC2) {           // warp-uniform trips over the row
            int kk[PR_NC2];
            bool valid[PR_NC2];
#pragma unroll
            for (int j = 0; j < PR_NC2; ++j) {
                const int k = k0 + lane + 32 * j;
                valid[j] = k < hi;
                kk[j] = min(k, hi - 1);
            }
            collect_trip<M, PR_NC2, NE, TABU>(V, G, F, b.p, P.ux, P.uy, P.dnp, P.dnm, kk, valid, wbuf, wcount, best, bestkey);
        }
    }
    __syncwarp();
    pairs_eval<M, NE, TABU>(V, F, wbuf, wcount, best, bestkey);
}

// ---- K7: whole-GPU persistent pruned descent (cooperative launch, one CTA per SM) ----------------
// Two grid barriers per scan + move: scan (stream the delta cache, four searches, exact evaluation of the few records
// that can win) -> argmin -> barrier -> move -> barrier.
template <class M, int BLK>
__global__ void __launch_bounds__(BLK, 1)
descend_pruned_kernel(TourView V, SpatialGrid G, unsigned long long *key, DescState *st, MoveRec *log, long long log_cap,
                      long long max_moves, unsigned long long *qctr, BigBox *boxes, uint4 *wbuf_all, NbCache C, long long moves0) {
    constexpr int WARPS = BLK / 32;
    constexpr int NE = BLK > 512 ? 2 : 4;                // records in flight per lane while evaluating (64 / 128 registers)
    cg::grid_group grid = cg::this_grid();
    __shared__ unsigned long long s_key;
    __shared__ int s_best;
    __shared__ int s_ring[WARPS][NB_RING];
    // What every thread of the grid needs to know after a barrier -- the winning key, the giant-box counter, the heap
    // counters -- is read by ONE thread per CTA and handed on through shared memory: 4 736 warps loading the same word
    // (volatile, so past L1) queue up at one L2 slice for microseconds, and there were three such rounds, one after the
    // other, per move (profiles/r02_pruned.md).
    __shared__ unsigned long long s_bq, s_bk;
    __shared__ int s_bhn, s_blive;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gtid = blockIdx.x * BLK + threadIdx.x;
    const int gsize = gridDim.x * BLK;
    // consecutive global warp numbers go to different SMs: the warps that own nodes (and the four that search) spread out
    const int gw = warp * gridDim.x + blockIdx.x;
    const int Wtot = gridDim.x * WARPS;
    uint4 *wbuf = wbuf_all + ((size_t)blockIdx.x * WARPS + warp) * PR_WBUF;
    PosPatch F = POSPATCH_NONE;
    bool optimum = false;
    int ep = NB_EP_NONE;
    // taken where nobody changes them: before the first scan of the launch (the barrier below), later between the last
    // two barriers of an iteration
    int hp = *(volatile int *)(C.heap_next + 2);
    int hsnap = min(*(volatile int *)(C.heap_next + hp), C.heap_cap);
    grid.sync();                                         // moves0: moves this cache has seen before the launch (epoch numbering)
#ifdef TSP_PHASE_TIMING
    unsigned long long tprev = 0;
    if (gtid == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tprev));
#define PHASE_MARK(slot)                                                       \
    if (gtid == 0) {                                                           \
        unsigned long long tn_;                                                \
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tn_));                 \
        st->phase_ns[slot] += (long long)(tn_ - tprev);                        \
        tprev = tn_;                                                           \
    }
#else
#define PHASE_MARK(slot)
#endif
    for (long long it = 0; it < max_moves; ++it) {
        if (threadIdx.x == 0) { s_key = KEY_EMPTY; s_best = -1; }
        __syncthreads();
        int best = -1;
        unsigned long long bestkey = KEY_EMPTY;
        long long ncand = 0;
#ifdef TSP_PHASE_TIMING
        unsigned long long tw0; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tw0));
#endif
        cached_scan<M, NE, false>(V, G, F, ep, gw, Wtot, hsnap, hp, qctr, boxes, wbuf, C, s_ring[warp], &s_best, best, bestkey, ncand);
#ifdef TSP_PHASE_TIMING
        {   // slowest and average warp of the scan (phase_ns[3] = max over scans and warps)
            unsigned long long tw1; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tw1));
            if (lane == 0) { atomicMax((unsigned long long *)&st->phase_ns[3], tw1 - tw0); atomicAdd((unsigned long long *)&st->phase_sum, tw1 - tw0); }
        }
#endif
        bestkey = warp_min_key(bestkey);
        for (int off = 16; off > 0; off >>= 1) ncand += __shfl_down_sync(FULL, ncand, off);
        if (lane == 0) {
            if (bestkey != KEY_EMPTY) atomicMin(&s_key, bestkey);
            if (ncand) atomicAdd((unsigned long long *)&st->candidates, (unsigned long long)ncand);
        }
        __syncthreads();
        unsigned long long *kslot = key + (it & 1);
        if (threadIdx.x == 0 && s_key != KEY_EMPTY) atomicMin(kslot, s_key);
        grid.sync();
        PHASE_MARK(0)
        if (threadIdx.x == 0) {
            s_bq = *(volatile unsigned long long *)qctr; s_bk = *(volatile unsigned long long *)kslot;
            s_bhn = *(volatile int *)(C.heap_next + hp); s_blive = *(volatile int *)(C.heap_next + 3);
        }
        __syncthreads();
        unsigned long long kbest = s_bk;
        int hn = s_bhn, live = s_blive;
        {   // giant search boxes met in this scan (rare): all warps fill their heap runs now, one more barrier
            const unsigned long long q = s_bq;
            if (q >> 40) {
                bestkey = KEY_EMPTY;
                ncand = 0;
                giant_phase<M, NE>(V, &G, F, C, gw, Wtot, q, boxes, wbuf, &best, &bestkey, &ncand);
                bestkey = warp_min_key(bestkey);
                for (int off = 16; off > 0; off >>= 1) ncand += __shfl_down_sync(FULL, ncand, off);
                if (lane == 0) {
                    if (bestkey != KEY_EMPTY) atomicMin(&s_key, bestkey);
                    if (ncand) atomicAdd((unsigned long long *)&st->candidates, (unsigned long long)ncand);
                }
                __syncthreads();
                if (threadIdx.x == 0 && s_key != KEY_EMPTY) atomicMin(kslot, s_key);
                grid.sync();
                if (gtid == 0) *qctr = 0;
                if (threadIdx.x == 0) {                          // (everybody has read the first values: barriers in between)
                    s_bk = *(volatile unsigned long long *)kslot;
                    s_bhn = *(volatile int *)(C.heap_next + hp); s_blive = *(volatile int *)(C.heap_next + 3);
                }
                __syncthreads();
                kbest = s_bk; hn = s_bhn; live = s_blive;
                PHASE_MARK(1)
            }
        }
        if (kbest == KEY_EMPTY) {
            if (gtid == 0) st->status = 1;
            optimum = true;
            break;
        }
        hsnap = heap_checkpoint(C, V.n, gtid, gsize, hp, hn, live);
        ep = (int)((moves0 + it) % NB_EPOCHS);
        const int a = key_a(kbest), b = key_b(kbest), delta = key_delta(kbest);
        int seg_s, seg_e;
        apply_move<M, true>(V, a, b, gtid, gsize, false, seg_s, seg_e, ep);
        if (gtid == 96) {                                // the bookkeeping: not the thread that carries the move's own chains (apply_move)
            key[(it + 1) & 1] = KEY_EMPTY;
            const long long lc = st->log_count;
            if (lc < log_cap) { log[lc].a = a; log[lc].b = b; log[lc].delta = delta; log[lc].reserved = 0; }
            st->log_count = lc + 1;
            st->moves += 1;
            st->cost += delta;
            st->last_a = a; st->last_b = b; st->last_delta = delta;
        }
#ifdef TSP_PHASE_TIMING
        if (gtid == 0) {                                 // slot 4: nodes in the reversed segments; slot 5: CTA 0's own share of the move phase (ns)
            unsigned long long tn_;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tn_));
            int L_ = seg_e - seg_s + 1; if (L_ <= 0) L_ += V.n;
            st->phase_ns[4] += L_;
            st->phase_ns[5] += (long long)(tn_ - tprev);
        }
#endif
        grid.sync();
        PHASE_MARK(2)
        F.node0 = b; F.pos0 = seg_s; F.node1 = V.tid[seg_e]; F.pos1 = seg_e;
        F.node2 = a; F.node3 = V.tid[seg_e + 1];                               // with node0/node1: the four nodes whose tour edges changed
        if (gtid == 0) { V.pos[F.node0] = F.pos0; V.pos[F.node1] = F.pos1; }   // deferred end-node positions
    }
    // move cap reached: no scan has followed the last move in this launch, and the next launch knows nothing of F --
    // bring the cache up to date now (the four lists, the records pointing at those four nodes), without candidates
    if (!optimum && F.node0 >= 0) {
        int best = -1;
        unsigned long long bestkey = KEY_EMPTY;
        long long ncand = 0;
        cached_scan<M, NE, true>(V, G, F, ep, gw, Wtot, hsnap, hp, qctr, boxes, wbuf, C, s_ring[warp], nullptr, best, bestkey, ncand);
    }
}

// ---- K5c: device-resident tabu loop over the pruned scan -------------------------------------------------
// tabu_best_move wants the best ADMISSIBLE pair even if it worsens the tour (metaheuristic.c:189,216), so "only
// improving pairs can win" does not hold.  But for any bound B every pair with delta < B has
// D(p,q) - dn[p] < h or D(p+1,q+1) - dn[q] < h with h = ceil(B/2): the same two radius searches with radii
// enlarged by h enumerate every pair below B.  An iteration searches with a bound guessed from the previous move;
// if an admissible pair below the bound is found, the argmin over the enumerated pairs is the exact argmin (ties
// included); otherwise the bound is multiplied by 4 and the scan repeated, and after three failures the kernel
// returns (out->done iterations completed): the host continues with the full-scan loop (K5b), which is the right
// tool once most nodes are tabu.  Tabu edges start no search and are rejected as partners (dns flags).  Everything
// else -- flags, stamps, cost trace, incumbent copy -- is K5b's.
template <class M, int BLK, int L>
__global__ void __launch_bounds__(BLK, 1)
tabu_pruned_kernel(TourView V, SpatialGrid G, unsigned long long *key, TabuCtl T, TabuOut *out, long long iters,
                   unsigned long long *qctr, BigBox *boxes, uint4 *wbuf_all, int bound0) {
    constexpr int WARPS = BLK / 32;
    constexpr int NC1 = BLK > 512 ? PR_NC_BIG : PR_NC;
    constexpr int NE = BLK > 512 ? 2 : 4;
    cg::grid_group grid = cg::this_grid();
    __shared__ unsigned long long s_key, s_bq, s_bk;      // s_bq, s_bk: one reader per CTA of the words the whole grid needs (see K7)
    __shared__ long long s_cost, s_bestcost, s_moves;
    __shared__ int s_bestiter, s_lastdelta;
    const int n = V.n;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gtid = blockIdx.x * BLK + threadIdx.x;
    const int gsize = gridDim.x * BLK;
    const long long gw = (long long)blockIdx.x * WARPS + warp;
    const long long Wtot = (long long)gridDim.x * WARPS;
    uint4 *wbuf = wbuf_all + (size_t)gw * PR_WBUF;
    PosPatch F = POSPATCH_NONE;
    if (threadIdx.x == 0) { s_cost = T.cost; s_bestcost = T.best_cost; s_moves = 0; s_bestiter = -1; s_lastdelta = 0; }
    int bound = bound0;
    unsigned slot = 0;
    long long k = 0;
    for (; k < iters; ++k) {
        const int iter = T.first_iter + (int)k;
        const int tenure = T.tenure[k];
        for (int p = gtid; p <= n; p += gsize) {       // dns flags, metaheuristic.c:208, :416-418
            const int u = V.tid[p], v = V.tid[p == n ? 1 : p + 1];
            const int tu = T.tabu_list[u], tv = T.tabu_list[v];
            const bool f = (iter - tu < tenure && tu != -1) || (iter - tv < tenure && tv != -1);
            V.dns[p] = V.dn[p] - (f ? BIGPEN : 0);
        }
        grid.sync();
        unsigned long long kbest = KEY_EMPTY;
        int B = bound;
        for (int attempt = 0; attempt < 3; ++attempt) {
            if (threadIdx.x == 0) s_key = KEY_EMPTY;
            __syncthreads();
            int best = B - 1;                           // pairs with delta < B
            unsigned long long bestkey = KEY_EMPTY;
            long long ncand = 0;
            const int h = (B + 1) / 2;
            pruned_phase1<M, NC1, NE, L, true>(V, G, F, gtid, gsize, qctr, boxes, wbuf, best, bestkey, ncand, h);
            grid.sync();
            if (threadIdx.x == 0) s_bq = *(volatile unsigned long long *)qctr;
            __syncthreads();
            pruned_phase2<M, NE, true>(V, G, F, gw, Wtot, s_bq, boxes, wbuf, best, bestkey, ncand, h);
            bestkey = warp_min_key(bestkey);
            if (lane == 0 && bestkey != KEY_EMPTY) atomicMin(&s_key, bestkey);
            __syncthreads();
            if (threadIdx.x == 0 && s_key != KEY_EMPTY) atomicMin(key + slot, s_key);
            grid.sync();
            if (threadIdx.x == 0) s_bk = *(volatile unsigned long long *)(key + slot);
            __syncthreads();
            kbest = s_bk;
            if (gtid == 0) { key[slot ^ 1] = KEY_EMPTY; *qctr = 0; }
            slot ^= 1;
            if (kbest != KEY_EMPTY) break;
            B = B < (1 << 26) ? B * 4 : (1 << 28);      // deltas stay below 2^29 (packed key)
            grid.sync();                                // the queue counter is reset before anybody enqueues again
        }
        if (kbest == KEY_EMPTY) break;                  // bound exhausted: the host goes on with the full scan
        const int a = key_a(kbest), b = key_b(kbest), delta = key_delta(kbest);
        if (gtid == 0) {                                // stamp a, b and their successors BEFORE the move (:235-238)
            const int i = V.pos[a], j = V.pos[b];
            const int sa = V.tid[i + 1], sb = V.tid[j + 1];
            T.tabu_list[a] = iter; T.tabu_list[b] = iter; T.tabu_list[sa] = iter; T.tabu_list[sb] = iter;
        }
        int seg_s, seg_e;
        apply_move<M>(V, a, b, gtid, gsize, false, seg_s, seg_e);
        if (threadIdx.x == 0) { s_cost += delta; s_moves += 1; s_lastdelta = delta; }
        if (gtid == 0) T.trace[k] = s_cost;             // updated just above by this very thread
        grid.sync();
        F.node0 = b; F.pos0 = seg_s; F.node1 = V.tid[seg_e]; F.pos1 = seg_e;
        if (gtid == 0) { V.pos[F.node0] = F.pos0; V.pos[F.node1] = F.pos1; }
        const bool improved = s_cost < s_bestcost;      // tsp_update_best_solution, strict '<' (tsp.c:671)
        if (improved)
            for (int p = gtid; p <= n; p += gsize) T.best_tid[p] = V.tid[p];
        __syncthreads();
        if (improved && threadIdx.x == 0) { s_bestcost = s_cost; s_bestiter = iter; }
        bound = delta > 0 ? min(2 * delta + 64, 1 << 26) : 64;
    }
    __syncthreads();
    if (gtid == 0) {
        out->cost = s_cost; out->best_cost = s_bestcost; out->moves = s_moves; out->best_iter = s_bestiter;
        out->done = k; out->last_delta = s_lastdelta;
    }
}

#ifndef TSP_K8_O8_SMEM
#define TSP_K8_O8_SMEM 1
#endif
constexpr bool K8_O8_SMEM = TSP_K8_O8_SMEM != 0;
// ---- K8: one CTA per descent, pruned scan, tours in CTA-private global-memory slots -------------------
// Multi-start work (h_greedy_2opt, heuristics.c:74-116): a pruned scan is a few thousand record
// evaluations, far too little to fill 148 SMs, so the SMs run independent descents instead.  CTAs pull
// start indices from an atomic counter in ascending order; the tour of the current start and its
// delta cache live in the CTA's own slot of global memory (any n; streamed from HBM with 148 descents resident), every barrier is
// a __syncthreads.  Each CTA keeps its best tour under the reference's incumbent rule (strict '<' in
// ascending start order, tsp.c:671); the host takes the (cost, start)-minimum.
struct MultiSlots {
    double2 *pxy;      // [slots][n+1]
    int *tid, *dn;     // [slots][n+1]
    int *pos;          // [slots][n]
    BigBox *boxes;     // [slots][n]  queue of giant search boxes
    int *nb;           // [slots][n * NB_WORDS]  delta cache of the slot's current descent: node lines,
    int4 *nb_heap;     // [slots][heap_cap]                                                    record heap,
    int heap_cap, restart_pct;
    int *nb_next;      // [slots][8]                                                           its counters,
    unsigned char *o8; // [slots][n]                                                           orientation / epoch bytes,
    unsigned char *cnt8; // [slots][nb_cnt8_stride(n)]                                         count bytes
    int *best_path;    // [slots][n]  successor array of the CTA's best tour
    long long *best_cost;   // [slots]
    int *best_start;        // [slots]
};

// per-start outcome of a batch (status_out): a start is handed out at most once, in ascending order
enum { MS_NOT_STARTED = 0, MS_COMPLETE = 1, MS_CUT = 2 };

// CL > 1: a thread-block CLUSTER of CL CTAs (CL SMs) shares one descent -- for batches with fewer starts than SMs,
// where one CTA per descent would leave SMs idle while every descent still takes its full single-SM time.  The
// tour stays in the cluster's global-memory slot; the few scalars of a descent (cost, current start, stop flag) live in
// the shared memory of the cluster's first CTA and are reached by the others through distributed
// shared memory; every barrier is a cluster barrier (barrier.cluster, release/acquire at cluster scope), which also
// orders the team's global-memory accesses to the tour (scripts/cluster_probe.cu: no stale read without extra fences).
template <int CL>
__device__ __forceinline__ void team_sync() {
    if (CL == 1) __syncthreads();
    else cg::this_cluster().sync();
}

// stop (may be NULL): a device word the host raises when the caller's time limit has passed.  It is looked at where the
// reference looks at its clock: before a start is taken (heuristics.c:83-90) and before every scan of a descent
// (refinment.c:17-24); a descent that is cut keeps the tour it has reached, which still competes for the incumbent
// exactly as ref_2opt's update_incumbent does after DEADLINE_EXCEEDED (refinment.c:29-34).
template <class M, int BLK, int CL>
__global__ void __launch_bounds__(BLK, 1)
multi_pruned_kernel(const double2 *__restrict__ xy, int n, SpatialGrid G, const int *__restrict__ order_all, int first_start,
                    int stride, int nstarts, int *next_start, MultiSlots S, long long *final_cost, long long *moves_out,
                    int *status_out, unsigned long long *cand_total, uint4 *wbuf_all, const int *stop) {
    constexpr int WARPS = BLK / 32;
    constexpr int NE = BLK > 512 ? 2 : 4;
    constexpr int TEAM = BLK * CL;                      // threads sharing a descent
    __shared__ unsigned long long s_key_[2], s_q_;
    __shared__ long long s_cost_, s_bestcost_;
    __shared__ int s_cur_, s_stop_, s_best;
    __shared__ int s_ring[WARPS][NB_RING];
    // dynamic shared memory: the stages of the heap stream and their mbarriers (cached_scan, STAGED) -- a team of fewer than
    // NB_SPLIT_WARPS warps streams through them; the 8-SM cluster (256 warps, a few records per lane) loads directly
#if TSP_NB_STAGE
    constexpr bool STG = WARPS * CL < NB_SPLIT_WARPS;
#else
    constexpr bool STG = false;                         // default: every team loads its records directly (measured faster, see StageRing)
#endif
    extern __shared__ __align__(128) unsigned char k8_dyn_smem[];
    StageRing sr;
    sr.buf = reinterpret_cast<int4 *>(k8_dyn_smem) + (size_t)(threadIdx.x >> 5) * NB_STAGES * NB_STAGE_RECS;
    sr.bar = reinterpret_cast<unsigned long long *>(k8_dyn_smem + (size_t)WARPS * NB_STAGES * NB_STAGE_RECS * sizeof(int4)) + (threadIdx.x >> 5) * NB_STAGES;
    sr.phase = 0;
    if (STG) {
        if ((threadIdx.x & 31) == 0)
            for (int st = 0; st < NB_STAGES; ++st) mbar_init(&sr.bar[st], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncthreads();
    }
    unsigned rank = 0;
    unsigned long long *s_q = &s_q_;
    long long *s_cost = &s_cost_, *s_bestcost = &s_bestcost_;
    int *s_cur = &s_cur_, *s_stop = &s_stop_;
    if (CL > 1) {                                       // the scalars of the cluster's first CTA, through DSMEM
        cg::cluster_group cluster = cg::this_cluster();
        rank = cluster.block_rank();
        s_q = cluster.map_shared_rank(&s_q_, 0);
        s_cost = cluster.map_shared_rank(&s_cost_, 0); s_bestcost = cluster.map_shared_rank(&s_bestcost_, 0);
        s_cur = cluster.map_shared_rank(&s_cur_, 0); s_stop = cluster.map_shared_rank(&s_stop_, 0);
    }
    const int t = (int)rank * BLK + threadIdx.x, warp = t >> 5, lane = t & 31;
    const size_t n1 = (size_t)n + 1, slot = blockIdx.x / CL;
    TourView V;
    V.n = n;
    V.pxy = S.pxy + slot * n1;
    V.tid = S.tid + slot * n1;
    V.dn = S.dn + slot * n1;
    V.dns = V.dn;
    V.pos = S.pos + slot * (size_t)n;
    BigBox *boxes = S.boxes + slot * (size_t)n;
    NbCache C;
    C.lines = S.nb + slot * (size_t)n * NB_WORDS;
    C.heap_cap = S.heap_cap;
    C.restart_pct = S.restart_pct;
    C.giant_age = 0;
    C.heap = S.nb_heap + slot * (size_t)C.heap_cap;
    C.heap_next = S.nb_next + slot * 8;
    C.cnt8 = S.cnt8 + slot * (size_t)nb_cnt8_stride(n);
    V.o8 = S.o8 + slot * (size_t)n;
    // One CTA per descent: the orientation / epoch bytes live in SHARED memory for the whole descent (n bytes of dynamic
    // shared memory; nothing outside this kernel reads them).  The scan looks up two of them per live record and the move
    // flips one per reversed node -- as global bytes these were L1 gathers and scattered one-byte stores to HBM (K8_O8_SMEM:
    // pla85900 300 -> 271, usa13509 96 -> 88, d15112 106 -> 92 us per scan + move, profiles/r02_pruned.md; on pr2392, whose
    // bytes fit L1 anyway, the smaller L1 costs 4 %).  Unconditional for this instantiation, so that the compiler KNOWS the
    // address space (LDS / STS, not generic loads: a run-time switch on n lost the whole gain).  A cluster keeps the bytes
    // in global memory: its CTAs would each need a copy and every flip in all of them.
    if (K8_O8_SMEM && CL == 1 && !STG) V.o8 = k8_dyn_smem;
    int *best_path = S.best_path + slot * (size_t)n;
    uint4 *wbuf = wbuf_all + ((size_t)blockIdx.x * WARPS + (threadIdx.x >> 5)) * PR_WBUF;
    if (t == 0) { *s_bestcost = -1; *s_stop = 0; S.best_cost[slot] = -1; S.best_start[slot] = -1; }
    long long ncand = 0;

    for (;;) {
        team_sync<CL>();
        if (t == 0) {
            const bool halted = stop != nullptr && *(volatile const int *)stop != 0;
            *s_cur = halted ? nstarts : atomicAdd(next_start, 1);
        }
        team_sync<CL>();
        const int cur = *s_cur;
        if (CL > 1) team_sync<CL>();                    // nobody leaves (or rewrites s_cur) while a peer still reads it remotely
        if (cur >= nstarts) break;
        const int *order = order_all + (size_t)cur * n;
        for (int p = t; p <= n; p += TEAM) {
            const int node = order[p == n ? 0 : p];
            V.pxy[p] = xy[node];
            V.tid[p] = node;
            if (p < n) { V.pos[node] = p; C.lines[(size_t)p * NB_WORDS] = NB_HDR_STALE; C.cnt8[p] = 1; V.o8[p] = NB_EP_INIT << 1; }    // a new tour: every cached list is stale
        }
        if (t == 0) { *s_cost = 0; *s_q = 0; C.heap_next[0] = 0; C.heap_next[1] = 0; C.heap_next[2] = 0; C.heap_next[3] = 0; C.heap_next[4] = 0; }
        if (threadIdx.x == 0) { s_key_[0] = KEY_EMPTY; s_key_[1] = KEY_EMPTY; s_best = -1; }
        team_sync<CL>();
        long long local = 0;
        for (int p = t; p < n; p += TEAM) {
            const double2 u = V.pxy[p], v = V.pxy[p + 1];
            const int k = M::K(u.x, u.y, v.x, v.y);
            V.dn[p] = k;
            if (p == 0) V.dn[n] = k;
            local += (long long)(k - KOFF);
        }
        for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(FULL, local, o);
        if (lane == 0) atomicAdd((unsigned long long *)s_cost, (unsigned long long)local);
        team_sync<CL>();
        long long moves = 0;
        unsigned par = 0;
        bool cut = false;
        PosPatch F = POSPATCH_NONE;
        // two barriers per scan + move.  The argmin key is double-buffered by iteration parity; the positions of the two
        // segment ends are written behind the last barrier and bridged by F.  The key stays per CTA and is combined by reading the peers' copies: a 64-bit atomicMin
        // is not atomic between a CTA's own shared-memory path and its peers' distributed-shared-memory path (measured,
        // scripts/cluster_probe.cu: lost updates), 64-bit atomicAdd is.
        int ep = NB_EP_NONE, hp = 0, hsnap = 0;
#ifdef TSP_PHASE_TIMING
        // thread 0 of team 0: ns in the scan (own part / incl. the wait at its barrier), the giant phase, the move (own part / incl. barrier)
        unsigned long long k8t[5] = {0, 0, 0, 0, 0}, k8a = 0, k8b = 0;
#define K8_MARK(slot) { asm volatile("mov.u64 %0, %globaltimer;" : "=l"(k8b)); k8t[slot] += k8b - k8a; k8a = k8b; }
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(k8a));
#else
#define K8_MARK(slot)
#endif
        for (;;) {
            int best = -1;
            unsigned long long bestkey = KEY_EMPTY;
            cached_scan<M, NE, false, STG>(V, G, F, ep, warp, WARPS * CL, hsnap, hp, s_q, boxes, wbuf, C, s_ring[threadIdx.x >> 5], &s_best, best, bestkey, ncand, &sr);
            bestkey = warp_min_key(bestkey);
            if (lane == 0 && bestkey != KEY_EMPTY) atomicMin(&s_key_[par], bestkey);
            K8_MARK(0)
            team_sync<CL>();
            K8_MARK(1)
            {   // giant search boxes met in this scan (rare): all warps of the team fill their heap runs now, one more barrier
                const unsigned long long q = *s_q;
                if (q >> 40) {
                    bestkey = KEY_EMPTY;
                    giant_phase<M, NE>(V, &G, F, C, warp, WARPS * CL, q, boxes, wbuf, &best, &bestkey, &ncand);
                    bestkey = warp_min_key(bestkey);
                    if (lane == 0 && bestkey != KEY_EMPTY) atomicMin(&s_key_[par], bestkey);
                    team_sync<CL>();
                    if (t == 0) *s_q = 0;
                }
            }
            unsigned long long kbest = s_key_[par];
            if (CL > 1) {
                cg::cluster_group cluster = cg::this_cluster();
#pragma unroll
                for (int r = 0; r < CL; ++r) kbest = min(kbest, *cluster.map_shared_rank(&s_key_[par], r));
            }
            if (kbest == KEY_EMPTY) break;                      // local optimum (refinment.c:27)
            hsnap = heap_checkpoint(C, n, t, TEAM, hp, *(volatile int *)(C.heap_next + hp), *(volatile int *)(C.heap_next + 3));   // between the last two barriers: nobody changes it here
            ep = (int)(moves % NB_EPOCHS);
            const int a = key_a(kbest), b = key_b(kbest);
            int seg_s, seg_e;
            K8_MARK(2)
            apply_move<M, true, (CL >= 8 ? 1 : TSP_AM_B)>(V, a, b, t, TEAM, false, seg_s, seg_e, ep);
            K8_MARK(3)
            if (t == 0) {
                *s_cost += key_delta(kbest);
                if (stop != nullptr) *s_stop = *(volatile const int *)stop;
            }
            if (threadIdx.x == 0) { s_key_[par ^ 1] = KEY_EMPTY; s_best = -1; }
            ++moves;
            par ^= 1;
            team_sync<CL>();
            K8_MARK(4)
            F.node0 = b; F.pos0 = seg_s; F.node1 = V.tid[seg_e]; F.pos1 = seg_e;
            F.node2 = a; F.node3 = V.tid[seg_e + 1];
            if (t == 0) { V.pos[F.node0] = F.pos0; V.pos[F.node1] = F.pos1; }
            if (stop != nullptr && *s_stop != 0) { cut = true; break; }      // time limit: checked before the next scan
        }
#ifdef TSP_PHASE_TIMING
        if (t == 0 && blockIdx.x == 0)
            printf("[tspcuda] K8 CTA 0 warp 1, all its scans so far: heap stream %llu ns, lines %llu ns, attention + evaluation tail %llu ns per scan (%llu scans)\n",
                   g_scan_ns[0] / (g_scan_ns[3] + 1), g_scan_ns[1] / (g_scan_ns[3] + 1), g_scan_ns[2] / (g_scan_ns[3] + 1), g_scan_ns[3]);
        if (t == 0 && blockIdx.x == 0)
            printf("[tspcuda] K8 team 0, start %d: %lld moves; per scan+move ns: scan own %llu, wait at its barrier %llu, giant phase + checkpoint %llu, move own %llu, bookkeeping + barrier %llu\n",
                   first_start + cur * stride, moves, k8t[0] / (moves + 1), k8t[1] / (moves + 1), k8t[2] / (moves + 1), k8t[3] / (moves + 1), k8t[4] / (moves + 1));
#endif
#undef K8_MARK
        // left behind a barrier with a team-uniform condition: every thread is here, the tour is final
        const long long fc = *s_cost;
        if (t == 0) { final_cost[cur] = fc; moves_out[cur] = moves; status_out[cur] = cut ? MS_CUT : MS_COMPLETE; }
        const long long bcst = *s_bestcost;
        const bool better = bcst < 0 || fc < bcst;                        // strict '<', ascending starts per team
        if (better)
            for (int p = t; p < n; p += TEAM) best_path[V.tid[p]] = V.tid[p + 1];
        team_sync<CL>();
        if (better && t == 0) {
            *s_bestcost = fc;
            S.best_cost[slot] = fc;
            S.best_start[slot] = first_start + cur * stride;
        }
    }
    for (int off = 16; off > 0; off >>= 1) ncand += __shfl_down_sync(FULL, ncand, off);
    if (lane == 0 && ncand) atomicAdd(cand_total, (unsigned long long)ncand);
}

}  // namespace tspcuda
