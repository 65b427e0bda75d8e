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
// Work distribution.  Phase 1: one thread per tour position; search boxes holding at most
// PR_SMALL_NODES nodes are walked by that thread, larger ones (the long edges of a nearest-neighbour
// tour) are queued.  Phase 2 (after a grid barrier): the queued boxes are cut into (box, cell row)
// units -- a cell row of a box is one contiguous range of the cell-sorted node array -- and the units
// are split evenly over all warps of the grid, lanes striding over the nodes of a row.
#pragma once
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

constexpr int PRUNED_BLOCK = 512;
constexpr int PR_SMALL_ROWS = 12;
constexpr int PR_SMALL_NODES = 160;
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

// ---- candidate evaluation ---------------------------------------------------------------------
// Everything a search around the node at position p needs, loaded once.
struct Probe {
    double ux, uy, sx, sy, px, py;     // node at p, its successor (p+1), its predecessor (p-1)
    int p, pm, ida, idm, dnp, dnm;     // pm = p-1 (cyclic); node ids at p and pm; K-offset edge costs
};

__device__ __forceinline__ void probe_load(const TourView &V, int p, Probe &P) {
    const int n = V.n;
    P.p = p;
    P.pm = p == 0 ? n - 1 : p - 1;
    const double2 u = V.pxy[p], s = V.pxy[p + 1], q = V.pxy[P.pm];
    P.ux = u.x; P.uy = u.y; P.sx = s.x; P.sy = s.y; P.px = q.x; P.py = q.y;
    P.ida = V.tid[p]; P.idm = V.tid[P.pm];
    P.dnp = V.dn[p]; P.dnm = V.dn[P.pm];
}

// pos[] of the two end nodes of the segment reversed by the PREVIOUS move is written by one thread
// right after the barrier that ends that move (every thread reads pos[a], pos[b] on entry of
// apply_move, so it cannot be written earlier); scans that start behind the same barrier take those
// two positions from registers instead of racing with that write.
struct PosPatch { int node0, pos0, node1, pos1; };

template <class M>
__device__ __forceinline__ void cand_eval(const TourView &V, const Probe &P, const PosPatch &F, int w, double wx, double wy,
                                          int &best, unsigned long long &bestkey) {
    const int n = V.n;
    const int K = M::K(P.ux, P.uy, wx, wy);
    const bool c1 = K < P.dnp, c2 = K < P.dnm;
    if (!(c1 | c2)) return;
    const int q = (w == F.node0) ? F.pos0 : (w == F.node1) ? F.pos1 : V.pos[w];
    const int diff = abs(P.p - q);
    if (diff <= 1 || diff == n - 1) return;       // same node / tour neighbours: refinment.c:55
    if (c1) {                                      // pair {p, q}: edges (p,p+1), (q,q+1) out, (p,q), (p+1,q+1) in
        const double2 ws = V.pxy[q + 1];
        const int K2 = M::K(P.sx, P.sy, ws.x, ws.y);
        const int d = (int)((unsigned)K + (unsigned)K2 - (unsigned)P.dnp - (unsigned)V.dn[q]);
        if (d <= best) {
            best = d;
            const unsigned long long key = pack_key(d, P.ida, w);
            if (key < bestkey) bestkey = key;
        }
    }
    if (c2) {                                      // pair {p-1, q-1}: (p,q) is its second new edge
        const int qm = q == 0 ? n - 1 : q - 1;
        const double2 wp = V.pxy[qm];
        const int K2 = M::K(P.px, P.py, wp.x, wp.y);
        const int d = (int)((unsigned)K2 + (unsigned)K - (unsigned)P.dnm - (unsigned)V.dn[qm]);
        if (d <= best) {
            best = d;
            const unsigned long long key = pack_key(d, P.idm, V.tid[qm]);
            if (key < bestkey) bestkey = key;
        }
    }
}

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

// ---- K7 ---------------------------------------------------------------------------------------
template <class M, int BLK>
__global__ void __launch_bounds__(BLK, 1)
descend_pruned_kernel(TourView V, SpatialGrid G, unsigned long long *key, DescState *st, MoveRec *log, long long log_cap,
                      long long max_moves, unsigned long long *qctr, BigBox *boxes) {
    constexpr int WARPS = BLK / 32;
    cg::grid_group grid = cg::this_grid();
    __shared__ unsigned long long s_key;
    const int n = V.n;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gtid = blockIdx.x * BLK + threadIdx.x;
    const int gsize = gridDim.x * BLK;
    const long long gw = (long long)blockIdx.x * WARPS + warp;
    const long long Wtot = (long long)gridDim.x * WARPS;
    PosPatch F = {-1, 0, -1, 0};

    for (long long it = 0; it < max_moves; ++it) {
        if (threadIdx.x == 0) s_key = KEY_EMPTY;
        __syncthreads();
        int best = -1;
        unsigned long long bestkey = KEY_EMPTY;
        long long ncand = 0;

        // ---- phase 1: one thread per position, small boxes inline, large boxes queued ----------
        for (int p = gtid; p < n; p += gsize) {
            Probe P;
            probe_load(V, p, P);
            int cx0, cx1, cy0, cy1;
            if (!probe_box<M>(G, P, cx0, cx1, cy0, cy1)) continue;
            const int nrows = cy1 - cy0 + 1;
            int total = PR_SMALL_NODES + 1;
            if (nrows <= PR_SMALL_ROWS) {
                total = 0;
                for (int cy = cy0; cy <= cy1; ++cy)
                    total += G.cell_start[cy * G.W + cx1 + 1] - G.cell_start[cy * G.W + cx0];
            }
            if (total > PR_SMALL_NODES) {
                const unsigned long long old = atomicAdd(qctr, (1ull << 40) | (unsigned long long)nrows);
                BigBox b;
                b.p = p; b.cy0 = cy0; b.cx0 = cx0; b.cx1 = cx1; b.nrows = nrows; b.pad = 0;
                b.base = (long long)(old & PR_UNIT_MASK);
                boxes[old >> 40] = b;
                continue;
            }
            ncand += total;
            for (int cy = cy0; cy <= cy1; ++cy) {
                const int lo = G.cell_start[cy * G.W + cx0], hi = G.cell_start[cy * G.W + cx1 + 1];
                for (int k = lo; k < hi; ++k) {
                    const double2 wxy = G.cxy[k];
                    cand_eval<M>(V, P, F, G.cnode[k], wxy.x, wxy.y, best, bestkey);
                }
            }
        }
        grid.sync();

        // ---- phase 2: queued boxes, (box,row) units split evenly over the warps of the grid ----
        {
            const unsigned long long q = *(volatile unsigned long long *)qctr;
            const long long nunits = (long long)(q & PR_UNIT_MASK);
            const int nboxes = (int)(q >> 40);
            if (nunits > 0) {
                const long long per = (nunits + Wtot - 1) / Wtot;
                const long long g0 = gw * per, g1 = min(nunits, g0 + per);
                if (g0 < g1) {
                    int lo_b = 0, hi_b = nboxes - 1;            // last box with base <= g0
                    while (lo_b < hi_b) {
                        const int mid = (lo_b + hi_b + 1) >> 1;
                        if (boxes[mid].base <= g0) lo_b = mid; else hi_b = mid - 1;
                    }
                    int bi = lo_b;
                    BigBox b = boxes[bi];
                    Probe P;
                    probe_load(V, b.p, P);
                    for (long long g = g0; g < g1; ++g) {
                        while (g >= b.base + b.nrows) { b = boxes[++bi]; probe_load(V, b.p, P); }
                        const int cy = b.cy0 + (int)(g - b.base);
                        const int lo = G.cell_start[cy * G.W + b.cx0], hi = G.cell_start[cy * G.W + b.cx1 + 1];
                        if (lane == 0) ncand += hi - lo;
                        for (int k = lo + lane; k < hi; k += 32) {
                            const double2 wxy = G.cxy[k];
                            cand_eval<M>(V, P, F, G.cnode[k], wxy.x, wxy.y, best, bestkey);
                        }
                    }
                }
            }
        }
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
        const unsigned long long kbest = *(volatile unsigned long long *)kslot;
        if (kbest == KEY_EMPTY) {
            if (gtid == 0) { st->status = 1; *qctr = 0; }
            break;
        }
        const int a = key_a(kbest), b = key_b(kbest), delta = key_delta(kbest);
        int seg_s, seg_e;
        apply_move<M>(V, a, b, gtid, gsize, false, seg_s, seg_e);
        if (gtid == 0) {
            key[(it + 1) & 1] = KEY_EMPTY;
            *qctr = 0;
            const long long lc = st->log_count;
            if (lc < log_cap) { log[lc].a = a; log[lc].b = b; log[lc].delta = delta; log[lc].reserved = 0; }
            st->log_count = lc + 1;
            st->moves += 1;
            st->cost += delta;
            st->last_a = a; st->last_b = b; st->last_delta = delta;
        }
        grid.sync();
        F.node0 = b; F.pos0 = seg_s; F.node1 = V.tid[seg_e]; F.pos1 = seg_e;
        if (gtid == 0) { V.pos[F.node0] = F.pos0; V.pos[F.node1] = F.pos1; }   // deferred end-node positions
    }
}

}  // namespace tspcuda
