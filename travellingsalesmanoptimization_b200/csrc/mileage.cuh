// mileage.cuh -- K11/K12: Extra Mileage (heuristics.c:156-210 h_ExtraMileage, :290-367
// h_extramileage_util) on the device, exact.
//
// The reference spends O(unvisited x tour edges) per insertion: over unvisited nodes i (ascending) and
// tour edges j (in edges[] order) it keeps the FIRST strict minimum of c(u,i) + c(i,v) - c(u,v).  An
// insertion replaces edges[j*] by (u,i*) in place and appends (i*,v); every other edge keeps its
// content and its index.  So the per-node minimum (best[i], edge[i]) = first minimum over the edge
// array survives an insertion unless edge[i] == j*: otherwise only the two new edges can change it
// (index j*: wins ties against a larger index; the appended edge has the largest index: strict '<'
// only).  Per insertion: every unvisited node is refreshed in O(1) by its owning thread, the few
// nodes whose best edge was replaced are rescanned over the whole edge array, one warp each, and the
// global argmin of (best[i], i) -- lexicographic, the reference's loop order -- picks the next
// insertion.  Costs are computed on the fly from the coordinates (K-offset integers, metric.cuh);
// all arithmetic is integer, so the result is the reference's bit for bit.
#pragma once
#include "kernels.cuh"

namespace tspcuda {

struct EmState {
    int *visited;            // [n] 0/1
    int *eu, *ev;            // [n] edge array in the reference's order
    int *best, *bedge;       // [n] per-node first-minimum mileage (plain integer) and its edge index
    int *path;               // [n] successor array being built
    int *list;               // [n] nodes whose best edge was replaced (rescan queue)
    unsigned long long *key; // [2] argmin slots (double-buffered)
    int *list_count;         // [2]
    unsigned long long *rk;  // [n] rescan keys (mileage, edge index) of the queued nodes, ~0 when idle
    int *done;               // [1] CTAs that have finished their share of the rescans of this insertion
    long long *cost;         // [1] tour cost: 2 c(A,B) + the mileage of every insertion
    int *init_pair;          // [2] nodeA, nodeB
};

constexpr int EM_BIAS = 64;                         // mileage >= -2 with rounded costs; bias keeps the key unsigned

template <class M>
__device__ __forceinline__ int em_delta(const double2 *xy, int u, int v, const double2 pi, int kuv) {
    const double2 pu = xy[u], pv = xy[v];
    return (M::K(pu.x, pu.y, pi.x, pi.y) - KOFF) + (M::K(pi.x, pi.y, pv.x, pv.y) - KOFF) - kuv;
}

// EM_MAX (heuristics.c:169-178): first pair i<j, ascending, with the strictly largest cost; (0,1) when every cost is 0.
template <class M>
__global__ void em_max_kernel(const double2 *__restrict__ xy, int n, unsigned long long *slot) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    unsigned long long best = 0;
    for (int i = warp; i < n - 1; i += nwarps) {
        const double2 pi = xy[i];
        for (int j = i + 1 + lane; j < n; j += 32) {
            const double2 pj = xy[j];
            const unsigned long long c = (unsigned long long)(M::K(pi.x, pi.y, pj.x, pj.y) - KOFF);
            const unsigned long long key = (c << 34) | (0x3FFFFFFFFull - (((unsigned long long)i << 17) | (unsigned long long)j));
            best = max(best, key);
        }
    }
    for (int off = 16; off > 0; off >>= 1) best = max(best, __shfl_down_sync(FULL, best, off));
    if (lane == 0 && best) atomicMax(slot, best);
}

template <class M, int BLK>
__global__ void __launch_bounds__(BLK)
extra_mileage_kernel(const double2 *__restrict__ xy, int n, EmState S) {
    cg::grid_group grid = cg::this_grid();
    __shared__ unsigned long long s_bk, s_kb;
    __shared__ int s_nlist, s_last;
    const int gtid = blockIdx.x * BLK + threadIdx.x, gsize = gridDim.x * BLK;
    const int lane = threadIdx.x & 31;
    const int gw = gtid >> 5, nw = gsize >> 5;
    const int A = S.init_pair[0], B = S.init_pair[1];

    // ---- set-up: edges (A,B), (B,A); every unvisited node starts on edge 0 (equal mileage on both, strict '<')
    unsigned long long mykey = ~0ull;
    {
        const double2 pa = xy[A], pb = xy[B];
        const int kab = M::K(pa.x, pa.y, pb.x, pb.y) - KOFF;
        for (int i = gtid; i < n; i += gsize) {
            const bool vis = (i == A || i == B);
            S.visited[i] = vis;
            S.path[i] = 0;
            if (!vis) {
                const int d = em_delta<M>(xy, A, B, xy[i], kab);
                S.best[i] = d; S.bedge[i] = 0;
                mykey = min(mykey, ((unsigned long long)(unsigned)(d + EM_BIAS) << 32) | (unsigned)i);
            }
        }
        if (gtid == 0) {
            S.eu[0] = A; S.ev[0] = B; S.eu[1] = B; S.ev[1] = A;
            S.key[1] = ~0ull; S.list_count[0] = 0; S.list_count[1] = 0;
            *S.cost = 2LL * kab;                    // solution.cost = 2 * c(A,B), heuristics.c:193
        }
    }
    for (int off = 16; off > 0; off >>= 1) mykey = min(mykey, __shfl_down_sync(FULL, mykey, off));
    if (lane == 0 && mykey != ~0ull) atomicMin(&S.key[0], mykey);
    grid.sync();
    if (gtid == 0) { S.path[A] = B; S.path[B] = A; }    // :195-196, after the zero-fill above

    for (int m = 2; m < n; ++m) {                   // m = edges so far = nodes in the tour
        const int par = m & 1;                      // key[par] holds the argmin of this insertion
        // (one reader per CTA of the words the whole grid needs, handed on through shared memory: see K7, pruned.cuh)
        if (threadIdx.x == 0) s_kb = *(volatile unsigned long long *)&S.key[par];
        __syncthreads();
        const unsigned long long kb = s_kb;
        if (kb == ~0ull) break;                     // no unvisited node left (cannot happen before m == n)
        const int istar = (int)(kb & 0xFFFFFFFFu);
        const int mileage = (int)(kb >> 32) - EM_BIAS;
        const int jstar = S.bedge[istar];
        const int u = S.eu[jstar], v = S.ev[jstar]; // old content of the replaced edge: nobody writes edges in this phase
        const double2 ps = xy[istar], pu = xy[u], pv = xy[v];
        const int k_us = M::K(pu.x, pu.y, ps.x, ps.y) - KOFF;      // new edge (u, i*) at index j*
        const int k_sv = M::K(ps.x, ps.y, pv.x, pv.y) - KOFF;      // new edge (i*, v) at index m
        // ---- phase A: O(1) refresh of every unvisited node; nodes that sat on edge j* are queued
        mykey = ~0ull;
        for (int i = gtid; i < n; i += gsize) {
            if (S.visited[i] || i == istar) continue;
            const double2 pi = xy[i];
            int bd = S.best[i], be = S.bedge[i];
            if (be == jstar) {
                S.list[atomicAdd(&S.list_count[par], 1)] = i;
                continue;
            }
            const int d1 = (M::K(pu.x, pu.y, pi.x, pi.y) - KOFF) + (M::K(pi.x, pi.y, ps.x, ps.y) - KOFF) - k_us;
            if (d1 < bd || (d1 == bd && jstar < be)) { bd = d1; be = jstar; }
            const int d2 = (M::K(ps.x, ps.y, pi.x, pi.y) - KOFF) + (M::K(pi.x, pi.y, pv.x, pv.y) - KOFF) - k_sv;
            if (d2 < bd) { bd = d2; be = m; }
            S.best[i] = bd; S.bedge[i] = be;
            mykey = min(mykey, ((unsigned long long)(unsigned)(bd + EM_BIAS) << 32) | (unsigned)i);
        }
        for (int off = 16; off > 0; off >>= 1) mykey = min(mykey, __shfl_down_sync(FULL, mykey, off));
        if (lane == 0 && mykey != ~0ull) atomicMin(&S.key[par ^ 1], mykey);
        grid.sync();
        // ---- phase B: the insertion itself (one thread) and the rescans (one warp per queued node).  Rescans
        // substitute the two new edges at indices j* and m instead of reading them, so they do not wait for the write.
        if (gtid == 0) {
            S.path[u] = istar; S.ev[jstar] = istar;                // edges[j*] = (u, i*)      :348-350
            S.path[istar] = v; S.eu[m] = istar; S.ev[m] = v;       // edges[m]  = (i*, v)      :352-354
            S.visited[istar] = 1;                                   // :357
            *S.cost += mileage;                                     // :360
            S.key[par] = ~0ull;                                     // slot of the insertion after next
        }
        // Once the edge array is long, the rescans are cut into (queued node, chunk of the edge array) items that just cover
        // the grid (round 2 first gave each queued node ONE CTA: up to 335 trips of two dependent gathers per thread while
        // the CTAs without a node idled, 80 us per insertion on pla85900): warp and CTA minima in shared memory, one global
        // atomicMin per item into rk[], and the CTA that finishes LAST (a counter behind a fence) writes the nodes' new
        // (best, edge) and their keys -- no third grid barrier.
        if (threadIdx.x == 0) s_nlist = *(volatile int *)&S.list_count[par];
        __syncthreads();
        const int nlist = s_nlist;
        if (m >= 4 * BLK) {
            // work items (queued node q, chunk of the edge array): nch chunks per node so that the items just cover the grid
            const int nch = max(1, min((int)gridDim.x / max(1, nlist), (m + BLK) / BLK));
            const int len = (m + nch) / nch;                        // ceil((m + 1) / nch) edges per chunk
            for (int item = blockIdx.x; item < nlist * nch; item += gridDim.x) {     // CTA-uniform trip count
                if (threadIdx.x == 0) s_bk = ~0ull;
                __syncthreads();
                const int q = item / nch, j0 = (item - q * nch) * len, j1 = min(m + 1, j0 + len);
                const int i = S.list[q];
                const double2 pi = xy[i];
                unsigned long long bk = ~0ull;                      // (mileage, edge index): first minimum in edge order
                for (int j = j0 + threadIdx.x; j < j1; j += BLK) {
                    int eu = S.eu[j], ev = S.ev[j];
                    if (j == jstar) { eu = u; ev = istar; }
                    if (j == m) { eu = istar; ev = v; }
                    const double2 a = xy[eu], b = xy[ev];
                    const int kuv = M::K(a.x, a.y, b.x, b.y) - KOFF;
                    const int d = (M::K(a.x, a.y, pi.x, pi.y) - KOFF) + (M::K(pi.x, pi.y, b.x, b.y) - KOFF) - kuv;
                    bk = min(bk, ((unsigned long long)(unsigned)(d + EM_BIAS) << 32) | (unsigned)j);
                }
                for (int off = 16; off > 0; off >>= 1) bk = min(bk, __shfl_down_sync(FULL, bk, off));
                if (lane == 0 && bk != ~0ull) atomicMin(&s_bk, bk);
                __syncthreads();
                if (threadIdx.x == 0 && s_bk != ~0ull) atomicMin(&S.rk[q], s_bk);
            }
            if (nlist > 0) {
                __syncthreads();
                if (threadIdx.x == 0) {
                    __threadfence();                                // this CTA's rk[] minima before its tick
                    s_last = atomicAdd(S.done, 1) == (int)gridDim.x - 1;
                }
                __syncthreads();
                if (s_last) {                                       // every CTA's minima are in rk[]
                    __threadfence();
                    for (int q = threadIdx.x; q < nlist; q += BLK) {
                        const unsigned long long rk = *(volatile unsigned long long *)&S.rk[q];
                        const int i = S.list[q];
                        const int bd = (int)(rk >> 32) - EM_BIAS;
                        S.best[i] = bd; S.bedge[i] = (int)(rk & 0xFFFFFFFFu);
                        atomicMin(&S.key[par ^ 1], ((unsigned long long)(unsigned)(bd + EM_BIAS) << 32) | (unsigned)i);
                        S.rk[q] = ~0ull;
                    }
                    if (threadIdx.x == 0) *S.done = 0;
                }
            }
        } else {
            for (int q = gw; q < nlist; q += nw) {
                const int i = S.list[q];
                const double2 pi = xy[i];
                unsigned long long bk = ~0ull;
                for (int j = lane; j <= m; j += 32) {
                    int eu = S.eu[j], ev = S.ev[j];
                    if (j == jstar) { eu = u; ev = istar; }
                    if (j == m) { eu = istar; ev = v; }
                    const double2 a = xy[eu], b = xy[ev];
                    const int kuv = M::K(a.x, a.y, b.x, b.y) - KOFF;
                    const int d = (M::K(a.x, a.y, pi.x, pi.y) - KOFF) + (M::K(pi.x, pi.y, b.x, b.y) - KOFF) - kuv;
                    bk = min(bk, ((unsigned long long)(unsigned)(d + EM_BIAS) << 32) | (unsigned)j);
                }
                for (int off = 16; off > 0; off >>= 1) bk = min(bk, __shfl_down_sync(FULL, bk, off));
                if (lane == 0) {
                    const int bd = (int)(bk >> 32) - EM_BIAS;
                    S.best[i] = bd; S.bedge[i] = (int)(bk & 0xFFFFFFFFu);
                    atomicMin(&S.key[par ^ 1], ((unsigned long long)(unsigned)(bd + EM_BIAS) << 32) | (unsigned)i);
                }
            }
        }
        grid.sync();
        if (gtid == 0) S.list_count[par] = 0;       // reused two insertions later
    }
}

}  // namespace tspcuda
