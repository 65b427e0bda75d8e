// patching.cuh -- K13/K14: the patching heuristic of the Benders loop (cx_patching,
// src/algorithms/cplex_model.c:693-755) on the device, exact.
//
// Input is a successor array that is NOT a tour: several disjoint cycles ("components", comp[v] = label
// of the cycle of node v, as cx_build_sol leaves them).  One patch merges two cycles: over all ordered
// node pairs (a,b) of different components, a-major / b-minor, the reference keeps the first strict
// minimum of
//     delta = [c(a, succ b) + c(b, succ a)] - [c(a, succ a) + c(b, succ b)]        (:712-714, this order)
// among the pairs with delta > 0 (:716), then sets path[a] = succ b, path[b] = succ a (:734-735),
// cost += delta (:738), relabels the whole cycle of b with comp[a] (:741-752) and decrements ncomp.
// delta is symmetric in (a,b), so the first minimum in the ordered loop is the minimum of the packed key
// (delta, a, b) over the UNORDERED pairs a < b -- the same 64-bit key as the 2-opt scan (metric.cuh::pack_key).
// Costs are computed on the fly (two fresh edge costs per pair, K-offset integers, exact), the per-node
// terms c(a, succ a) and the successor coordinates are kept per node and refreshed for the two nodes a
// patch touches.  The relabel walk "from succ b until b" covers exactly the cycle of b, i.e. every node
// carrying comp[b] -- done as one parallel pass.
#pragma once
#include "kernels.cuh"

namespace tspcuda {


struct PatchState {
    int *path, *comp;          // [n] successor array and component labels (in/out)
    double2 *sxy;              // [n] coordinates of succ(v)
    int *dn;                   // [n] K-offset cost of (v, succ v)
    unsigned long long *key;   // [1] argmin slot
    long long *cost;           // [1] running cost (integer-valued)
    int *info;                 // [0] patches applied, [1] 1 = a scan found no pair with delta > 0
    int *log;                  // [3 * cap] (a, b, delta) per patch, may be null
    int log_cap;
};

template <class M>
__global__ void patch_prep_kernel(const double2 *__restrict__ xy, int n, PatchState S) {
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) {
        const double2 p = xy[v], s = xy[S.path[v]];
        S.sxy[v] = s;
        S.dn[v] = M::K(p.x, p.y, s.x, s.y);
    }
}

// K13: the upper triangle of (a-tile, b-tile) pairs, b-tile >= a-tile, is linearised row by row and dealt out to the
// CTAs round-robin (tile pair t goes to CTA t mod gridDim.x), so every CTA gets the same number of tile pairs to within
// one -- all CTAs are resident at once, and a split by a-tile would leave the CTAs of the last rows idle for half the
// scan.  Within a tile pair thread t owns node a = ta * TILE + t in registers; the b side is staged in shared memory
// and read as broadcasts.
template <class M, int TILE>
__global__ void __launch_bounds__(TILE)
patch_scan_kernel(const double2 *__restrict__ xy, int n, PatchState S) {
    __shared__ double2 s_xy[TILE], s_sxy[TILE];
    __shared__ int s_dn[TILE], s_comp[TILE];
    __shared__ unsigned long long s_key;
    if (S.info[1]) return;                               // an earlier scan of this batch found nothing
    const long long N = (n + TILE - 1) / TILE;
    const long long T = N * (N + 1) / 2;
    if (threadIdx.x == 0) s_key = KEY_EMPTY;
    int best = 0x3FFFFFFF;
    unsigned long long bestkey = KEY_EMPTY;
    for (long long t = blockIdx.x; t < T; t += gridDim.x) {
        // row ta starts at R(ta) = ta * N - ta (ta - 1) / 2
        long long ta = (long long)(((double)(2 * N + 1) - sqrt((double)((2 * N + 1) * (2 * N + 1) - 8 * t))) * 0.5);
        ta = max(0ll, min(N - 1, ta));
        while (ta > 0 && ta * N - ta * (ta - 1) / 2 > t) --ta;
        while ((ta + 1) * N - (ta + 1) * ta / 2 <= t) ++ta;
        const int tb = (int)(ta + (t - (ta * N - ta * (ta - 1) / 2)));
        const int a = (int)ta * TILE + threadIdx.x;
        const bool va = a < n;
        const double2 pa = va ? xy[a] : make_double2(0, 0), psa = va ? S.sxy[a] : make_double2(0, 0);
        const int dna = va ? S.dn[a] : 0, ca = va ? S.comp[a] : -1;
        __syncthreads();
        const int bl = tb * TILE + threadIdx.x;
        if (bl < n) { s_xy[threadIdx.x] = xy[bl]; s_sxy[threadIdx.x] = S.sxy[bl]; s_dn[threadIdx.x] = S.dn[bl]; s_comp[threadIdx.x] = S.comp[bl]; }
        __syncthreads();
        if (!va) continue;
        const int cnt = min(TILE, n - tb * TILE);
        const int k0 = (tb == (int)ta) ? threadIdx.x + 1 : 0; // b > a
        for (int k = k0; k < cnt; ++k) {
            if (s_comp[k] == ca) continue;                    // :703
            const double2 pb = s_xy[k], psb = s_sxy[k];
            // swapped - current (:712-714); the four K offsets cancel (and so do the two sign flips of the packed
            // FP32x2 cost sequence, metric.cuh: 2 * 2^31 wraps away)
            int k1, k2;
            M::KK(pa.x, pa.y, psb.x, psb.y, pb.x, pb.y, psa.x, psa.y, k1, k2);
            const int d = (int)((unsigned)k1 + (unsigned)k2 - (unsigned)dna - (unsigned)s_dn[k]);
            if (d > 0 && d <= best) {                         // :716 (delta > 0), first strict minimum by the key
                best = d;
                const unsigned long long key = pack_key(d, a, tb * TILE + k);
                if (key < bestkey) bestkey = key;
            }
        }
    }
    bestkey = warp_min_key(bestkey);
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && bestkey != KEY_EMPTY) atomicMin(&s_key, bestkey);
    __syncthreads();
    if (threadIdx.x == 0 && s_key != KEY_EMPTY) atomicMin(S.key, s_key);
}

// K14: apply the patch found by K13 (:727-754) and re-arm the argmin slot.  ONE CTA: a patch is O(n) trivial
// work (the relabel pass) between two scans, and a single CTA needs no grid barrier.
template <class M>
__global__ void __launch_bounds__(1024) patch_apply_kernel(const double2 *__restrict__ xy, int n, PatchState S) {
    if (S.info[1]) return;
    const unsigned long long k = *S.key;
    if (k == KEY_EMPTY) {                                     // the reference would index path[-1] here (:730-735)
        __syncthreads();
        if (threadIdx.x == 0) S.info[1] = 1;
        return;
    }
    const int a = key_a(k), b = key_b(k), delta = key_delta(k);
    const int sa = S.path[a], sb = S.path[b];
    const int ca = S.comp[a], cb = S.comp[b];
    __syncthreads();                                          // everybody has read the key, old successors and labels
    for (int v = threadIdx.x; v < n; v += blockDim.x)
        if (S.comp[v] == cb) S.comp[v] = ca;                  // :741-752: the cycle of b takes the label of a
    if (threadIdx.x == 0) {
        S.path[a] = sb; S.path[b] = sa;                       // :734-735
        const double2 pa = xy[a], pb = xy[b], psa = xy[sa], psb = xy[sb];
        S.sxy[a] = psb; S.sxy[b] = psa;
        S.dn[a] = M::K(pa.x, pa.y, psb.x, psb.y);
        S.dn[b] = M::K(pb.x, pb.y, psa.x, psa.y);
        *S.cost += delta;                                     // :738
        const int done = S.info[0];
        if (S.log && done < S.log_cap) { S.log[3 * done] = a; S.log[3 * done + 1] = b; S.log[3 * done + 2] = delta; }
        S.info[0] = done + 1;
        *S.key = KEY_EMPTY;
    }
}

}  // namespace tspcuda
