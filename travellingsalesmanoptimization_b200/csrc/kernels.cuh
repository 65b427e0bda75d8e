// kernels.cuh -- device code of libtspcuda.so (sm_100a).
//
// Data layout in HBM (per tour slot, all tour-POSITION ordered; position n mirrors position 0):
//   pxy[n+1]  double2   coordinates of the node at tour position p
//   tid[n+1]  int32     node id at position p
//   dn [n+1]  int32     K-offset cost of the tour edge (p, p+1)            (metric.cuh)
//   dns[n+1]  int32     dn, minus BIGPEN where the edge touches a tabu node (== dn for 2-opt)
//   pos[n]    int32     position of node v
// A 2-opt move (a,b) of the reference (refinment.c:49-69, node ids a<b) removes the tour edges
// (a,succ a), (b,succ b).  In position space the pair {i,j} = {pos a, pos b}, i<j, has
//   delta(i,j) = D(i,j) + D(i+1,j+1) - dn[i] - dn[j],      D(p,q) = cost(node at p, node at q)
// so along a diagonal every D value serves two pairs: ONE new edge cost per pair evaluation
// instead of the two the reference's node-id loop needs.  Ties are broken exactly as the
// reference does (strict '<' in ascending (a,b) node-id order) through the packed key
// (delta, min id, max id), metric.cuh::pack_key.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include "metric.cuh"

namespace cg = cooperative_groups;

namespace tspcuda {

#ifndef TSP_R
#define TSP_R 4
#endif
#ifndef TSP_UNROLL
#define TSP_UNROLL 2
#endif
constexpr int R = TSP_R;             // tour positions (rows) owned by one lane
constexpr int SR = 32 * R - 1;       // pair rows per warp strip (last D row is the next strip's first)
constexpr int UNROLL = TSP_UNROLL;
constexpr int BLOCK_SMALL = 256;      // descent kernel block size for n < BIG_N (fewer, lighter CTAs: cheaper grid syncs)
// n >= BIG_N: ONE CTA per SM.  With two 512-thread CTAs per SM the warp scheduler favours the older CTA: it ends its
// scan a third earlier and then idles at the grid barrier while the other one finishes at half occupancy (measured:
// scan-end spread 0.79 ms of 2.4 ms on pla85900).  1024 threads (8 warps per scheduler, 64 registers) is best up to
// HUGE_N, 896 threads (7 per scheduler, 72 registers) above: 1.67 / 1.74e12 pair-evals/s on pla85900, against
// 1.55e12 with 2 x 512 (profiles/).  Warp counts that are not a multiple of 4 leave the schedulers uneven and lose.
#ifndef TSP_BLOCK_BIG
#define TSP_BLOCK_BIG 1024
#endif
#ifndef TSP_BLOCK_HUGE
#define TSP_BLOCK_HUGE 896
#endif
constexpr int BLOCK_BIG = TSP_BLOCK_BIG;
constexpr int BLOCK_HUGE = TSP_BLOCK_HUGE;
constexpr int HUGE_N = 32768;
#ifndef TSP_BIG_N
#define TSP_BIG_N 7000
#endif
constexpr int BIG_N = TSP_BIG_N;
constexpr unsigned FULL = 0xFFFFFFFFu;

struct TourView {
    double2 *pxy;
    int *tid;
    int *dn;
    int *dns;
    int *pos;
    int n;
    unsigned char *o8;     // [n] per NODE, or NULL: bit 0 flips whenever the node's successor and predecessor trade places,
};                         //   i.e. whenever it lies in a reversed segment (the delta cache of pruned.cuh; apply_move keeps it)

struct DescState {
    long long moves;       // moves applied since upload
    long long cost;        // current tour cost (plain integer, no offset)
    long long log_count;   // entries written to the move log
    int status;            // 0 running / move cap hit, 1 local optimum (no admissible move)
    int last_a, last_b, last_delta;
    long long candidates;  // candidate nodes examined by the pruned scan (pruned.cuh)
    long long phase_sum;   // -DTSP_PHASE_TIMING: sum over warps and scans of the phase-1 duration
    long long phase_ns[6]; // -DTSP_PHASE_TIMING: ns in scan phase 1 / phase 2 / apply (each incl. its barrier), -, units, boxes
    long long nbstat[8];   // -DTSP_NB_STATS: delta-cache events (rebuilds, queued boxes, long lists, dirty, hits, evaluated, attend calls, box nodes)
};
#ifdef TSP_NB_STATS
#define NBSTAT(slot, v) atomicAdd((unsigned long long *)&g_nbstat[slot], (unsigned long long)(v))
__device__ unsigned long long g_nbstat[8];
#else
#define NBSTAT(slot, v)
#endif

struct MoveRec { int a, b, delta, reserved; };

// number of pair columns before strip s:  sum_{t<s} (n - 2 - t*SR)
__device__ __forceinline__ long long strip_prefix(int s, int n) {
    return (long long)s * (n - 2) - (long long)SR * ((long long)s * (s - 1) / 2);
}

// ---------------------------------------------------------------------------------------------
// One warp sweeps pair columns [j0, j1] of strip s (pair rows i0 .. i0+SR-1, i0 = s*SR).
// Lane l owns D rows i0 + l*R + r.  Step for column c computes D(row, c) for the R rows (R edge
// costs), fetches D(next lane's first row, c) by shuffle, and closes the R pairs of column c-1.
// ---------------------------------------------------------------------------------------------
template <class M, bool MASKED>
__device__ __forceinline__ void sweep_block(const TourView &V, int i0, int c, int cnt, const double (&rx)[R],
                                            const double (&ry)[R], const int (&base)[R], int (&Dp)[R],
                                            const double2 *sxy, const int *sdn, int &best,
                                            unsigned long long &bestkey, int *blk_best) {
    const int lane = threadIdx.x & 31;
    const int n = V.n;
#pragma unroll UNROLL
    for (int k = 0; k < cnt; ++k) {
        const double2 cj = sxy[k];
        const int dj = sdn[k];
        int Dn[R + 1];
        static_assert(R % 2 == 0, "rows are processed in pairs (packed FP32x2 cost sequence)");
#pragma unroll
        for (int r = 0; r < R; r += 2) M::K2(rx[r], ry[r], rx[r + 1], ry[r + 1], cj.x, cj.y, Dn[r], Dn[r + 1]);
        Dn[R] = __shfl_down_sync(FULL, Dn[0], 1);
        int m[R];
#pragma unroll
        for (int r = 0; r < R; ++r) m[r] = (int)((unsigned)Dp[r] + (unsigned)Dn[r + 1] + (unsigned)base[r]);
        if (MASKED) {
            const int j = c + k - 1;
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (j - (i0 + lane * R + r) < 2) m[r] += BIGPEN;
        }
        int mm = m[0];
#pragma unroll
        for (int r = 1; r < R; ++r) mm = min(mm, m[r]);
        mm -= dj;
        if (__any_sync(FULL, mm <= best)) {
            // rare path: a pair at least as good as the warp's running best
            const int j = c + k - 1;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int d = m[r] - dj;
                const int i = i0 + lane * R + r;
                if (d <= best && !(i == 0 && j == n - 1)) {      // (0,n-1): a == succ_b, refinment.c:55
                    unsigned long long key = pack_key(d, V.tid[i], V.tid[j]);
                    if (key < bestkey) bestkey = key;
                }
            }
            int mine = (bestkey == KEY_EMPTY) ? best : min(best, key_delta(bestkey));
            best = __reduce_min_sync(FULL, mine);
            if (lane == 0) atomicMin(blk_best, best);
        }
#pragma unroll
        for (int r = 0; r < R; ++r) Dp[r] = Dn[r];
    }
}

template <class M>
__device__ void sweep_strip(const TourView &V, int s, int j0, int j1, int &best, unsigned long long &bestkey,
                            double2 *sxy, int *sdn, int *blk_best) {
    const int lane = threadIdx.x & 31;
    const int n = V.n;
    const int i0 = s * SR;
    double rx[R], ry[R];
    int base[R], Dp[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = i0 + lane * R + r;
        const int ic = min(i, n);
        const double2 p = V.pxy[ic];
        rx[r] = p.x;
        ry[r] = p.y;
        const bool valid = (i <= n - 3) && !(lane == 31 && r == R - 1);
        base[r] = (int)(0u - (unsigned)V.dns[ic]) + (valid ? 0 : BIGPEN);
    }
    {   // warm-up: D(row, j0)
        const double2 c0 = V.pxy[j0];
#pragma unroll
        for (int r = 0; r < R; r += 2) M::K2(rx[r], ry[r], rx[r + 1], ry[r + 1], c0.x, c0.y, Dp[r], Dp[r + 1]);
    }
    int c = j0 + 1;
    while (c <= j1 + 1) {
        const int cnt = min(32, j1 + 2 - c);
        __syncwarp();
        sxy[lane] = V.pxy[min(c + lane, n)];
        sdn[lane] = V.dns[min(c + lane - 1, n)];
        __syncwarp();
        best = min(best, *(volatile int *)blk_best);
        if ((c - 1) < i0 + 32 * R + 1)
            sweep_block<M, true>(V, i0, c, cnt, rx, ry, base, Dp, sxy, sdn, best, bestkey, blk_best);
        else
            sweep_block<M, false>(V, i0, c, cnt, rx, ry, base, Dp, sxy, sdn, best, bestkey, blk_best);
        c += cnt;
    }
}

// Scan the pair columns [lo, hi) of the linearised (strip, column) space.
template <class M>
__device__ void scan_range(const TourView &V, long long lo, long long hi, int nstrips, int &best,
                           unsigned long long &bestkey, double2 *sxy, int *sdn, int *blk_best) {
    const int n = V.n;
    long long idx = lo;
    while (idx < hi) {
        int s_lo = 0, s_hi = nstrips - 1;
        while (s_lo < s_hi) {
            const int mid = (s_lo + s_hi + 1) >> 1;
            if (strip_prefix(mid, n) <= idx) s_lo = mid; else s_hi = mid - 1;
        }
        const int s = s_lo;
        const int j0 = s * SR + 2 + (int)(idx - strip_prefix(s, n));
        const long long rem = hi - idx;
        const int j1 = (int)min((long long)(n - 1), (long long)j0 + rem - 1);
        sweep_strip<M>(V, s, j0, j1, best, bestkey, sxy, sdn, blk_best);
        idx += (j1 - j0 + 1);
    }
}

__device__ __forceinline__ unsigned long long warp_min_key(unsigned long long k) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        unsigned long long o = __shfl_down_sync(FULL, k, off);
        k = (o < k) ? o : k;
    }
    return k;
}

// ---------------------------------------------------------------------------------------------
// Apply move (a,b): reverse the forward segment succ(a) .. b exactly as ref_reverse_path does
// (refinment.c:95-114: path[a]=b, path[succ_a]=succ_b, orientation of the rest preserved).
// Work is split over `nthreads` cooperating threads (a whole grid or one CTA); thread `t`.
// Each array element is read and written by exactly one thread, so no barrier is needed inside.
// ---------------------------------------------------------------------------------------------
// pos[] of the two segment END nodes (b and succ_a) is NOT written here: every cooperating thread
// reads pos[a], pos[b] on entry, so those two entries must stay stable until all threads have
// passed that read.  The caller writes them after its next barrier: pos[b] = s_out, and
// pos[tid[e_out]] = e_out (tid[e_out] is succ_a after the reversal).
// ep (delta cache only, V.o8 != NULL): the epoch of this move; the o8 bytes of the four nodes whose tour edges change
// (a, succ a, b, succ b) are stamped with it in bits 1..7 (pruned.cuh).
#ifndef TSP_AM_B
#define TSP_AM_B 4
#endif
template <class M, bool CACHE = false, int AM_B = 1>
__device__ __forceinline__ void apply_move(const TourView &V, int a, int b, int t, int nthreads, bool has_dns,
                                           int &s_out, int &e_out, int ep = 0) {
    const int n = V.n;
    const int i = V.pos[a], j = V.pos[b];
    const int s = (i + 1 == n) ? 0 : i + 1;           // position of succ(a)
    const int e = j;                                   // position of b
    int L = j - i; if (L < 0) L += n;                  // nodes in the segment succ(a)..b
    const int half = L >> 1;
    const int ehalf = (L - 1) >> 1;                    // interior edges: L-1 of them, mirrored
    s_out = s; e_out = e;
    if (t == 0) {
        // the two new tour edges: (a,b) at position i, (succ_a, succ_b) at position e
        const double2 pa = V.pxy[i], psa = V.pxy[s], pb = V.pxy[e], psb = V.pxy[e + 1];
        const int k_ab = M::K(pa.x, pa.y, pb.x, pb.y);
        const int k_ss = M::K(psa.x, psa.y, psb.x, psb.y);
        V.dn[i] = k_ab; V.dn[e] = k_ss;
        if (has_dns) { V.dns[i] = k_ab; V.dns[e] = k_ss; }
        if (i == 0) { V.dn[n] = k_ab; if (has_dns) V.dns[n] = k_ab; }
        if (e == 0) { V.dn[n] = k_ss; if (has_dns) V.dns[n] = k_ss; }
    }
    // AM_B pairs per thread and trip, all loads of a trip -- the o8 bytes included -- before its stores: when few threads share
    // a long segment (one CTA per descent: pla85900 reverses 48 000 nodes per move on average, 24 pairs per thread) a pair at
    // a time is a chain of dependent memory round trips per pair.  No pair is touched by two threads or twice by one, so the
    // order of loads and stores within a trip is free.  K8 passes AM_B = 4 (pla85900 318 -> 300, d15112 139 -> 106 us per scan);
    // where the move is spread over the whole grid (at most one pair per thread) 1 is best, and batches cost registers
    // (K7 pla85900 0.109 s with 1, 0.117 with 2, 0.126 with 4: profiles/r02_pruned.md).
    for (int k0 = t; k0 < half; k0 += AM_B * nthreads) {
        double2 xp[AM_B], xq[AM_B];
        int tp[AM_B], tq[AM_B], pp[AM_B], qq[AM_B], op[AM_B], oq[AM_B];
#pragma unroll
        for (int j = 0; j < AM_B; ++j) {
            const int k = k0 + j * nthreads;
            pp[j] = -1;
            if (k < half) {
                int p = s + k; if (p >= n) p -= n;
                int q = e - k; if (q < 0) q += n;
                pp[j] = p; qq[j] = q;
                xp[j] = V.pxy[p]; xq[j] = V.pxy[q];
                tp[j] = V.tid[p]; tq[j] = V.tid[q];
            }
        }
        if (CACHE) {
#pragma unroll
            for (int j = 0; j < AM_B; ++j)
                if (pp[j] >= 0) { oq[j] = V.o8[tq[j]]; op[j] = V.o8[tp[j]]; }
        }
#pragma unroll
        for (int j = 0; j < AM_B; ++j) {
            if (pp[j] < 0) continue;
            const int k = k0 + j * nthreads, p = pp[j], q = qq[j];
            V.pxy[p] = xq[j]; V.pxy[q] = xp[j];
            V.tid[p] = tq[j]; V.tid[q] = tp[j];
            if (k != 0) { V.pos[tq[j]] = p; V.pos[tp[j]] = q; }   // k == 0: the end nodes, deferred (see above)
            if (p == 0) { V.pxy[n] = xq[j]; V.tid[n] = tq[j]; }
            if (q == 0) { V.pxy[n] = xp[j]; V.tid[n] = tp[j]; }
            if (CACHE) {                                // one owner per node: plain byte updates
                if (k == 0) {                           // the segment ends: succ a and b
                    V.o8[tq[j]] = (unsigned char)(((oq[j] ^ 1) & 1) | (ep << 1));
                    V.o8[tp[j]] = (unsigned char)(((op[j] ^ 1) & 1) | (ep << 1));
                } else { V.o8[tq[j]] = (unsigned char)(oq[j] ^ 1); V.o8[tp[j]] = (unsigned char)(op[j] ^ 1); }
            }
        }
    }
    // The three single-thread jobs of a move -- the new edge costs and the end pair (thread 0 above), the stamps of a and
    // succ b, the middle node -- are chains of dependent loads: three different threads (of different warps where there are
    // any) take them, so that the longest of them, not their sum, is what the barrier after the move waits for.  a, succ b
    // (outside the segment) and the middle node are touched by nobody else.
    if (CACHE && t == nthreads - 1) {                   // a and succ b keep their direction
        const int sb = V.tid[e + 1];
        V.o8[a] = (unsigned char)((V.o8[a] & 1) | (ep << 1));
        V.o8[sb] = (unsigned char)((V.o8[sb] & 1) | (ep << 1));
    }
    if (CACHE && (L & 1) && t == (nthreads > 32 ? nthreads - 33 : 0)) {   // the middle node of an odd segment stays in place but turns round too
        int m = s + half; if (m >= n) m -= n;
        V.o8[V.tid[m]] ^= 1;
    }
    for (int k0 = t; k0 < ehalf; k0 += AM_B * nthreads) {
        int dp[AM_B], dq[AM_B], pp[AM_B], qq[AM_B];
#pragma unroll
        for (int j = 0; j < AM_B; ++j) {
            const int k = k0 + j * nthreads;
            pp[j] = -1;
            if (k < ehalf) {
                int p = s + k; if (p >= n) p -= n;
                int q = e - 1 - k; if (q < 0) q += n;
                pp[j] = p; qq[j] = q;
                dp[j] = V.dn[p]; dq[j] = V.dn[q];
            }
        }
#pragma unroll
        for (int j = 0; j < AM_B; ++j) {
            if (pp[j] < 0) continue;
            const int p = pp[j], q = qq[j];
            V.dn[p] = dq[j]; V.dn[q] = dp[j];
            if (has_dns) { V.dns[p] = dq[j]; V.dns[q] = dp[j]; }
            if (p == 0) { V.dn[n] = dq[j]; if (has_dns) V.dns[n] = dq[j]; }
            if (q == 0) { V.dn[n] = dp[j]; if (has_dns) V.dns[n] = dp[j]; }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K2: whole-GPU persistent descent (cooperative launch, grid = resident CTAs).
//   loop: scan -> 64-bit atomicMin -> grid.sync -> apply best move (all CTAs) -> grid.sync
// init_best = -1  : 2-opt, only improving moves (best_delta < -1e-7 on integer deltas)
// init_best = BIGPEN/2 : tabu, best admissible move even if worsening (metaheuristic.c:189,225)
// ---------------------------------------------------------------------------------------------
template <class M, int BLK>
__global__ void __launch_bounds__(BLK, (BLK >= 512 ? 1 : 1024 / BLK))
descend_kernel(TourView V, unsigned long long *key, DescState *st, MoveRec *log, long long log_cap,
               long long max_moves, int init_best, int dns_is_dn) {
    constexpr int WARPS = BLK / 32;
    cg::grid_group grid = cg::this_grid();
    __shared__ double2 s_xy[WARPS][32];
    __shared__ int s_dn[WARPS][32];
    __shared__ unsigned long long s_key;
    __shared__ int s_best;

    const int n = V.n;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nstrips = (n - 2 + SR - 1) / SR;
    const long long total = strip_prefix(nstrips, n);
    const long long gw = (long long)blockIdx.x * WARPS + warp;
    const long long W = (long long)gridDim.x * WARPS;
    const long long lo = total * gw / W, hi = total * (gw + 1) / W;
    const int gtid = blockIdx.x * BLK + threadIdx.x;
    const int gsize = gridDim.x * BLK;

    for (long long it = 0; it < max_moves; ++it) {
        if (threadIdx.x == 0) { s_key = KEY_EMPTY; s_best = init_best; }
        __syncthreads();
        int best = init_best;
        unsigned long long bestkey = KEY_EMPTY;
        scan_range<M>(V, lo, hi, nstrips, best, bestkey, s_xy[warp], s_dn[warp], &s_best);
        bestkey = warp_min_key(bestkey);
        if (lane == 0 && bestkey != KEY_EMPTY) atomicMin(&s_key, bestkey);
        __syncthreads();
        unsigned long long *kslot = key + (it & 1);
        if (threadIdx.x == 0 && s_key != KEY_EMPTY) atomicMin(kslot, s_key);
        grid.sync();
        // one reader per CTA, handed on through shared memory: thousands of warps loading the same word (volatile, so past
        // L1) queue up at one L2 slice for microseconds (profiles/r02_pruned.md).  s_key is re-armed behind a later barrier.
        if (threadIdx.x == 0) s_key = *(volatile unsigned long long *)kslot;
        __syncthreads();
        const unsigned long long kbest = s_key;
        if (kbest == KEY_EMPTY) {
            if (gtid == 0) st->status = 1;
            break;
        }
        const int a = key_a(kbest), b = key_b(kbest), delta = key_delta(kbest);
        int seg_s, seg_e;
        apply_move<M>(V, a, b, gtid, gsize, dns_is_dn == 0, seg_s, seg_e);
        if (gtid == 0) {
            key[(it + 1) & 1] = KEY_EMPTY;
            const long long lc = st->log_count;
            if (lc < log_cap) { log[lc].a = a; log[lc].b = b; log[lc].delta = delta; log[lc].reserved = 0; }
            st->log_count = lc + 1;
            st->moves += 1;
            st->cost += delta;
            st->last_a = a; st->last_b = b; st->last_delta = delta;
        }
        grid.sync();
        if (gtid == 0) { V.pos[b] = seg_s; V.pos[V.tid[seg_e]] = seg_e; }   // deferred end-node positions
    }
}

// ---------------------------------------------------------------------------------------------
// K5b: device-resident tabu search, the loop of mh_TabuSearch (metaheuristic.c:128-164) around
// tabu_best_move (:188-245) in ONE cooperative launch: per iteration  flags -> scan -> move -> stamp.
// The tenure of every iteration comes from the host (the reference's own policy functions,
// metaheuristic.c:7-59, computed ahead: they do not depend on the search state).  Every thread
// keeps the running cost and the best cost in registers (same deterministic updates), so the
// "new incumbent" decision (strict '<', tsp.c:671) needs no extra communication.
// ---------------------------------------------------------------------------------------------
struct TabuCtl {
    int *tabu_list;            // [n] last iteration in which the node took part in a move, -1 never
    const int *tenure;         // [iters] tenure in force at iteration first_iter + k
    long long *trace;          // [iters] tour cost after iteration k (results/TabuResults.dat)
    int *best_tid;             // [n+1] tour-position order of the best tour seen (valid if improved)
    long long cost, best_cost; // in: current tour cost, incumbent cost
    int first_iter, pad;
};
struct TabuOut { long long cost, best_cost, moves, done; int best_iter, last_delta; };   // done = iterations completed by the launch

template <class M, int BLK>
__global__ void __launch_bounds__(BLK, (BLK >= 512 ? 1 : 1024 / BLK))
tabu_kernel(TourView V, unsigned long long *key, TabuCtl T, TabuOut *out, long long iters) {
    constexpr int WARPS = BLK / 32;
    cg::grid_group grid = cg::this_grid();
    __shared__ double2 s_xy[WARPS][32];
    __shared__ int s_dn[WARPS][32];
    __shared__ unsigned long long s_key;
    __shared__ int s_best;

    const int n = V.n;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nstrips = (n - 2 + SR - 1) / SR;
    const long long total = strip_prefix(nstrips, n);
    const long long gw = (long long)blockIdx.x * WARPS + warp;
    const long long W = (long long)gridDim.x * WARPS;
    const long long lo = total * gw / W, hi = total * (gw + 1) / W;
    const int gtid = blockIdx.x * BLK + threadIdx.x;
    const int gsize = gridDim.x * BLK;
    // running cost / best cost / counters live in shared memory (one replica per CTA, same
    // deterministic updates everywhere) so that they hold no registers during the scan
    __shared__ long long s_cost, s_bestcost, s_moves;
    __shared__ int s_bestiter;
    if (threadIdx.x == 0) { s_cost = T.cost; s_bestcost = T.best_cost; s_moves = 0; s_bestiter = -1; }

    for (long long k = 0; k < iters; ++k) {
        const int iter = T.first_iter + (int)k;
        const int tenure = T.tenure[k];
        // dns[p] = dn[p] - BIGPEN where node(p) or node(p+1) is tabu (metaheuristic.c:208, :416-418)
        for (int p = gtid; p <= n; p += gsize) {
            const int u = V.tid[p], v = V.tid[p == n ? 1 : p + 1];
            const int tu = T.tabu_list[u], tv = T.tabu_list[v];
            const bool f = (iter - tu < tenure && tu != -1) || (iter - tv < tenure && tv != -1);
            V.dns[p] = V.dn[p] - (f ? BIGPEN : 0);
        }
        if (threadIdx.x == 0) { s_key = KEY_EMPTY; s_best = BIGPEN / 2; }
        grid.sync();
        {
            int best = BIGPEN / 2;                   // best admissible move even if worsening (:189)
            unsigned long long bestkey = KEY_EMPTY;
            scan_range<M>(V, lo, hi, nstrips, best, bestkey, s_xy[warp], s_dn[warp], &s_best);
            bestkey = warp_min_key(bestkey);
            if (lane == 0 && bestkey != KEY_EMPTY) atomicMin(&s_key, bestkey);
        }
        __syncthreads();
        unsigned long long *kslot = key + (k & 1);
        if (threadIdx.x == 0 && s_key != KEY_EMPTY) atomicMin(kslot, s_key);
        grid.sync();
        // one reader per CTA, handed on through shared memory: thousands of warps loading the same word (volatile, so past
        // L1) queue up at one L2 slice for microseconds (profiles/r02_pruned.md).  s_key is re-armed behind a later barrier.
        if (threadIdx.x == 0) s_key = *(volatile unsigned long long *)kslot;
        __syncthreads();
        const unsigned long long kbest = s_key;
        int seg_s = -1, seg_e = -1, b = -1;
        if (kbest != KEY_EMPTY) {                    // every pair tabu: no move this iteration (:225)
            const int a = key_a(kbest);
            b = key_b(kbest);
            if (gtid == 0) {                         // stamp a, b and their successors BEFORE the move (:235-238)
                const int i = V.pos[a], j = V.pos[b];
                const int sa = V.tid[i + 1], sb = V.tid[j + 1];
                T.tabu_list[a] = iter; T.tabu_list[b] = iter; T.tabu_list[sa] = iter; T.tabu_list[sb] = iter;
            }
            apply_move<M>(V, a, b, gtid, gsize, false, seg_s, seg_e);
            if (threadIdx.x == 0) { s_cost += key_delta(kbest); s_moves += 1; }
        }
        if (gtid == 0) { key[(k + 1) & 1] = KEY_EMPTY; T.trace[k] = s_cost; }
        grid.sync();
        if (b >= 0 && gtid == 0) { V.pos[b] = seg_s; V.pos[V.tid[seg_e]] = seg_e; }
        const bool improved = s_cost < s_bestcost;   // tsp_update_best_solution, strict '<' (tsp.c:671)
        if (improved)
            for (int p = gtid; p <= n; p += gsize) T.best_tid[p] = V.tid[p];
        __syncthreads();
        if (improved && threadIdx.x == 0) { s_bestcost = s_cost; s_bestiter = iter; }
    }
    __syncthreads();
    if (gtid == 0) { out->cost = s_cost; out->best_cost = s_bestcost; out->moves = s_moves; out->best_iter = s_bestiter; out->done = iters; out->last_delta = 0; }
}

// ---------------------------------------------------------------------------------------------
// K6: one CTA per descent, whole tour resident in shared memory (n <= CTA_MAX_N).  Used for
// multi-start work on small instances (h_greedy_2opt on pr2392-class inputs), where one scan is a
// few microseconds of whole-GPU work and a grid-wide barrier per move would dominate: here a move
// costs three __syncthreads.  CTAs pull starts from an atomic counter; up to three CTAs share an SM.
// The scan and the move are the SAME device functions as in descend_kernel (scan_range,
// apply_move), only pointed at shared memory, so results are identical by construction.
// ---------------------------------------------------------------------------------------------
constexpr int CTA_MAX_N = 7600;          // 28 B/node + staging must fit 227 KB
constexpr int CTA_SMALL_N = 3600;        // up to here two 512-thread CTAs share an SM; above, one of 1024

__host__ __device__ inline size_t cta_smem_bytes(int n, int CTA_WARPS) {
    const size_t n1 = (size_t)n + 1;
    size_t b = 16 * n1;                  // pxy
    b += 4 * n1 * 2;                     // tid, dn
    b += 4 * (size_t)n;                  // pos
    b = (b + 15) & ~(size_t)15;
    b += (size_t)CTA_WARPS * 32 * (16 + 4);   // per-warp column staging
    return b + 64;
}

template <class M, int CTA_BLOCK>
__global__ void __launch_bounds__(CTA_BLOCK, 1024 / CTA_BLOCK)
descend_cta_kernel(const double2 *__restrict__ xy, int n, const int *__restrict__ order_all, int nstarts,
                   int *next_start, long long *final_cost, long long *moves_out, int *path_all, int *status_out, const int *stop) {
    constexpr int CTA_WARPS = CTA_BLOCK / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned long long s_key;
    __shared__ int s_best, s_cur, s_stop;
    __shared__ long long s_cost;
    const size_t n1 = (size_t)n + 1;
    TourView V;
    V.n = n;
    V.pxy = reinterpret_cast<double2 *>(smem_raw);
    V.tid = reinterpret_cast<int *>(smem_raw + 16 * n1);
    V.dn = V.tid + n1;
    V.dns = V.dn;
    V.o8 = nullptr;
    V.pos = V.dn + n1;
    size_t off = (16 * n1 + 4 * n1 * 2 + 4 * (size_t)n + 15) & ~(size_t)15;
    double2 *stage_xy = reinterpret_cast<double2 *>(smem_raw + off);
    int *stage_dn = reinterpret_cast<int *>(smem_raw + off + (size_t)CTA_WARPS * 32 * 16);

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int nstrips = (n - 2 + SR - 1) / SR;
    const long long total = strip_prefix(nstrips, n);
    const long long lo = total * warp / CTA_WARPS, hi = total * (warp + 1) / CTA_WARPS;

    // stop (may be NULL): raised by the host when the caller's time limit has passed; looked at before a start is taken
    // (heuristics.c:83-90) and before every scan (refinment.c:17-24).  status_out: 1 = descent complete, 2 = cut.
    if (t == 0) s_stop = 0;
    for (;;) {
        if (t == 0) s_cur = (stop != nullptr && *(volatile const int *)stop != 0) ? nstarts : atomicAdd(next_start, 1);
        __syncthreads();
        const int cur = s_cur;
        if (cur >= nstarts) break;
        const int *order = order_all + (size_t)cur * n;
        for (int p = t; p <= n; p += CTA_BLOCK) {
            const int node = order[p == n ? 0 : p];
            V.pxy[p] = xy[node];
            V.tid[p] = node;
            if (p < n) V.pos[node] = p;
        }
        if (t == 0) s_cost = 0;
        __syncthreads();
        long long local = 0;
        for (int p = t; p < n; p += CTA_BLOCK) {
            const double2 u = V.pxy[p], v = V.pxy[p + 1];
            const int k = M::K(u.x, u.y, v.x, v.y);
            V.dn[p] = k;
            if (p == 0) V.dn[n] = k;
            local += (long long)(k - KOFF);
        }
        for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(FULL, local, o);
        if (lane == 0) atomicAdd((unsigned long long *)&s_cost, (unsigned long long)local);
        long long moves = 0;
        bool cut = false;
        for (;;) {
            if (t == 0) { s_key = KEY_EMPTY; s_best = -1; }
            __syncthreads();
            int best = -1;
            unsigned long long bestkey = KEY_EMPTY;
            scan_range<M>(V, lo, hi, nstrips, best, bestkey, stage_xy + warp * 32, stage_dn + warp * 32, &s_best);
            bestkey = warp_min_key(bestkey);
            if (lane == 0 && bestkey != KEY_EMPTY) atomicMin(&s_key, bestkey);
            __syncthreads();
            const unsigned long long kbest = s_key;
            if (kbest == KEY_EMPTY) break;                      // local optimum (refinment.c:27)
            const int a = key_a(kbest), b = key_b(kbest);
            int seg_s, seg_e;
            apply_move<M>(V, a, b, t, CTA_BLOCK, false, seg_s, seg_e);
            if (t == 0) {
                s_cost += key_delta(kbest);
                if (stop != nullptr) s_stop = *(volatile const int *)stop;
            }
            ++moves;
            __syncthreads();
            if (t == 0) { V.pos[b] = seg_s; V.pos[V.tid[seg_e]] = seg_e; }
            if (stop != nullptr && s_stop != 0) { cut = true; break; }
        }
        if (t == 0) { final_cost[cur] = s_cost; moves_out[cur] = moves; status_out[cur] = cut ? 2 : 1; }
        if (path_all) {
            int *path = path_all + (size_t)cur * n;
            for (int p = t; p < n; p += CTA_BLOCK) path[V.tid[p]] = V.tid[p + 1];
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Tour set-up / export kernels.
// ---------------------------------------------------------------------------------------------
// GEO instances: DDD.MM (latitude, longitude) -> radians, in place, once per instance.
__global__ void geo_convert_kernel(double2 *xy, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double2 p = xy[i];
        xy[i] = make_double2(geo_radians(p.x), geo_radians(p.y));
    }
}

// order[p] = node at position p (host walked the successor array).  Fills pxy, tid, pos.
__global__ void tour_init_kernel(TourView V, const double2 *xy, const int *order) {
    const int n = V.n;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p <= n; p += gridDim.x * blockDim.x) {
        const int node = order[p == n ? 0 : p];
        V.pxy[p] = xy[node];
        V.tid[p] = node;
        if (p < n) V.pos[node] = p;
    }
}

// dn[p] = K(edge p,p+1); cost = sum of plain costs (block-reduced, atomically added).
template <class M>
__global__ void tour_edges_kernel(TourView V, DescState *st) {
    const int n = V.n;
    long long local = 0;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const double2 u = V.pxy[p], v = V.pxy[p + 1];
        const int k = M::K(u.x, u.y, v.x, v.y);
        V.dn[p] = k;
        if (p == 0) V.dn[n] = k;
        local += (long long)(k - KOFF);
    }
    for (int off = 16; off > 0; off >>= 1) local += __shfl_down_sync(FULL, local, off);
    if ((threadIdx.x & 31) == 0 && local != 0) atomicAdd((unsigned long long *)&st->cost, (unsigned long long)local);
}

// dns[p] = dn[p] - BIGPEN if node(p) or node(p+1) is tabu (metaheuristic.c:208, :416-418):
// a pair (i,j) is admissible iff none of a, succ a, b, succ b is tabu <=> dns unflagged at i and j.
__global__ void tabu_flags_kernel(TourView V, const int *tabu_list, int tenure, int iter) {
    const int n = V.n;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p <= n; p += gridDim.x * blockDim.x) {
        const int u = V.tid[p], v = V.tid[p == n ? 1 : p + 1];
        const int tu = tabu_list[u], tv = tabu_list[v];
        const bool f = (iter - tu < tenure && tu != -1) || (iter - tv < tenure && tv != -1);
        V.dns[p] = V.dn[p] - (f ? BIGPEN : 0);
    }
}

__global__ void tour_export_kernel(TourView V, int *path) {
    const int n = V.n;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x)
        path[V.tid[p]] = V.tid[p + 1];
}

// ---------------------------------------------------------------------------------------------
// K1: cost matrix (src/tsp.c:608-636).  Row-major n*n doubles, diagonal -1.0.  HBM-write bound:
// each thread produces two consecutive cells and issues one 16-byte store (flat index even, so
// the store is 16-byte aligned whatever the parity of n); grid-stride over flat cell pairs.
// ---------------------------------------------------------------------------------------------
// plain cost -> double without an I2F.F64 conversion: 2^52 + k has k in its low mantissa bits
__device__ __forceinline__ double cost_as_double(int koff) {
    return __hiloint2double(0x43300000, koff - KOFF) - 4503599627370496.0;
}

#ifndef TSP_K1_ROWS
#define TSP_K1_ROWS 8
#endif
#ifndef TSP_K1_BLOCK
#define TSP_K1_BLOCK 128
#endif
// A work unit is TSP_K1_ROWS rows x one CHUNK of columns, and a CTA does ONE unit (grid = number of units).  Several rows
// at a time, so that the column coordinates -- which every row needs and which do not fit L1 for large n -- are read from
// L2 once per row group instead of once per row.  Units are numbered chunk-fastest and the hardware hands CTAs out in
// index order, so the CTAs resident at any moment write one compact, advancing band of the matrix.  Rounds 1-2 dealt the
// same units out grid-stride to 16 CTAs per SM: those drift apart, the write front frays, and pla85900 (59 GB) stayed at
// 0.80-0.86 of the copy bandwidth whatever the occupancy; a persistent grid (lock-step CTAs) was worse still, 0.70.  One
// unit per CTA writes at the rate cudaMemset reaches on the same buffer (profiles/r02_k1.md: 7.3 TB/s, 1.11 x the
// measured copy bandwidth).  Rows of a group have the same 16-byte alignment: consecutive rows when n is even, every
// other row when n is odd (row i starts at cell i*n).  chunk_pairs: 16-byte cell pairs per unit and row (a multiple of
// the block size; 4 block-wide trips measured best).  TSP_K1_ROT: CTA b starts with trip b mod trips (measured +1-3 %).
#ifndef TSP_K1_MINB
#define TSP_K1_MINB 1
#endif
#ifndef TSP_K1_ROT
#define TSP_K1_ROT 1
#endif
// The two cells (row, column j) and (row, column j + 1) as doubles.  Metrics with a packed scan form compute both costs
// with FP32x2 instructions (K2 with the roles swapped: (xj - xa)^2 == (xa - xj)^2 bit for bit); the cost leaves the
// sequence as the bit pattern 0xCB000000 + cost, which one integer subtraction and the 2^52 trick turn into a double.
template <class M>
__device__ __forceinline__ double2 k1_two_cells(const double2 pr, const double2 p0, const double2 p1) {
    if constexpr (M::SCAN_XOR == SCAN_FLIP) {
        int k0, k1;
        M::K2(p0.x, p0.y, p1.x, p1.y, pr.x, pr.y, k0, k1);
        return make_double2(__hiloint2double(0x43300000, (int)((unsigned)k0 - ((unsigned)KOFF ^ SCAN_FLIP))) - 4503599627370496.0,
                            __hiloint2double(0x43300000, (int)((unsigned)k1 - ((unsigned)KOFF ^ SCAN_FLIP))) - 4503599627370496.0);
    } else {
        return make_double2(cost_as_double(M::K(pr.x, pr.y, p0.x, p0.y)), cost_as_double(M::K(pr.x, pr.y, p1.x, p1.y)));
    }
}

// One unit: RB rows x the column pairs [t0, t1).  The loop body has NO data-dependent control flow -- the diagonal is
// computed like any other cell (cost 0) and overwritten with -1.0 afterwards by the very thread that stored it (program
// order, same address) -- so the 2 * RB cost chains of a trip are independent straight-line code the scheduler can
// interleave.  (Round 1/2's loop selected -1.0 per cell: nvcc turned each select into a branch around the cost sequence,
// leaving 16 serial dependent chains per trip, 27 instructions per cell and `wait` as the top stall, r02_ncu.md.)
template <class M, bool FULL_GROUP>
__device__ __forceinline__ void k1_unit(const double2 *__restrict__ xy, int n, double *__restrict__ out, int first, int s, int off,
                                        int t0, int t1) {
    constexpr int RB = TSP_K1_ROWS;
    double2 pr[RB];
#pragma unroll
    for (int k = 0; k < RB; ++k) pr[k] = xy[FULL_GROUP ? first + k * s : min(first + k * s, n - 1)];
    double *base = out + (size_t)first * (size_t)n + off;
    const size_t rstride = (size_t)s * (size_t)n;
#if TSP_K1_ROT
    const int T = (t1 - t0 + TSP_K1_BLOCK - 1) / TSP_K1_BLOCK, rot = (int)(blockIdx.x % (unsigned)T);
    for (int m = 0; m < T; ++m) {
        int trip = m + rot; if (trip >= T) trip -= T;
        const int t = t0 + trip * TSP_K1_BLOCK + threadIdx.x;
        if (t >= t1) continue;
#else
    for (int t = t0 + threadIdx.x; t < t1; t += TSP_K1_BLOCK) {
#endif
        const double2 p0 = xy[off + 2 * t], p1 = xy[off + 2 * t + 1];
        double *cell = base + 2 * (size_t)t;
#pragma unroll
        for (int k = 0; k < RB; ++k) {
            const double2 v = k1_two_cells<M>(pr[k], p0, p1);
            if (FULL_GROUP || first + k * s < n)
                __stcs(reinterpret_cast<double2 *>(cell + k * rstride), v);   // written once: streaming store (plain / .wt / .cg stores: 0.57-0.72)
        }
    }
#pragma unroll
    for (int k = 0; k < RB; ++k) {                       // diagonal cells of this unit, each by the thread that wrote it above
        const int i = first + k * s, ti = (i - off) >> 1;
        if (i >= off && ti >= t0 && ti < t1 && ((ti - t0) % TSP_K1_BLOCK) == (int)threadIdx.x && (FULL_GROUP || i < n))
            out[(size_t)i * (size_t)n + i] = -1.0;
    }
}

template <class M>
__global__ void __launch_bounds__(TSP_K1_BLOCK, TSP_K1_MINB)
cost_matrix_kernel(const double2 *__restrict__ xy, int n, double *__restrict__ out, int chunk_pairs) {
    constexpr int RB = TSP_K1_ROWS;
    const int s = (n & 1) ? 2 : 1;                       // row stride inside a group
    const int span = RB * s;                             // consecutive rows covered by s groups
    const int ngroups = ((n + span - 1) / span) * s;
    const int nchunks = max(1, ((n >> 1) + chunk_pairs - 1) / chunk_pairs);  // covers (n - off) / 2 pairs for off = 0 and 1
    const long long nunits = (long long)ngroups * nchunks;
    for (long long u = blockIdx.x; u < nunits; u += gridDim.x) {
        const int g = (int)(u / nchunks), ch = (int)(u - (long long)g * nchunks);
        const int first = (g / s) * span + (g % s);      // rows first, first + s, ...
        const int off = (int)(((size_t)first * (size_t)n) & 1);
        const int npairs = (n - off) >> 1;
        const int t0 = ch * chunk_pairs, t1 = min(npairs, t0 + chunk_pairs);
        if (first + (RB - 1) * s < n) k1_unit<M, true>(xy, n, out, first, s, off, t0, t1);
        else k1_unit<M, false>(xy, n, out, first, s, off, t0, t1);
        if (ch == 0 && threadIdx.x < RB) {               // the unpaired first / last cell of each row
            const int k = threadIdx.x, i = first + k * s;
            if (i < n) {
                double *row = out + (size_t)i * (size_t)n;
                const double2 pi = xy[i];
                if (off) { const double2 p = xy[0]; row[0] = (i == 0) ? -1.0 : cost_as_double(M::K(pi.x, pi.y, p.x, p.y)); }
                if ((n - off) & 1) { const double2 p = xy[n - 1]; row[n - 1] = (i == n - 1) ? -1.0 : cost_as_double(M::K(pi.x, pi.y, p.x, p.y)); }
            }
        }
    }
}

__global__ void matrix_checksum_kernel(const double *m, unsigned long long cells, long long *sum) {
    long long local = 0;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < cells; t += stride)
        local += (long long)m[t];
    for (int off = 16; off > 0; off >>= 1) local += __shfl_down_sync(FULL, local, off);
    if ((threadIdx.x & 31) == 0) atomicAdd((unsigned long long *)sum, (unsigned long long)local);
}

// ---------------------------------------------------------------------------------------------
// K4: nearest neighbour, one CTA per start (heuristics.c:216-288).  The unvisited set is a
// compacted candidate list (swap-remove); the winner of a step is argmin (cost, node id), which
// is the reference's "ascending i, strict <" rule.  Output is the visit ORDER (= tour positions),
// the successor array is derived by tour_export_kernel.
// ---------------------------------------------------------------------------------------------
constexpr int NN_BLOCK = 512;

template <class M>
__global__ void __launch_bounds__(NN_BLOCK)
nn_kernel(const double2 *__restrict__ xy, int n, int first_start, int stride, double2 *cand_xy_all,
          int *cand_id_all, int *order_all, long long *cost_out) {
    __shared__ unsigned long long s_wkey[NN_BLOCK / 32];
    __shared__ int s_wslot[NN_BLOCK / 32];
    __shared__ unsigned long long s_key;
    __shared__ int s_slot;
    const int sidx = blockIdx.x;
    const int start = first_start + sidx * stride;
    double2 *cxy = cand_xy_all + (size_t)sidx * n;
    int *cid = cand_id_all + (size_t)sidx * n;
    int *order = order_all + (size_t)sidx * n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int k = tid; k < n - 1; k += NN_BLOCK) {
        const int node = k + (k >= start ? 1 : 0);
        cxy[k] = xy[node];
        cid[k] = node;
    }
    if (tid == 0) order[0] = start;
    __syncthreads();
    double2 cur = xy[start];
    long long cost = 0;
    int m = n - 1;
    for (int step = 1; step < n; ++step, --m) {
        unsigned long long bk = KEY_EMPTY;
        int bslot = -1;
        for (int k = tid; k < m; k += NN_BLOCK) {
            const double2 p = cxy[k];
            const int d = M::K(cur.x, cur.y, p.x, p.y) - KOFF;
            const unsigned long long key = ((unsigned long long)(unsigned)d << 32) | (unsigned)cid[k];
            if (key < bk) { bk = key; bslot = k; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const unsigned long long ok = __shfl_down_sync(FULL, bk, off);
            const int os = __shfl_down_sync(FULL, bslot, off);
            if (ok < bk) { bk = ok; bslot = os; }
        }
        if (lane == 0) { s_wkey[warp] = bk; s_wslot[warp] = bslot; }
        __syncthreads();
        if (warp == 0) {
            bk = (lane < NN_BLOCK / 32) ? s_wkey[lane] : KEY_EMPTY;
            bslot = (lane < NN_BLOCK / 32) ? s_wslot[lane] : -1;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const unsigned long long ok = __shfl_down_sync(FULL, bk, off);
                const int os = __shfl_down_sync(FULL, bslot, off);
                if (ok < bk) { bk = ok; bslot = os; }
            }
            if (lane == 0) {
                s_key = bk; s_slot = bslot;
                // swap-remove the winner; its coordinates become the new current point
                const int node = (int)(bk & 0xFFFFFFFFu);
                order[step] = node;
                const double2 last = cxy[m - 1];
                const int lastid = cid[m - 1];
                cxy[m - 1] = cxy[bslot];      // park the winner's coords where everyone can read them
                cxy[bslot] = last;
                cid[bslot] = lastid;
            }
        }
        __syncthreads();
        cost += (long long)(s_key >> 32);
        cur = cxy[m - 1];
    }
    if (tid == 0) {
        const double2 s0 = xy[start];
        long long closing = (n > 1) ? (long long)(M::K(cur.x, cur.y, s0.x, s0.y) - KOFF) : -1;   // heuristics.c:282
        cost_out[sidx] = cost + closing;
    }
}

// ---------------------------------------------------------------------------------------------
// Measurement / self-check kernels.
// ---------------------------------------------------------------------------------------------
template <int KIND>
__global__ void pipe_rate_kernel(int iters, double seed, double *sink) {
    double d[8];
    float f[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { d[k] = seed + k + threadIdx.x * 1e-3; f[k] = (float)d[k]; }
    const double c1 = 1.0000000001, c2 = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (KIND == 0) d[k] = __dadd_rn(d[k], c2);
            else if (KIND == 1) d[k] = __dmul_rn(d[k], c1);
            else if (KIND == 2) d[k] = __fma_rn(d[k], c1, c2);
            else if (KIND == 3) { f[k] = __double2float_rn(d[k]); d[k] = (double)f[k]; }
            else if (KIND == 4) { asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(f[k]) : "f"(f[k])); }
            else if (KIND == 5) f[k] = __fmaf_rn(f[k], 1.0000001f, 1e-7f);
            // 6 / 7: the whole edge-cost sequence; the result flips the low mantissa bit of its own input (one
            // integer instruction, no conversion) so that the chain cannot be hoisted
            else if (KIND == 6) { int kk = EucGeneral::K(d[k], d[(k + 1) & 7], 3.25, 7.5); d[k] = __hiloint2double(__double2hiint(d[k]), __double2loint(d[k]) ^ (kk & 1)); }
            else if (KIND == 8) { f[k] = __double2float_rn(d[k]); d[k] = __hiloint2double(__double2hiint(d[k]), __double2loint(d[k]) ^ (__float_as_int(f[k]) & 1)); }
            else if (KIND == 9) { int ka, kb; EucInt::K2(d[k], d[(k + 1) & 7], d[(k + 1) & 7], d[k], 3.0, 7.0, ka, kb); d[k] = __hiloint2double(__double2hiint(d[k]), __double2loint(d[k]) ^ ((ka ^ kb) & 1)); }
            else if (KIND == 10) { int ka, kb; EucIntF32::K2(f[k], f[(k + 1) & 7], f[(k + 1) & 7], f[k], 3.0f, 7.0f, ka, kb); f[k] = __int_as_float(__float_as_int(f[k]) ^ ((ka ^ kb) & 1)); }
            else if (KIND == 7) { int kk = EucInt::K(d[k], d[(k + 1) & 7], 3.0, 7.0); d[k] = __hiloint2double(__double2hiint(d[k]), __double2loint(d[k]) ^ (kk & 1)); }
        }
    }
    double acc = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += d[k] + (double)f[k];
    if (acc == 123.456) sink[0] = acc;
}

// EucIntF32 against EucInt over pseudo-random integer coordinates below 2^23 (measurement companion of pipe_rate kind 10)
__global__ void f32cost_selfcheck_kernel(unsigned seed, int trials, unsigned long long *mismatches) {
    unsigned long long bad = 0;
    unsigned s = seed ^ (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u;
    for (int it = 0; it < trials; ++it) {
        int v[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) { s = s * 1664525u + 1013904223u; v[k] = (int)(s >> 9); }       // 23 bits
        if ((it & 7) == 0) { v[2] = v[0] + (v[2] & 1023) - 512; v[3] = v[1] + (v[3] & 1023) - 512; v[4] = v[0]; v[5] = v[1]; }   // near / coincident
        int ka, kb, ra, rb;
        EucIntF32::K2((float)v[0], (float)v[1], (float)v[2], (float)v[3], (float)v[4], (float)v[5], ka, kb);
        EucInt::K2((double)v[0], (double)v[1], (double)v[2], (double)v[3], (double)v[4], (double)v[5], ra, rb);
        if (ka != ra) bad++;
        if (kb != rb) bad++;
    }
    if (bad) atomicAdd(mismatches, bad);
}

__global__ void sqrt_selfcheck_kernel(unsigned lo_bits, unsigned hi_bits, unsigned long long *mismatches) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long bad = 0;
    for (unsigned long long b = lo_bits + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b < hi_bits; b += stride) {
        const float f = __uint_as_float((unsigned)b);
        const float r0 = __fsqrt_rn(f);
        const float r1 = sqrt_rn_pos(f);
        if (__float_as_uint(r0) != __float_as_uint(r1)) bad++;
        if (b < ((127u + 46u) << 23)) {               // costs below 2^23: the packed FP32x2 sequence, both halves
            const float f2 = __uint_as_float((unsigned)b ^ 0x2A5A5u);
            int k0, k1;
            koff2_from_negsumsq(-f, -f2, k0, k1);
            if ((unsigned)k0 != ((unsigned)koff_from_sumsq(f) ^ SCAN_FLIP)) bad++;
            if ((unsigned)k1 != ((unsigned)koff_from_sumsq(f2) ^ SCAN_FLIP)) bad++;
        }
    }
    if (bad) atomicAdd(mismatches, bad);
}

}  // namespace tspcuda
