// tspcuda.cu -- host side of libtspcuda.so: the C ABI of include/tspcuda.h over the sm_100a
// kernels of kernels.cuh.  No CPU implementation of any algorithm lives here: without a CUDA
// device every compute entry point fails with TSPCUDA_UNAVAILABLE.
#include <cuda_runtime.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/tspcuda.h"
#include "kernels.cuh"
#include "pruned.cuh"
#include "matrix.cuh"
#include "mileage.cuh"
#include "patching.cuh"

using namespace tspcuda;

static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fail(e_ == cudaErrorMemoryAllocation ? TSPCUDA_RESOURCE_EXHAUSTED : TSPCUDA_INTERNAL, \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);   \
    } while (0)

enum MetricId { M_EUC_GENERAL = 0, M_EUC_INTEGER = 1, M_CEIL = 2, M_ATT = 3, M_GEO = 4 };

#define DISPATCH_METRIC(mid, EXPR)                                        \
    switch (mid) {                                                        \
        case M_EUC_GENERAL: { using M = EucGeneral; EXPR; } break;        \
        case M_EUC_INTEGER: { using M = EucInt; EXPR; } break;            \
        case M_CEIL: { using M = Ceil2D; EXPR; } break;                   \
        case M_GEO: { using M = Geo; EXPR; } break;                       \
        default: { using M = Att; EXPR; } break;                          \
    }

// dynamic shared memory of multi_pruned_kernel<M, blk, cl>: the stages of its heap stream (pruned.cuh, STG)
#if TSP_NB_STAGE
static size_t k8_dyn_smem(int blk, int cl, int) { return (blk / 32) * cl < NB_SPLIT_WARPS ? stage_smem_bytes(blk / 32) : 0; }
#else
// one CTA per descent: the n orientation / epoch bytes of the descent (pruned.cuh, K8_O8_SMEM); n <= 131 072
static size_t k8_dyn_smem(int, int cl, int n) { return (K8_O8_SMEM && cl == 1) ? (((size_t)n + 127) & ~(size_t)127) : 0; }
#endif

// Grow-only device scratch of the multi-start entry points, kept in the context between calls: cudaFree of the per-call
// buffers sporadically took 0.1 - 1.5 s on the B200 boxes (TSPCUDA_VERBOSE phase clock, profiles/r01_cluster.md), more
// than the descents of a small batch.  Released by tspcuda_destroy.
struct ScratchBuf { void *p = nullptr; size_t bytes = 0; };
enum {
    SC_NN_CXY, SC_NN_CID, SC_NN_ORDER, SC_NN_COST, SC_NN_SLOT, SC_BEST_PATH, SC_NEXT, SC_DCOST, SC_DMOVES, SC_DPATHS,
    SC_M_PXY, SC_M_TID, SC_M_DN, SC_M_POS, SC_M_BOXES, SC_M_BPATH, SC_M_BCOST, SC_M_BSTART, SC_M_CAND, SC_M_NB, SC_M_NBHEAP, SC_M_NBNEXT, SC_M_O8, SC_M_CNT8, SC_DSTATUS,
    SC_T_TEN, SC_T_TRACE, SC_T_BEST, SC_T_OUT, SC_NB_LIST, SC_COUNT
};

struct tspcuda_ctx {
    ScratchBuf scratch[SC_COUNT];
    int n = 0, wt = 0, device = 0, metric = 0;
    int num_sms = 0, max_coop_blocks = 0, block = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double2 *d_xy = nullptr;
    TourView tour{};
    int *d_path = nullptr, *d_order = nullptr, *d_tabu = nullptr;
    unsigned long long *d_key = nullptr;
    DescState *d_state = nullptr, *h_state = nullptr;
    MoveRec *d_log = nullptr;
    long long log_cap = 0;
    int *h_buf = nullptr;            // pinned, n ints
    double *d_matrix = nullptr;
    long long *d_sum = nullptr;
    bool tour_loaded = false;
    std::vector<int> h_cache;        // successor array the device tour currently represents
    tspcuda_stats stats{};
    double est_scan_s = 0;
    long long dev_moves = 0;         // host mirror of d_state->moves
    cudaEvent_t tev0 = nullptr, tev1 = nullptr;
    void *d_flush = nullptr;
    // K7 (pruned.cuh): static uniform grid over the node coordinates + the big-box queue
    SpatialGrid grid{};
    int *d_cell_start = nullptr, *d_cell_count = nullptr, *d_cnode = nullptr;
    double2 *d_cxy = nullptr;
    unsigned long long *d_qctr = nullptr;
    BigBox *d_boxes = nullptr;
    uint4 *d_wbuf = nullptr;         // per-warp record buffers of the pruned kernels (num_sms * 32 warps)
    int *d_nb = nullptr;             // delta cache of the resident tour (pruned.cuh: one 64-byte line per node,
    int4 *d_nb_heap = nullptr;       //   a record heap for the long lists,
    int *d_nb_next = nullptr;        //   the heap's counters,
    unsigned char *d_o8 = nullptr;   //   one orientation / epoch byte and one count byte per node)
    unsigned char *d_cnt8 = nullptr; //   and a byte-per-node flag array for nb_mark_*_kernel
    unsigned char *d_nbflag = nullptr;
    int nb_heap_per_node = NB_HEAP_PER_NODE;
    int nb_restart_pct = 100;             // NbCache::restart_pct (TSPCUDA_NB_RESTART_PCT, measurements)
    int nb_giant_age = 0;                 // NbCache::giant_age of the next descents (the VNS loop raises it)
    long long nb_moves = 0;          // moves the delta cache has followed since it was last invalidated (epoch numbering)
    bool nb_valid = false;           // false: the tour changed behind the cache's back (upload, full-scan or tabu moves)
    int *d_stop = nullptr;           // time-limit flag of the batched multi-start kernels, raised through aux_stream
    int *h_one = nullptr;            // pinned constant 1: source of that copy (a DMA copy needs no SM, a memset kernel might)
    cudaStream_t aux_stream = nullptr;
    int pruned_blocks = 0;           // cooperative grid of descend_pruned_kernel (0: unavailable)
    int k8_clusters[4] = {0, 0, 0, 0};   // co-resident K8 clusters of 1/2/4/8 CTAs (cudaOccupancyMaxActiveClusters), filled at first use
    int pruned_block = 0;            // its CTA size
    int scan_mode = TSPCUDA_SCAN_AUTO;
    double est_pruned_s = 0, est_full_s = 0;
    long long chunk_no = 0;
    bool probe_due = false;
    long long dev_candidates = 0;    // host mirror of d_state->candidates
};

extern "C" const char *tspcuda_last_error(void) { return g_err.c_str(); }
extern "C" const char *tspcuda_version(void) { return "tspcuda 0.1 (sm_100a)"; }

extern "C" int tspcuda_device_count(void) {
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) { cudaGetLastError(); return 0; }
    return c;
}

static int use(tspcuda_ctx *c) {
    if (!c) return fail(TSPCUDA_INVALID_ARGUMENT, "null context");
    CK(cudaSetDevice(c->device));
    return TSPCUDA_OK;
}

template <class M, int BLK>
static int coop_blocks_per_sm() {
    int b = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, descend_kernel<M, BLK>, BLK, 0);
    return b;
}

extern "C" void tspcuda_destroy(tspcuda_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    cudaFree(c->d_xy); cudaFree(c->tour.pxy); cudaFree(c->tour.tid); cudaFree(c->tour.dn);
    cudaFree(c->tour.dns); cudaFree(c->tour.pos); cudaFree(c->d_path); cudaFree(c->d_order);
    cudaFree(c->d_tabu); cudaFree(c->d_key); cudaFree(c->d_state); cudaFree(c->d_log);
    cudaFree(c->d_matrix); cudaFree(c->d_sum); cudaFree(c->d_flush);
    cudaFree(c->d_cell_start); cudaFree(c->d_cell_count); cudaFree(c->d_cnode); cudaFree(c->d_cxy);
    cudaFree(c->d_qctr); cudaFree(c->d_boxes); cudaFree(c->d_wbuf); cudaFree(c->d_nb); cudaFree(c->d_nb_heap); cudaFree(c->d_nb_next); cudaFree(c->d_o8); cudaFree(c->d_cnt8); cudaFree(c->d_nbflag); cudaFree(c->d_stop);
    if (c->h_one) cudaFreeHost(c->h_one);
    if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
    for (ScratchBuf &b : c->scratch) cudaFree(b.p);
    if (c->tev0) cudaEventDestroy(c->tev0);
    if (c->tev1) cudaEventDestroy(c->tev1);
    if (c->h_state) cudaFreeHost(c->h_state);
    if (c->h_buf) cudaFreeHost(c->h_buf);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" int tspcuda_create(tspcuda_ctx **out, const double *xy, int n, int weight_type, int device) {
    if (!out || !xy) return fail(TSPCUDA_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    if (n <= 0) return fail(TSPCUDA_INVALID_ARGUMENT, "n = %d", n);
    if (n > TSPCUDA_MAX_NODES) return fail(TSPCUDA_OUT_OF_RANGE, "n = %d exceeds %d (17-bit node ids in the packed key)", n, TSPCUDA_MAX_NODES);
    if (weight_type < TSPCUDA_EUC_REF || weight_type > TSPCUDA_GEO) return fail(TSPCUDA_UNIMPLEMENTED, "weight type %d", weight_type);
    int ndev = tspcuda_device_count();
    if (ndev <= 0) return fail(TSPCUDA_UNAVAILABLE, "no CUDA device: libtspcuda has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(TSPCUDA_INVALID_ARGUMENT, "device %d of %d", device, ndev);

    bool integer = true;
    double xmin = xy[0], xmax = xy[0], ymin = xy[1], ymax = xy[1];
    for (int i = 0; i < n; i++) {
        const double x = xy[2 * i], y = xy[2 * i + 1];
        if (!std::isfinite(x) || !std::isfinite(y)) return fail(TSPCUDA_INVALID_ARGUMENT, "non-finite coordinate at node %d", i);
        if (x != std::floor(x) || y != std::floor(y) || std::fabs(x) >= 33554432.0 || std::fabs(y) >= 33554432.0) integer = false;
        xmin = std::min(xmin, x); xmax = std::max(xmax, x); ymin = std::min(ymin, y); ymax = std::max(ymax, y);
    }
    const double diag = std::sqrt((xmax - xmin) * (xmax - xmin) + (ymax - ymin) * (ymax - ymin));
    if (diag >= 8388000.0 && weight_type != TSPCUDA_GEO)
        return fail(TSPCUDA_OUT_OF_RANGE, "bounding-box diagonal %.0f: edge costs must stay below 2^23 for the K-offset encoding", diag);

    tspcuda_ctx *c = new tspcuda_ctx();
    c->n = n; c->wt = weight_type; c->device = device;
    c->metric = weight_type == TSPCUDA_EUC_REF ? (integer ? M_EUC_INTEGER : M_EUC_GENERAL)
                : (weight_type == TSPCUDA_CEIL_2D ? M_CEIL : (weight_type == TSPCUDA_ATT ? M_ATT : M_GEO));
    if (const char *force = getenv("TSPCUDA_FORCE_METRIC")) {     // test hook: a SLOWER exact variant of the same rule
        if (weight_type == TSPCUDA_EUC_REF) {
            if (!strcmp(force, "general")) c->metric = M_EUC_GENERAL;
        }
    }
#define CKC(call)                                                                                     \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            int code_ = fail(e_ == cudaErrorMemoryAllocation ? TSPCUDA_RESOURCE_EXHAUSTED : TSPCUDA_INTERNAL, \
                             "%s failed: %s", #call, cudaGetErrorString(e_));                          \
            tspcuda_destroy(c);                                                                       \
            return code_;                                                                             \
        }                                                                                             \
    } while (0)
    CKC(cudaSetDevice(device));
    cudaDeviceProp prop;
    CKC(cudaGetDeviceProperties(&prop, device));
    if (!prop.cooperativeLaunch) { tspcuda_destroy(c); return fail(TSPCUDA_UNAVAILABLE, "device lacks cooperative launch"); }
    c->num_sms = prop.multiProcessorCount;
    int bps = 0;
    c->block = n >= HUGE_N ? BLOCK_HUGE : (n >= BIG_N ? BLOCK_BIG : BLOCK_SMALL);
    if (c->block == BLOCK_HUGE) { DISPATCH_METRIC(c->metric, (bps = coop_blocks_per_sm<M, BLOCK_HUGE>())); }
    else if (c->block == BLOCK_BIG) { DISPATCH_METRIC(c->metric, (bps = coop_blocks_per_sm<M, BLOCK_BIG>())); }
    else { DISPATCH_METRIC(c->metric, (bps = coop_blocks_per_sm<M, BLOCK_SMALL>())); }
    if (bps <= 0) { tspcuda_destroy(c); return fail(TSPCUDA_INTERNAL, "descent kernel cannot be resident (built for sm_100a; device is sm_%d%d)", prop.major, prop.minor); }
    c->max_coop_blocks = bps * c->num_sms;
    CKC(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CKC(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
    CKC(cudaMalloc(&c->d_stop, sizeof(int)));
    CKC(cudaMallocHost(&c->h_one, sizeof(int)));
    *c->h_one = 1;
    CKC(cudaEventCreate(&c->ev0));
    CKC(cudaEventCreate(&c->ev1));
    CKC(cudaEventCreate(&c->tev0));
    CKC(cudaEventCreate(&c->tev1));
    const size_t n1 = (size_t)n + 1;
    CKC(cudaMalloc(&c->d_xy, sizeof(double2) * n));
    CKC(cudaMalloc(&c->tour.pxy, sizeof(double2) * n1));
    CKC(cudaMalloc(&c->tour.tid, sizeof(int) * n1));
    CKC(cudaMalloc(&c->tour.dn, sizeof(int) * n1));
    CKC(cudaMalloc(&c->tour.dns, sizeof(int) * n1));
    CKC(cudaMalloc(&c->tour.pos, sizeof(int) * n));
    c->tour.n = n;
    CKC(cudaMalloc(&c->d_path, sizeof(int) * n));
    CKC(cudaMalloc(&c->d_order, sizeof(int) * n));
    CKC(cudaMalloc(&c->d_tabu, sizeof(int) * n));
    CKC(cudaMalloc(&c->d_key, sizeof(unsigned long long) * 2));
    CKC(cudaMalloc(&c->d_state, sizeof(DescState)));
    CKC(cudaMalloc(&c->d_sum, sizeof(long long)));
    c->log_cap = std::max<long long>(4096, 8LL * n);
    CKC(cudaMalloc(&c->d_log, sizeof(MoveRec) * c->log_cap));
    CKC(cudaMallocHost(&c->h_state, sizeof(DescState)));
    CKC(cudaMallocHost(&c->h_buf, sizeof(int) * n));
    CKC(cudaMemcpyAsync(c->d_xy, xy, sizeof(double2) * n, cudaMemcpyHostToDevice, c->stream));
    if (c->metric == M_GEO) {
        geo_convert_kernel<<<std::max(1, std::min(4 * c->num_sms, (n + 255) / 256)), 256, 0, c->stream>>>(c->d_xy, n);
        c->stats.kernel_launches += 1;
        CKC(cudaGetLastError());
    }
    CKC(cudaStreamSynchronize(c->stream));
    c->stats.h2d_bytes += (int64_t)sizeof(double2) * n;
    const double pairs = 0.5 * (double)n * (double)(n - 1);
    c->est_scan_s = std::max(pairs / 5e11, 8e-6);
    c->est_full_s = pairs / 1.0e12 + 8e-6;
    if (const char *sm = getenv("TSPCUDA_SCAN")) {
        if (!strcmp(sm, "full")) c->scan_mode = TSPCUDA_SCAN_FULL;
        else if (!strcmp(sm, "pruned")) c->scan_mode = TSPCUDA_SCAN_PRUNED;
    }
    {   // K7: uniform grid, about two nodes per cell, built on the device
        const double w = xmax - xmin, h = ymax - ymin;
        const double target = std::max(1.0, 0.5 * (double)n);
        double cell = (w > 0 && h > 0) ? std::sqrt(w * h / target) : std::max(w, h) / target;
        if (!(cell > 0)) cell = 1.0;
        cell = std::max(cell, std::max(w, h) / 2048.0);
        SpatialGrid &G = c->grid;
        G.gx0 = xmin; G.gy0 = ymin; G.inv_cell = 1.0 / cell;
        G.W = (int)std::min(2048.0, std::floor(w / cell) + 1.0);
        G.H = (int)std::min(2048.0, std::floor(h / cell) + 1.0);
        const int ncells = G.W * G.H;
        CKC(cudaMalloc(&c->d_cell_start, sizeof(int) * (ncells + 1)));
        CKC(cudaMalloc(&c->d_cell_count, sizeof(int) * ncells));
        CKC(cudaMalloc(&c->d_cnode, sizeof(int) * n));
        CKC(cudaMalloc(&c->d_cxy, sizeof(double2) * n));
        CKC(cudaMalloc(&c->d_qctr, sizeof(unsigned long long)));
        CKC(cudaMalloc(&c->d_boxes, sizeof(BigBox) * n));
        CKC(cudaMalloc(&c->d_wbuf, sizeof(uint4) * (size_t)PR_WBUF * 32 * c->num_sms));
        CKC(cudaMalloc(&c->d_nb, sizeof(int) * (size_t)NB_WORDS * n));
        if (const char *hv = getenv("TSPCUDA_NB_HEAP")) c->nb_heap_per_node = std::max(0, atoi(hv));
        if (const char *hv = getenv("TSPCUDA_NB_RESTART_PCT")) c->nb_restart_pct = std::max(1, std::min(100000, atoi(hv)));
        CKC(cudaMalloc(&c->d_nb_heap, sizeof(int4) * (size_t)nb_heap_cap(n, c->nb_heap_per_node)));
        CKC(cudaMalloc(&c->d_nb_next, sizeof(int) * 8));
        CKC(cudaMemsetAsync(c->d_nb_next, 0, sizeof(int) * 8, c->stream));
        CKC(cudaMalloc(&c->d_o8, n));
        CKC(cudaMalloc(&c->d_cnt8, nb_cnt8_stride(n)));
        CKC(cudaMalloc(&c->d_nbflag, n));
        CKC(cudaMemsetAsync(c->d_nbflag, 0, n, c->stream));
        G.cell_start = c->d_cell_start; G.cnode = c->d_cnode; G.cxy = c->d_cxy;
        CKC(cudaMemsetAsync(c->d_cell_count, 0, sizeof(int) * ncells, c->stream));
        CKC(cudaMemsetAsync(c->d_qctr, 0, sizeof(unsigned long long), c->stream));
        const int g = std::max(1, std::min(4 * c->num_sms, (n + 255) / 256));
        grid_count_kernel<<<g, 256, 0, c->stream>>>(G, c->d_xy, n, c->d_cell_count);
        grid_scan_kernel<<<1, 1024, 0, c->stream>>>(c->d_cell_count, c->d_cell_start, ncells);
        grid_fill_kernel<<<g, 256, 0, c->stream>>>(G, c->d_xy, n, c->d_cell_count, c->d_cnode, c->d_cxy);
        c->stats.kernel_launches += 3;
        CKC(cudaGetLastError());
        CKC(cudaStreamSynchronize(c->stream));
        // one CTA per SM (cheapest grid barrier); 1024-thread CTAs when 512 per SM would leave some
        // threads with two tour positions in phase 1
        c->pruned_block = (long long)c->num_sms * PRUNED_BLOCK >= n ? PRUNED_BLOCK : 1024;
        int pb = 0;
        if (c->pruned_block == PRUNED_BLOCK) { DISPATCH_METRIC(c->metric, (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pb, descend_pruned_kernel<M, PRUNED_BLOCK>, PRUNED_BLOCK, 0))); }
        else { DISPATCH_METRIC(c->metric, (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pb, descend_pruned_kernel<M, 1024>, 1024, 0))); }
        c->pruned_blocks = (pb > 0 && c->metric != M_GEO) ? c->num_sms : 0;     // the grid search is Euclidean
    }
    *out = c;
    return TSPCUDA_OK;
}

extern "C" int tspcuda_set_scan_mode(tspcuda_ctx *c, int mode) {
    if (!c) return fail(TSPCUDA_INVALID_ARGUMENT, "null context");
    if (mode < TSPCUDA_SCAN_AUTO || mode > TSPCUDA_SCAN_PRUNED) return fail(TSPCUDA_INVALID_ARGUMENT, "scan mode %d", mode);
    c->scan_mode = mode;
    return TSPCUDA_OK;
}

extern "C" int tspcuda_nnodes(const tspcuda_ctx *c) { return c ? c->n : -1; }

extern "C" int tspcuda_get_stats(tspcuda_ctx *c, tspcuda_stats *out) {
    if (!c || !out) return fail(TSPCUDA_INVALID_ARGUMENT, "null argument");
    *out = c->stats;
    return TSPCUDA_OK;
}
extern "C" int tspcuda_reset_stats(tspcuda_ctx *c) {
    if (!c) return fail(TSPCUDA_INVALID_ARGUMENT, "null context");
    c->stats = tspcuda_stats{};
    return TSPCUDA_OK;
}

static inline int grid_for(int n, int cap) { return std::max(1, std::min(cap, (n + 255) / 256)); }

// ---- tour slot -----------------------------------------------------------------------------
// order already in c->d_order (device): build pxy/tid/pos/dn and the cost.
static int tour_from_device_order(tspcuda_ctx *c, bool keep_nb = false) {
    const int n = c->n;
    const int g = grid_for(n + 1, 4 * c->num_sms);
    CK(cudaMemsetAsync(c->d_state, 0, sizeof(DescState), c->stream));
    c->dev_moves = 0;
    c->dev_candidates = 0;
    if (!keep_nb) c->nb_valid = false;
    tour_init_kernel<<<g, 256, 0, c->stream>>>(c->tour, c->d_xy, c->d_order);
    DISPATCH_METRIC(c->metric, (tour_edges_kernel<M><<<g, 256, 0, c->stream>>>(c->tour, c->d_state)));
    c->stats.kernel_launches += 2;
    CK(cudaGetLastError());
    return TSPCUDA_OK;
}

extern "C" int tspcuda_tour_upload(tspcuda_ctx *c, const int *path) {
    int e = use(c);
    if (e) return e;
    if (!path) return fail(TSPCUDA_INVALID_ARGUMENT, "null path");
    const int n = c->n;
    if (c->tour_loaded && (int)c->h_cache.size() == n && std::memcmp(c->h_cache.data(), path, sizeof(int) * n) == 0)
        return TSPCUDA_OK;      // the device tour already is this successor array
    // successor array -> visit order (data marshalling; also the single-cycle check of tsp.c:642-728)
    int *order = c->h_buf;
    int cur = 0;
    for (int k = 0; k < n; k++) {
        if (cur < 0 || cur >= n) return fail(TSPCUDA_INVALID_ARGUMENT, "path entry out of range");
        order[k] = cur;
        cur = path[cur];
    }
    if (cur != 0) return fail(TSPCUDA_INVALID_ARGUMENT, "path is not a single cycle through all nodes");
    {   // permutation check
        std::vector<unsigned char> seen(n, 0);
        for (int k = 0; k < n; k++) {
            if (seen[order[k]]) return fail(TSPCUDA_INVALID_ARGUMENT, "path is not a single cycle through all nodes");
            seen[order[k]] = 1;
        }
    }
    CK(cudaMemcpyAsync(c->d_order, order, sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
    c->stats.h2d_bytes += (int64_t)sizeof(int) * n;
    e = tour_from_device_order(c);
    if (e) return e;
    CK(cudaStreamSynchronize(c->stream));      // h_buf is reused
    c->h_cache.assign(path, path + n);
    c->tour_loaded = true;
    return TSPCUDA_OK;
}

static int read_state(tspcuda_ctx *c) {
    CK(cudaMemcpyAsync(c->h_state, c->d_state, sizeof(DescState), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->stats.d2h_bytes += sizeof(DescState);
    return TSPCUDA_OK;
}

template <class M>
static cudaError_t launch_descend(tspcuda_ctx *c, int grid, long long max_moves, int init_best, int dns_is_dn) {
    TourView V = c->tour;
    if (dns_is_dn) V.dns = V.dn;
    unsigned long long *key = c->d_key;
    DescState *st = c->d_state;
    MoveRec *log = c->d_log;
    long long log_cap = c->log_cap;
    void *args[] = {&V, &key, &st, &log, &log_cap, &max_moves, &init_best, &dns_is_dn};
    if (c->block == BLOCK_HUGE)
        return cudaLaunchCooperativeKernel((void *)descend_kernel<M, BLOCK_HUGE>, dim3(grid), dim3(BLOCK_HUGE), args, 0, c->stream);
    if (c->block == BLOCK_BIG)
        return cudaLaunchCooperativeKernel((void *)descend_kernel<M, BLOCK_BIG>, dim3(grid), dim3(BLOCK_BIG), args, 0, c->stream);
    return cudaLaunchCooperativeKernel((void *)descend_kernel<M, BLOCK_SMALL>, dim3(grid), dim3(BLOCK_SMALL), args, 0, c->stream);
}

template <class M>
static cudaError_t launch_descend_pruned(tspcuda_ctx *c, int grid, long long max_moves) {
    TourView V = c->tour;
    V.dns = V.dn;
    V.o8 = c->d_o8;
    SpatialGrid G = c->grid;
    unsigned long long *key = c->d_key, *qctr = c->d_qctr;
    DescState *st = c->d_state;
    MoveRec *log = c->d_log;
    long long log_cap = c->log_cap;
    BigBox *boxes = c->d_boxes;
    uint4 *wbuf = c->d_wbuf;
    NbCache nb;
    nb.lines = c->d_nb; nb.heap = c->d_nb_heap; nb.heap_next = c->d_nb_next; nb.cnt8 = c->d_cnt8; nb.heap_cap = nb_heap_cap(c->n, c->nb_heap_per_node); nb.giant_age = c->nb_giant_age; nb.restart_pct = c->nb_restart_pct;
    if (!c->nb_valid) {                                        // the tour was replaced or moved by another kernel: every list is stale
        nb_invalidate_kernel<<<grid_for(c->n, 4 * c->num_sms), 256, 0, c->stream>>>(nb, c->d_o8, c->n);
        c->stats.kernel_launches += 1;
        c->nb_valid = true;
        c->nb_moves = 0;
    }
    cudaError_t me = cudaMemsetAsync(c->d_qctr, 0, sizeof(unsigned long long), c->stream);      // the giant-box queue (shared with the tabu kernel)
    if (me != cudaSuccess) return me;
    long long moves0 = c->nb_moves;
    void *args[] = {&V, &G, &key, &st, &log, &log_cap, &max_moves, &qctr, &boxes, &wbuf, &nb, &moves0};
    if (c->pruned_block == 1024)
        return cudaLaunchCooperativeKernel((void *)descend_pruned_kernel<M, 1024>, dim3(grid), dim3(1024), args, 0, c->stream);
    return cudaLaunchCooperativeKernel((void *)descend_pruned_kernel<M, PRUNED_BLOCK>, dim3(grid), dim3(PRUNED_BLOCK), args, 0, c->stream);
}

// Run up to max_moves scan+move iterations on the device-resident tour (max_moves < 0: until no
// admissible move is left).  mode 0 = 2-opt (improving moves only), 1 = tabu (best admissible).
// The loop lives on the device; the host only re-launches in chunks of ~0.1 s so that the
// reference's time limit (refinment.c:17-24) can be honoured between chunks.
static int descend(tspcuda_ctx *c, long long max_moves, int mode, double time_limit_s, bool *deadline) {
    const int n = c->n;
    if (deadline) *deadline = false;
    if (n < 4) {   // refinment.c:55 skips every pair of a tour with fewer than 4 nodes
        const int one = 1;
        CK(cudaMemcpyAsync(&c->d_state->status, &one, sizeof(int), cudaMemcpyHostToDevice, c->stream));
        return read_state(c);
    }
    const int nstrips = (n - 2 + SR - 1) / SR;
    const long long total_cols = (long long)nstrips * (n - 2) - (long long)SR * ((long long)nstrips * (nstrips - 1) / 2);
    const int grid = (int)std::max<long long>(1, std::min<long long>(c->max_coop_blocks, total_cols / ((c->block / 32) * 48)));
    const auto t0 = std::chrono::steady_clock::now();
    const long long base = c->dev_moves;
    const int init_best = mode == 0 ? -1 : BIGPEN / 2;
    const int zero = 0;
    CK(cudaMemcpyAsync(&c->d_state->status, &zero, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    // 2-opt descents may use the exact pruned scan (K7).  It returns the same move as the full scan,
    // so the choice is only about speed: AUTO probes it with a short chunk and keeps whichever
    // kernel has the lower measured time per scan, re-probing every 8th chunk.
    const bool can_prune = mode == 0 && c->pruned_blocks > 0 && c->scan_mode != TSPCUDA_SCAN_FULL;
    const int pgrid = c->pruned_blocks;          // every SM: phase 2 splits the queued boxes over all warps
    for (;;) {
        const long long done = c->dev_moves - base;
        if (max_moves >= 0 && done >= max_moves) break;
        bool pruned = false, probe = false;
        const long long remaining = max_moves >= 0 ? max_moves - done : (1LL << 60);
        if (can_prune) {
            if (c->scan_mode == TSPCUDA_SCAN_PRUNED) pruned = true;
            else if (c->est_pruned_s <= 0) pruned = probe = true;                       // never measured
            else if (c->est_pruned_s <= c->est_full_s) pruned = true;
            else if (remaining >= 64 ? c->probe_due : (c->chunk_no & 7) == 7) pruned = probe = true;
        }
        c->chunk_no++;
        const double est = pruned ? std::max(c->est_pruned_s, 2e-6) : c->est_scan_s;
        long long chunk = (long long)std::max(4.0, std::min(8192.0, 0.1 / est));
        if (probe) chunk = 8;                                                          // (the first pruned scan after full scans rebuilds the delta cache)
        else if (can_prune && !pruned) chunk = std::min<long long>(chunk, 96);         // full scans in AUTO: re-probe soon
        c->probe_due = can_prune && !pruned;                                         // after a full chunk, probe again
        if (max_moves >= 0) chunk = std::min(chunk, max_moves - done);
        CK(cudaMemsetAsync(c->d_key, 0xFF, sizeof(unsigned long long) * 2, c->stream));
        CK(cudaEventRecord(c->ev0, c->stream));
        cudaError_t le = cudaSuccess;
        if (pruned) { DISPATCH_METRIC(c->metric, le = launch_descend_pruned<M>(c, pgrid, chunk)); }
        else { DISPATCH_METRIC(c->metric, le = launch_descend<M>(c, grid, chunk, init_best, mode == 0 ? 1 : 0)); }
        CK(le);
        CK(cudaEventRecord(c->ev1, c->stream));
        c->stats.kernel_launches += 1;
        int e = read_state(c);
        if (e) return e;
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        const long long applied = c->h_state->moves - c->dev_moves;
        c->dev_moves = c->h_state->moves;
        if (!pruned && applied > 0) c->nb_valid = false;       // moves the delta cache has not seen
        if (pruned) c->nb_moves += applied;
        const long long scans = applied + (c->h_state->status == 1 ? 1 : 0);
        c->stats.moves += applied;
        c->stats.scans += scans;
        if (pruned) {
            c->stats.pruned_scans += scans;
            c->stats.candidate_evals += c->h_state->candidates - c->dev_candidates;
            c->stats.pruned_kernel_ms += ms;
            if (scans > 0) c->est_pruned_s = std::max(1e-6, 1e-3 * ms / (double)scans);
        } else {
            c->stats.pair_evals += scans * ((long long)n * (n - 1) / 2);
            c->stats.scan_kernel_ms += ms;
            if (scans > 0) {
                c->est_scan_s = std::max(1e-6, 1e-3 * ms / (double)scans);
                if (mode == 0) c->est_full_s = c->est_scan_s;
            }
        }
        c->dev_candidates = c->h_state->candidates;
        c->stats.last_kernel_ms = ms;
#ifdef TSP_PHASE_TIMING
        if (pruned && c->h_state->status == 1)
            fprintf(stderr, "[tspcuda] pruned phases: scan1 %.1f ms (slowest warp ever %.1f us, mean warp %.2f us)  scan2 %.1f ms  apply %.1f ms | nodes reversed per move %.0f, CTA 0 own part of the move phase %.0f ns\n",
                    c->h_state->phase_ns[0] * 1e-6, c->h_state->phase_ns[3] * 1e-3,
                    c->h_state->phase_sum * 1e-3 / ((double)(c->h_state->moves + 1) * c->pruned_blocks * (c->pruned_block / 32)),
                    c->h_state->phase_ns[1] * 1e-6, c->h_state->phase_ns[2] * 1e-6,
                    (double)c->h_state->phase_ns[4] / std::max<long long>(1, c->h_state->moves + 1),
                    (double)c->h_state->phase_ns[5] / std::max<long long>(1, c->h_state->moves + 1));
#endif
#ifdef TSP_NB_STATS
        if (pruned && c->h_state->status == 1) {
            unsigned long long h[8];
            cudaMemcpyFromSymbol(h, g_nbstat, sizeof(h));
            const double sc = (double)std::max<long long>(1, c->h_state->moves + 1);
            fprintf(stderr, "[tspcuda] delta cache per scan: rebuilds %.1f (box nodes %.0f)  giants/no-room %.2f  heap slots %.0f  dirty %.1f  heap live %.0f  evaluated %.1f  attend calls %.1f\n",
                    h[0] / sc, h[7] / sc, h[1] / sc, h[2] / sc, h[3] / sc, h[4] / sc, h[5] / sc, h[6] / sc);
            unsigned long long z[8] = {0};
            cudaMemcpyToSymbol(g_nbstat, z, sizeof(z));
        }
#endif
        if (c->h_state->status == 1) break;
        if (time_limit_s >= 0) {
            const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (el > time_limit_s) { if (deadline) *deadline = true; break; }
        }
    }
    c->tour_loaded = (c->dev_moves == base) ? c->tour_loaded : false;   // host cache no longer matches
    return TSPCUDA_OK;
}

extern "C" int tspcuda_tour_descend(tspcuda_ctx *c, int64_t max_moves, int64_t *moves_out, int *at_local_optimum) {
    int e = use(c);
    if (e) return e;
    const long long base = c->dev_moves;
    e = descend(c, max_moves, 0, -1.0, nullptr);
    if (e) return e;
    if (moves_out) *moves_out = c->dev_moves - base;
    if (at_local_optimum) *at_local_optimum = c->h_state->status == 1;
    return TSPCUDA_OK;
}

extern "C" int tspcuda_tour_download(tspcuda_ctx *c, int *path_out, double *cost_out) {
    int e = use(c);
    if (e) return e;
    const int n = c->n;
    if (path_out) {
        tour_export_kernel<<<grid_for(n, 4 * c->num_sms), 256, 0, c->stream>>>(c->tour, c->d_path);
        c->stats.kernel_launches += 1;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(c->h_buf, c->d_path, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
        c->stats.d2h_bytes += (int64_t)sizeof(int) * n;
    }
    e = read_state(c);
    if (e) return e;
    if (path_out) {
        std::memcpy(path_out, c->h_buf, sizeof(int) * n);
        c->h_cache.assign(path_out, path_out + n);
        c->tour_loaded = true;
    }
    if (cost_out) *cost_out = (double)c->h_state->cost;
    return TSPCUDA_OK;
}

extern "C" int tspcuda_tour_movelog(tspcuda_ctx *c, tspcuda_move *log, int64_t cap, int64_t *count_out) {
    int e = use(c);
    if (e) return e;
    e = read_state(c);
    if (e) return e;
    const long long have = std::min<long long>(c->h_state->log_count, c->log_cap);
    const long long cnt = std::min<long long>(have, cap);
    if (log && cnt > 0) {
        static_assert(sizeof(tspcuda_move) == sizeof(MoveRec), "move record layout");
        CK(cudaMemcpyAsync(log, c->d_log, sizeof(MoveRec) * cnt, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        c->stats.d2h_bytes += (int64_t)sizeof(MoveRec) * cnt;
    }
    if (count_out) *count_out = c->h_state->log_count;
    return TSPCUDA_OK;
}

// ---- 2-opt entry points ---------------------------------------------------------------------
extern "C" int tspcuda_two_opt(tspcuda_ctx *c, int *path, double *cost_out, int64_t max_moves, double time_limit_s,
                               int64_t *moves_out, tspcuda_move *log, int64_t log_cap) {
    int e = use(c);
    if (e) return e;
    c->tour_loaded = false;                       // ref_2opt recomputes the cost from the path it is given
    e = tspcuda_tour_upload(c, path);
    if (e) return e;
    bool deadline = false;
    e = descend(c, max_moves, 0, time_limit_s, &deadline);
    if (e) return e;
    if (moves_out) *moves_out = c->dev_moves;
    if (log && log_cap > 0) {
        int64_t cnt = 0;
        e = tspcuda_tour_movelog(c, log, log_cap, &cnt);
        if (e) return e;
    }
    e = tspcuda_tour_download(c, path, cost_out);
    if (e) return e;
    return deadline ? TSPCUDA_DEADLINE_EXCEEDED : TSPCUDA_OK;
}

extern "C" int tspcuda_two_opt_once(tspcuda_ctx *c, int *path, double *cost_inout, double *best_delta_out,
                                    tspcuda_move *move_out) {
    int e = use(c);
    if (e) return e;
    if (!cost_inout) return fail(TSPCUDA_INVALID_ARGUMENT, "null cost");
    e = tspcuda_tour_upload(c, path);
    if (e) return e;
    const long long base = c->dev_moves;
    e = descend(c, 1, 0, -1.0, nullptr);
    if (e) return e;
    const bool moved = c->dev_moves > base;
    if (move_out) {
        move_out->a = moved ? c->h_state->last_a : -1;
        move_out->b = moved ? c->h_state->last_b : -1;
        move_out->delta = moved ? c->h_state->last_delta : 0;
        move_out->reserved = 0;
    }
    if (best_delta_out) *best_delta_out = moved ? (double)c->h_state->last_delta : 0.0;
    if (moved) {
        *cost_inout += (double)c->h_state->last_delta;                  // refinment.c:85
        e = tspcuda_tour_download(c, path, nullptr);
        if (e) return e;
    }
    return TSPCUDA_OK;
}

extern "C" int tspcuda_tabu_move(tspcuda_ctx *c, int *path, double *cost_inout, int *tabu_list, int tenure, int iter,
                                 tspcuda_move *move_out) {
    int e = use(c);
    if (e) return e;
    if (!cost_inout || !tabu_list) return fail(TSPCUDA_INVALID_ARGUMENT, "null argument");
    const int n = c->n;
    e = tspcuda_tour_upload(c, path);
    if (e) return e;
    if (move_out) { move_out->a = -1; move_out->b = -1; move_out->delta = 0; move_out->reserved = 0; }
    if (n < 4) return TSPCUDA_OK;
    CK(cudaMemcpyAsync(c->d_tabu, tabu_list, sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
    c->stats.h2d_bytes += (int64_t)sizeof(int) * n;
    tabu_flags_kernel<<<grid_for(n + 1, 4 * c->num_sms), 256, 0, c->stream>>>(c->tour, c->d_tabu, tenure, iter);
    c->stats.kernel_launches += 1;
    CK(cudaGetLastError());
    const long long base = c->dev_moves;
    e = descend(c, 1, 1, -1.0, nullptr);
    if (e) return e;
    if (c->dev_moves > base) {
        const int a = c->h_state->last_a, b = c->h_state->last_b;
        const int sa = path[a], sb = path[b];                            // successors BEFORE the move
        *cost_inout += (double)c->h_state->last_delta;                  // metaheuristic.c:232
        tabu_list[a] = iter; tabu_list[b] = iter; tabu_list[sa] = iter; tabu_list[sb] = iter;   // :235-238
        if (move_out) { move_out->a = a; move_out->b = b; move_out->delta = c->h_state->last_delta; }
        e = tspcuda_tour_download(c, path, nullptr);
        if (e) return e;
    }
    return TSPCUDA_OK;
}

struct DevBuf {
    void *p = nullptr;
    bool owned = true;
    ~DevBuf() { if (owned) cudaFree(p); }
    template <class T> T *as() { return reinterpret_cast<T *>(p); }
};

static cudaError_t scratch_get(tspcuda_ctx *c, int idx, size_t bytes, void **out) {
    ScratchBuf &b = c->scratch[idx];
    if (b.bytes < bytes) {
        cudaFree(b.p);
        b.p = nullptr; b.bytes = 0;
        cudaError_t e = cudaMalloc(&b.p, bytes);
        if (e != cudaSuccess) return e;
        b.bytes = bytes;
    }
    *out = b.p;
    return cudaSuccess;
}
static cudaError_t scratch_get(tspcuda_ctx *c, int idx, size_t bytes, DevBuf &buf) {
    buf.owned = false;
    return scratch_get(c, idx, bytes, &buf.p);
}

__global__ void order_to_path_kernel(const int *order, int n, int *path) {
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x)
        path[order[p]] = order[p + 1 == n ? 0 : p + 1];
}

// ---- device-resident tabu search (metaheuristic.c:128-164 around :188-245) ---------------------
template <class M>
static cudaError_t launch_tabu(tspcuda_ctx *c, int grid, TabuCtl T, TabuOut *out, long long iters) {
    TourView V = c->tour;
    unsigned long long *key = c->d_key;
    void *args[] = {&V, &key, &T, &out, &iters};
    if (c->block == BLOCK_HUGE)
        return cudaLaunchCooperativeKernel((void *)tabu_kernel<M, BLOCK_HUGE>, dim3(grid), dim3(BLOCK_HUGE), args, 0, c->stream);
    if (c->block == BLOCK_BIG)
        return cudaLaunchCooperativeKernel((void *)tabu_kernel<M, BLOCK_BIG>, dim3(grid), dim3(BLOCK_BIG), args, 0, c->stream);
    return cudaLaunchCooperativeKernel((void *)tabu_kernel<M, BLOCK_SMALL>, dim3(grid), dim3(BLOCK_SMALL), args, 0, c->stream);
}

extern "C" int tspcuda_tabu_search(tspcuda_ctx *c, int *path, double *cost_inout, int *tabu_list, const int *tenures,
                                   int first_iter, int n_iters, double incumbent_cost, double *cost_trace_out,
                                   int *best_path_out, double *best_cost_out, int *best_iter_out) {
    int e = use(c);
    if (e) return e;
    if (!path || !cost_inout || !tabu_list || (n_iters > 0 && !tenures)) return fail(TSPCUDA_INVALID_ARGUMENT, "null argument");
    if (n_iters < 0) return fail(TSPCUDA_INVALID_ARGUMENT, "n_iters = %d", n_iters);
    const int n = c->n;
    if (best_iter_out) *best_iter_out = -1;
    if (best_cost_out) *best_cost_out = incumbent_cost;
    e = tspcuda_tour_upload(c, path);
    if (e) return e;
    if (n < 4 || n_iters == 0) {                       // no admissible pair exists (refinment.c:55 rule)
        for (int k = 0; k < n_iters && cost_trace_out; k++) cost_trace_out[k] = *cost_inout;
        return TSPCUDA_OK;
    }
    DevBuf d_ten, d_trace, d_best, d_out;
    CK(scratch_get(c, SC_T_TEN, sizeof(int) * n_iters, d_ten));
    CK(scratch_get(c, SC_T_TRACE, sizeof(long long) * n_iters, d_trace));
    CK(scratch_get(c, SC_T_BEST, sizeof(int) * ((size_t)n + 1), d_best));
    CK(scratch_get(c, SC_T_OUT, sizeof(TabuOut), d_out));
    CK(cudaMemcpyAsync(d_ten.p, tenures, sizeof(int) * n_iters, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_tabu, tabu_list, sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
    c->stats.h2d_bytes += (int64_t)sizeof(int) * ((int64_t)n + n_iters);
    const int nstrips = (n - 2 + SR - 1) / SR;
    const long long total_cols = (long long)nstrips * (n - 2) - (long long)SR * ((long long)nstrips * (nstrips - 1) / 2);
    const int grid = (int)std::max<long long>(1, std::min<long long>(c->max_coop_blocks, total_cols / ((c->block / 32) * 48)));
    TabuCtl T;
    T.tabu_list = c->d_tabu; T.best_tid = d_best.as<int>();
    // the reference's "no incumbent yet" is DBL_MAX (tsp.c:37), API callers may pass +inf: both mean "anything is better"
    if (std::isnan(incumbent_cost) || std::isnan(*cost_inout) || !std::isfinite(*cost_inout) || std::fabs(*cost_inout) >= 9.0e18)
        return fail(TSPCUDA_INVALID_ARGUMENT, "tour cost / incumbent cost is not a usable number");
    T.cost = (long long)*cost_inout;
    T.best_cost = incumbent_cost >= 9.0e18 ? LLONG_MAX : (incumbent_cost <= -9.0e18 ? LLONG_MIN : (long long)incumbent_cost);
    T.pad = 0;
    c->nb_valid = false;                               // the tabu kernels move the tour without maintaining the delta cache
    TabuOut h_out{};
    h_out.best_iter = -1;
    long long moves = 0;
    const int CHUNK = 4096;                            // iterations per launch
    // K5c (pruned scan) while it keeps finding an admissible pair below its bound, K5b (full scan) for FALLBACK
    // iterations whenever it gives up; both continue the same device-resident state.
    const bool can_prune = c->pruned_blocks > 0 && c->scan_mode != TSPCUDA_SCAN_FULL && n >= 4096;
    const int FALLBACK = 32;
    int cooldown = 0, bound = 64;
    for (int done = 0; done < n_iters;) {
        const bool pruned = can_prune && cooldown == 0;
        const int cnt = pruned ? std::min(CHUNK, n_iters - done) : std::min(std::min(CHUNK, n_iters - done), cooldown > 0 ? cooldown : CHUNK);
        T.tenure = d_ten.as<int>() + done;
        T.trace = d_trace.as<long long>() + done;
        T.first_iter = first_iter + done;
        CK(cudaMemsetAsync(c->d_key, 0xFF, sizeof(unsigned long long) * 2, c->stream));
        CK(cudaEventRecord(c->ev0, c->stream));
        cudaError_t le = cudaSuccess;
        if (pruned) {
            CK(cudaMemsetAsync(c->d_qctr, 0, sizeof(unsigned long long), c->stream));
            TourView V = c->tour;
            SpatialGrid G = c->grid;
            unsigned long long *key = c->d_key, *qctr = c->d_qctr;
            TabuOut *o = d_out.as<TabuOut>();
            long long iters = cnt;
            BigBox *boxes = c->d_boxes;
            uint4 *wbuf = c->d_wbuf;
            void *args[] = {&V, &G, &key, &T, &o, &iters, &qctr, &boxes, &wbuf, &bound};
            const bool four = 4LL * n <= (long long)c->pruned_blocks * PRUNED_BLOCK;
            if (four) { DISPATCH_METRIC(c->metric, le = cudaLaunchCooperativeKernel((void *)tabu_pruned_kernel<M, PRUNED_BLOCK, 4>, dim3(c->pruned_blocks), dim3(PRUNED_BLOCK), args, 0, c->stream)); }
            else { DISPATCH_METRIC(c->metric, le = cudaLaunchCooperativeKernel((void *)tabu_pruned_kernel<M, PRUNED_BLOCK, 1>, dim3(c->pruned_blocks), dim3(PRUNED_BLOCK), args, 0, c->stream)); }
        } else {
            DISPATCH_METRIC(c->metric, le = launch_tabu<M>(c, grid, T, d_out.as<TabuOut>(), cnt));
        }
        CK(le);
        CK(cudaEventRecord(c->ev1, c->stream));
        c->stats.kernel_launches += 1;
        TabuOut o{};
        CK(cudaMemcpyAsync(&o, d_out.p, sizeof(TabuOut), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        c->stats.last_kernel_ms = ms;
        c->stats.scans += o.done;
        c->stats.moves += o.moves;
        if (pruned) {
            c->stats.pruned_kernel_ms += ms;
            c->stats.pruned_scans += o.done;
            cooldown = o.done < cnt ? FALLBACK : 0;
            if (o.last_delta > 0) bound = std::min(2 * o.last_delta + 64, 1 << 26);
        } else {
            c->stats.scan_kernel_ms += ms;
            c->stats.pair_evals += o.done * ((long long)n * (n - 1) / 2);
            if (cooldown > 0) cooldown = std::max(0, cooldown - (int)o.done);
        }
        moves += o.moves;
        T.cost = o.cost; T.best_cost = o.best_cost;
        h_out.cost = o.cost; h_out.best_cost = o.best_cost;
        if (o.best_iter >= 0) h_out.best_iter = o.best_iter;
        done += (int)o.done;
    }
    c->tour_loaded = false;                            // (the tabu kernels do not touch d_state->moves: dev_moves stays its mirror)
    (void)moves;
    // results: trace, tabu list, final tour, best tour
    if (cost_trace_out) {
        std::vector<long long> tr(n_iters);
        CK(cudaMemcpyAsync(tr.data(), d_trace.p, sizeof(long long) * n_iters, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        for (int k = 0; k < n_iters; k++) cost_trace_out[k] = (double)tr[k];
        c->stats.d2h_bytes += (int64_t)sizeof(long long) * n_iters;
    }
    CK(cudaMemcpyAsync(tabu_list, c->d_tabu, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
    c->stats.d2h_bytes += (int64_t)sizeof(int) * n;
    {   // d_state->cost mirrors the device tour for later tour_download calls
        const long long cst = h_out.cost;
        CK(cudaMemcpyAsync(&c->d_state->cost, &cst, sizeof(long long), cudaMemcpyHostToDevice, c->stream));
    }
    e = tspcuda_tour_download(c, path, nullptr);
    if (e) return e;
    *cost_inout = (double)h_out.cost;
    if (h_out.best_iter >= 0) {
        if (best_path_out) {
            order_to_path_kernel<<<grid_for(n, 4 * c->num_sms), 256, 0, c->stream>>>(d_best.as<int>(), n, c->d_path);
            c->stats.kernel_launches += 1;
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(best_path_out, c->d_path, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
            CK(cudaStreamSynchronize(c->stream));
            c->stats.d2h_bytes += (int64_t)sizeof(int) * n;
        }
        if (best_cost_out) *best_cost_out = (double)h_out.best_cost;
        if (best_iter_out) *best_iter_out = h_out.best_iter;
    }
    return TSPCUDA_OK;
}

// ---- nearest neighbour ----------------------------------------------------------------------
struct NNScratch;
static bool nn_use_grid(const tspcuda_ctx *c) {
    if (c->metric == M_GEO || c->pruned_blocks <= 0) return false;      // the grid is Euclidean
    if (const char *f = getenv("TSPCUDA_NN")) {
        if (!strcmp(f, "grid")) return true;
        if (!strcmp(f, "list")) return false;
    }
    return c->n >= 12000;          // below, a step of the plain kernel is a handful of iterations per thread
}

struct NNScratch {
    double2 *cxy = nullptr;
    int *cid = nullptr, *order = nullptr, *slot = nullptr;
    long long *cost = nullptr;
};

static int nn_batch(tspcuda_ctx *c, NNScratch &S, int first_start, int count, int stride, long long *h_cost) {
    const int n = c->n;
    if (S.slot) {                                     // K4b: grid-accelerated (nn_alloc decided)
        const size_t smem = sizeof(unsigned) * (((size_t)n + 31) / 32);
        DISPATCH_METRIC(c->metric, (nn_grid_kernel<M><<<count, NNG_BLOCK, smem, c->stream>>>(c->d_xy, n, c->grid, first_start, stride, S.cxy, S.cid, S.slot, S.order, S.cost)));
    } else {
        DISPATCH_METRIC(c->metric, (nn_kernel<M><<<count, NN_BLOCK, 0, c->stream>>>(c->d_xy, n, first_start, stride, S.cxy, S.cid, S.order, S.cost)));
    }
    c->stats.kernel_launches += 1;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h_cost, S.cost, sizeof(long long) * count, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->stats.d2h_bytes += (int64_t)sizeof(long long) * count;
    return TSPCUDA_OK;
}

static int nn_alloc(tspcuda_ctx *c, NNScratch &S, int batch) {
    const size_t n = (size_t)c->n;
    CK(scratch_get(c, SC_NN_CXY, sizeof(double2) * n * batch, (void **)&S.cxy));          // the context's grow-only scratch
    CK(scratch_get(c, SC_NN_CID, sizeof(int) * n * batch, (void **)&S.cid));
    CK(scratch_get(c, SC_NN_ORDER, sizeof(int) * n * batch, (void **)&S.order));
    CK(scratch_get(c, SC_NN_COST, sizeof(long long) * batch, (void **)&S.cost));
    if (nn_use_grid(c)) CK(scratch_get(c, SC_NN_SLOT, sizeof(int) * n * batch, (void **)&S.slot));
    return TSPCUDA_OK;
}

// ---- time limits of the batched entry points ---------------------------------------------------------------
// The reference looks at its clock before every start (heuristics.c:42-50, :83-90) and before every 2-opt scan
// (refinment.c:17-24).  The batched kernels look at a device word instead, which the host raises once the limit has passed.
struct Deadline {
    bool on = false;
    std::chrono::steady_clock::time_point end;
    explicit Deadline(double limit_s) {
        if (limit_s >= 0) { on = true; end = std::chrono::steady_clock::now() + std::chrono::duration_cast<std::chrono::steady_clock::duration>(std::chrono::duration<double>(limit_s)); }
    }
    bool passed() const { return on && std::chrono::steady_clock::now() >= end; }
    double left() const { return on ? std::max(0.0, std::chrono::duration<double>(end - std::chrono::steady_clock::now()).count()) : -1.0; }
};

// Wait for everything queued on the context's stream; when the deadline passes meanwhile, raise the stop word by a DMA
// copy on the auxiliary stream (a copy engine needs no SM; a memset kernel queued behind persistent CTAs might never run).
static int wait_with_deadline(tspcuda_ctx *c, const Deadline &dl, bool *raised) {
    if (dl.on) {
        for (;;) {
            const cudaError_t q = cudaStreamQuery(c->stream);
            if (q == cudaSuccess) break;
            if (q != cudaErrorNotReady) CK(q);
            if (!*raised && dl.passed()) {
                CK(cudaMemcpyAsync(c->d_stop, c->h_one, sizeof(int), cudaMemcpyHostToDevice, c->aux_stream));
                *raised = true;
            }
            usleep(100);
        }
    }
    CK(cudaStreamSynchronize(c->stream));
    return TSPCUDA_OK;
}

extern "C" int tspcuda_nn_multi(tspcuda_ctx *c, int first_start, int nstarts, int stride, double time_limit_s, double *costs_out,
                                int *best_start_out, int *best_path_out, double *best_cost_out, int *starts_done_out) {
    int e = use(c);
    if (e) return e;
    const int n = c->n;
    if (starts_done_out) *starts_done_out = 0;
    if (nstarts <= 0 || stride <= 0) return fail(TSPCUDA_INVALID_ARGUMENT, "nstarts/stride");
    const long long last = (long long)first_start + (long long)(nstarts - 1) * stride;
    if (first_start < 0 || last >= n) return fail(TSPCUDA_UNAVAILABLE, "starting node not correct");   // heuristics.c:223-226
    const Deadline dl(time_limit_s);
    // with a time limit the clock is looked at between batches (the reference: before every start, heuristics.c:42-50)
    const int batch = std::min(nstarts, (dl.on ? 2 : 8) * c->num_sms);
    NNScratch S;
    e = nn_alloc(c, S, batch);
    if (e) return e;
    std::vector<long long> hc(batch);
    long long best = -1;
    int best_start = -1, done = 0;
    for (; done < nstarts; done += batch) {
        if (dl.passed()) break;
        const int cnt = std::min(batch, nstarts - done);
        const int fs = first_start + done * stride;
        e = nn_batch(c, S, fs, cnt, stride, hc.data());
        if (e) return e;
        int bi = -1;
        for (int k = 0; k < cnt; k++) {
            if (costs_out) costs_out[done + k] = (double)hc[k];
            if (best_start < 0 || hc[k] < best) { best = hc[k]; best_start = fs + k * stride; bi = k; }   // strict '<', heuristics.c:58
        }
        if (bi >= 0 && best_path_out) {
            order_to_path_kernel<<<grid_for(n, 4 * c->num_sms), 256, 0, c->stream>>>(S.order + (size_t)bi * n, n, c->d_path);
            c->stats.kernel_launches += 1;
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(best_path_out, c->d_path, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
            CK(cudaStreamSynchronize(c->stream));
            c->stats.d2h_bytes += (int64_t)sizeof(int) * n;
        }
    }
    done = std::min(done, nstarts);
    if (starts_done_out) *starts_done_out = done;
    if (best_start_out) *best_start_out = best_start;
    if (best_cost_out) *best_cost_out = (double)best;
    return done < nstarts ? TSPCUDA_DEADLINE_EXCEEDED : TSPCUDA_OK;
}

extern "C" int tspcuda_nn(tspcuda_ctx *c, int start, int *path_out, double *cost_out) {
    if (!c) return fail(TSPCUDA_INVALID_ARGUMENT, "null context");
    if (start < 0 || start >= c->n) return fail(TSPCUDA_UNAVAILABLE, "starting node not correct");     // heuristics.c:223-226
    int bs = -1;
    return tspcuda_nn_multi(c, start, 1, 1, -1.0, nullptr, &bs, path_out, cost_out, nullptr);
}

// ---- multi-start NN + 2-opt (heuristics.c:74-116) --------------------------------------------
template <class M, int BLK>
static cudaError_t launch_cta_descents(tspcuda_ctx *c, const int *d_order, int count, int *d_next, long long *d_cost,
                                       long long *d_moves, int *d_paths, int *d_status, const int *d_stop) {
    const size_t smem = cta_smem_bytes(c->n, BLK / 32);
    cudaError_t e = cudaFuncSetAttribute(descend_cta_kernel<M, BLK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, descend_cta_kernel<M, BLK>, BLK, smem);
    if (e != cudaSuccess) return e;
    const int grid = std::max(1, std::min(count, std::max(1, per_sm) * c->num_sms));
    descend_cta_kernel<M, BLK><<<grid, BLK, smem, c->stream>>>(c->d_xy, c->n, d_order, count, d_next, d_cost, d_moves, d_paths, d_status, d_stop);
    return cudaGetLastError();
}

// TSPCUDA_VERBOSE: where a multi-start call spends its wall clock (allocation / NN / descents / release); declared first
// in the function so that its destructor runs after the buffers' own
struct PhaseClock {
    bool on = getenv("TSPCUDA_VERBOSE") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now(), tl = t0;
    double alloc = 0, nn = 0, descents = 0, rest = 0;
    void lap(double &acc) {
        if (!on) return;
        const auto t = std::chrono::steady_clock::now();
        acc += std::chrono::duration<double>(t - tl).count();
        tl = t;
    }
    ~PhaseClock() {
        if (!on) return;
        const double total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        fprintf(stderr, "tspcuda: multistart %.3f s = alloc %.3f + NN %.3f + descents %.3f + copies %.3f + release %.3f\n", total, alloc, nn,
                descents, rest, total - alloc - nn - descents - rest);
    }
};

// K8 with clusters (a descent shared by 2/4/8 SMs): from this n up a scan is long enough for the cluster barriers not to
// matter, and from this many starts up a batch beats descending them one after the other on the whole GPU (K7)
constexpr int K8_CLUSTER_MIN_N = 8192;
constexpr int K8_CLUSTER_MIN_STARTS = 12;

extern "C" int tspcuda_multistart_two_opt(tspcuda_ctx *c, int first_start, int nstarts, int stride, double time_limit_s,
                                          double *nn_costs_out, double *final_costs_out, int64_t *moves_out, int *best_start_out,
                                          int *best_path_out, double *best_cost_out, int *starts_done_out) {
    PhaseClock clk;
    int e = use(c);
    if (e) return e;
    const int n = c->n;
    if (starts_done_out) *starts_done_out = 0;
    if (nstarts <= 0 || stride <= 0) return fail(TSPCUDA_INVALID_ARGUMENT, "nstarts/stride");
    const long long last = (long long)first_start + (long long)(nstarts - 1) * stride;
    if (first_start < 0 || last >= n) return fail(TSPCUDA_UNAVAILABLE, "starting node not correct");
    const Deadline dl(time_limit_s);
    // K8 (one CTA per descent, pruned scan, any n) when there are enough starts to occupy the SMs;
    // K6 (one CTA per descent, full scan, tour in shared memory) when every pair must be evaluated;
    // otherwise the starts descend one after the other on the whole GPU (K2 / K7).
    const char *k8env = getenv("TSPCUDA_K8_CLUSTER");       // measurement knob: 0 = never batch (K7 one start after the other), 1/2/4/8 = cluster size
    const bool multi_path = n >= 4 && c->pruned_blocks > 0 && c->scan_mode != TSPCUDA_SCAN_FULL && !(k8env && atoi(k8env) == 0) &&
                            nstarts >= (k8env ? 1 : n >= K8_CLUSTER_MIN_N ? K8_CLUSTER_MIN_STARTS : c->num_sms / 4);
    const bool cta_path = !multi_path && n >= 4 && n <= CTA_MAX_N;       // whole tour fits one CTA's shared memory
    const int batch = std::min(nstarts, ((cta_path || multi_path) ? 16 : 8) * c->num_sms);
    NNScratch S;
    e = nn_alloc(c, S, batch);
    if (e) return e;
    DevBuf best_path, next, dcost, dmoves, dpaths, dstatus;
    CK(scratch_get(c, SC_BEST_PATH, sizeof(int) * n, best_path));
    std::vector<long long> hc(batch), hf(batch), hm(batch);
    std::vector<int> hs(batch);
    DevBuf m_pxy, m_tid, m_dn, m_pos, m_boxes, m_bpath, m_bcost, m_bstart, m_cand, m_nb, m_nbheap, m_nbnext, m_o8, m_cnt8;
    MultiSlots MS{};
    const int mslots = multi_path ? std::min(batch, c->num_sms) : 0;
    if (multi_path) {
        const size_t n1 = (size_t)n + 1, sl = (size_t)mslots;
        CK(scratch_get(c, SC_M_PXY, sizeof(double2) * n1 * sl, m_pxy));
        CK(scratch_get(c, SC_M_TID, sizeof(int) * n1 * sl, m_tid));
        CK(scratch_get(c, SC_M_DN, sizeof(int) * n1 * sl, m_dn));
        CK(scratch_get(c, SC_M_POS, sizeof(int) * (size_t)n * sl, m_pos));
        CK(scratch_get(c, SC_M_BOXES, sizeof(BigBox) * (size_t)n * sl, m_boxes));
        CK(scratch_get(c, SC_M_NB, sizeof(int) * (size_t)NB_WORDS * (size_t)n * sl, m_nb));
        CK(scratch_get(c, SC_M_NBHEAP, sizeof(int4) * (size_t)nb_heap_cap(n, c->nb_heap_per_node) * sl, m_nbheap));
        MS.heap_cap = nb_heap_cap(n, c->nb_heap_per_node);
        MS.restart_pct = c->nb_restart_pct;
        CK(scratch_get(c, SC_M_NBNEXT, sizeof(int) * 8 * sl, m_nbnext));
        CK(scratch_get(c, SC_M_O8, (size_t)n * sl, m_o8));
        CK(scratch_get(c, SC_M_CNT8, (size_t)nb_cnt8_stride(n) * sl, m_cnt8));
        CK(scratch_get(c, SC_M_BPATH, sizeof(int) * (size_t)n * sl, m_bpath));
        CK(scratch_get(c, SC_M_BCOST, sizeof(long long) * sl, m_bcost));
        CK(scratch_get(c, SC_M_BSTART, sizeof(int) * sl, m_bstart));
        CK(scratch_get(c, SC_M_CAND, sizeof(unsigned long long), m_cand));
        MS.pxy = m_pxy.as<double2>(); MS.tid = m_tid.as<int>(); MS.dn = m_dn.as<int>(); MS.pos = m_pos.as<int>();
        MS.boxes = m_boxes.as<BigBox>(); MS.nb = m_nb.as<int>(); MS.nb_heap = m_nbheap.as<int4>(); MS.nb_next = m_nbnext.as<int>(); MS.o8 = m_o8.as<unsigned char>(); MS.cnt8 = m_cnt8.as<unsigned char>(); MS.best_path = m_bpath.as<int>();
        MS.best_cost = m_bcost.as<long long>(); MS.best_start = m_bstart.as<int>();
    }
    std::vector<long long> hbc(std::max(1, mslots));
    std::vector<int> hbs(std::max(1, mslots));
    if (cta_path || multi_path) {
        CK(scratch_get(c, SC_NEXT, sizeof(int), next));
        CK(scratch_get(c, SC_DCOST, sizeof(long long) * batch, dcost));
        CK(scratch_get(c, SC_DMOVES, sizeof(long long) * batch, dmoves));
        CK(scratch_get(c, SC_DSTATUS, sizeof(int) * batch, dstatus));
    }
    if (cta_path) CK(scratch_get(c, SC_DPATHS, sizeof(int) * (size_t)n * batch, dpaths));
    const int *d_stop = dl.on ? c->d_stop : nullptr;
    bool raised = false, cut_short = false;
    if (dl.on) CK(cudaMemsetAsync(c->d_stop, 0, sizeof(int), c->stream));
    long long best = -1;
    int best_start = -1, done = 0;
    clk.lap(clk.alloc);
    for (; done < nstarts && !cut_short; done += batch) {
        if (dl.passed()) break;                                          // heuristics.c:83-90: no further start is taken
        const int cnt = std::min(batch, nstarts - done);
        const int fs = first_start + done * stride;
        clk.lap(clk.rest);
        e = nn_batch(c, S, fs, cnt, stride, hc.data());
        if (e) return e;
        clk.lap(clk.nn);
        if (multi_path || cta_path) {
            CK(cudaMemsetAsync(next.p, 0, sizeof(int), c->stream));
            CK(cudaMemsetAsync(dstatus.p, 0, sizeof(int) * cnt, c->stream));
            CK(cudaMemsetAsync(dmoves.p, 0, sizeof(long long) * cnt, c->stream));
            CK(cudaMemsetAsync(dcost.p, 0xFF, sizeof(long long) * cnt, c->stream));      // -1: start never taken
        }
        int g = 0;
        if (multi_path) {
            CK(cudaMemsetAsync(m_cand.p, 0, sizeof(unsigned long long), c->stream));
            CK(cudaEventRecord(c->ev0, c->stream));
            // fewer starts than SMs: a cluster of CL CTAs per descent (pruned.cuh), CL the largest of 8/4/2 that still
            // gives every start of the batch its own cluster.  A cluster occupies CL SMs of ONE GPC, so fewer than
            // num_sms / CL clusters are co-resident (GPCs of 16-20 SMs): the count the driver reports decides, not the SM
            // count (148 SMs: 37 clusters of 4 do not fit and the last one would start after the others have drained the queue)
            if (!c->k8_clusters[0]) {
                for (int k = 0, v = 1; k < 4; ++k, v *= 2) {
                    cudaLaunchConfig_t q = {};
                    q.gridDim = dim3(c->num_sms / v * v); q.blockDim = dim3(1024);
                    q.dynamicSmemBytes = k8_dyn_smem(1024, v, n);
                    cudaLaunchAttribute qa[1];
                    qa[0].id = cudaLaunchAttributeClusterDimension;
                    qa[0].val.clusterDim.x = v; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
                    q.attrs = qa; q.numAttrs = 1;
                    int nc = 0;
                    cudaError_t qe = cudaSuccess;
#define K8_QUERY(CLV) DISPATCH_METRIC(c->metric, (qe = cudaFuncSetAttribute(multi_pruned_kernel<M, 1024, CLV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)q.dynamicSmemBytes), \
                                                   qe = qe != cudaSuccess ? qe : cudaOccupancyMaxActiveClusters(&nc, multi_pruned_kernel<M, 1024, CLV>, &q)))
                    if (v == 8) { K8_QUERY(8); } else if (v == 4) { K8_QUERY(4); } else if (v == 2) { K8_QUERY(2); } else { K8_QUERY(1); }
#undef K8_QUERY
                    if (qe != cudaSuccess) { cudaGetLastError(); nc = 0; }
                    c->k8_clusters[k] = std::max(1, std::min(nc, c->num_sms / v));
                }
                if (getenv("TSPCUDA_VERBOSE"))
                    fprintf(stderr, "tspcuda: co-resident K8 clusters of 1/2/4/8 CTAs: %d %d %d %d\n", c->k8_clusters[0], c->k8_clusters[1],
                            c->k8_clusters[2], c->k8_clusters[3]);
            }
            int cl = 1, ck = 0;
            if (n >= K8_CLUSTER_MIN_N) while (cl < 8 && cnt <= c->k8_clusters[ck + 1]) { cl *= 2; ++ck; }
            if (k8env) {
                cl = std::max(1, std::min(8, atoi(k8env)));
                ck = cl >= 8 ? 3 : cl >= 4 ? 2 : cl >= 2 ? 1 : 0;
                cl = 1 << ck;
            }
            g = std::max(1, std::min(std::min(cnt, mslots), c->k8_clusters[ck]));   // teams (clusters): one tour slot each, cl record-buffer sets
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(g * cl); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = 0; cfg.stream = c->stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            cudaError_t le = cudaSuccess;
#define K8_LAUNCH(BLKV, CLV)                                                                                                  \
    cfg.dynamicSmemBytes = k8_dyn_smem(BLKV, CLV, n);                                                                            \
    DISPATCH_METRIC(c->metric, (le = cudaFuncSetAttribute(multi_pruned_kernel<M, BLKV, CLV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes))); \
    if (le == cudaSuccess) DISPATCH_METRIC(c->metric, (le = cudaLaunchKernelEx(&cfg, multi_pruned_kernel<M, BLKV, CLV>, (const double2 *)c->d_xy, n, c->grid, \
        (const int *)S.order, fs, stride, cnt, next.as<int>(), MS, dcost.as<long long>(), dmoves.as<long long>(),                     \
        dstatus.as<int>(), m_cand.as<unsigned long long>(), c->d_wbuf, d_stop)))
            // TSPCUDA_K8_BLOCK=512 (measurement knob): 16 warps of 128 registers per descent instead of 32 of 64
            const char *kb = getenv("TSPCUDA_K8_BLOCK");
            if (kb && atoi(kb) == 512) {
                cfg.blockDim = dim3(512);
                if (cl == 8) { K8_LAUNCH(512, 8); } else if (cl == 4) { K8_LAUNCH(512, 4); } else if (cl == 2) { K8_LAUNCH(512, 2); } else { K8_LAUNCH(512, 1); }
            } else {
                if (cl == 8) { K8_LAUNCH(1024, 8); } else if (cl == 4) { K8_LAUNCH(1024, 4); } else if (cl == 2) { K8_LAUNCH(1024, 2); } else { K8_LAUNCH(1024, 1); }
            }
#undef K8_LAUNCH
            CK(le);
            CK(cudaGetLastError());
            CK(cudaEventRecord(c->ev1, c->stream));
        } else if (cta_path) {
            // K6: every start of the batch descends inside one CTA, no grid barrier per move
            CK(cudaEventRecord(c->ev0, c->stream));
            cudaError_t le = cudaSuccess;
            if (n <= CTA_SMALL_N) { DISPATCH_METRIC(c->metric, (le = launch_cta_descents<M, 512>(c, S.order, cnt, next.as<int>(), dcost.as<long long>(), dmoves.as<long long>(), dpaths.as<int>(), dstatus.as<int>(), d_stop))); }
            else { DISPATCH_METRIC(c->metric, (le = launch_cta_descents<M, 1024>(c, S.order, cnt, next.as<int>(), dcost.as<long long>(), dmoves.as<long long>(), dpaths.as<int>(), dstatus.as<int>(), d_stop))); }
            CK(le);
            CK(cudaEventRecord(c->ev1, c->stream));
        }
        if (multi_path || cta_path) {
            c->stats.kernel_launches += 1;
            unsigned long long hcand = 0;
            // the clock is watched while the kernel runs; the results are fetched afterwards (a device-to-host copy into
            // pageable memory blocks the calling thread until everything queued before it has finished)
            e = wait_with_deadline(c, dl, &raised);
            if (e) return e;
            if (multi_path) {
                CK(cudaMemcpyAsync(hbc.data(), m_bcost.p, sizeof(long long) * g, cudaMemcpyDeviceToHost, c->stream));
                CK(cudaMemcpyAsync(hbs.data(), m_bstart.p, sizeof(int) * g, cudaMemcpyDeviceToHost, c->stream));
                CK(cudaMemcpyAsync(&hcand, m_cand.p, sizeof(hcand), cudaMemcpyDeviceToHost, c->stream));
            }
            CK(cudaMemcpyAsync(hf.data(), dcost.p, sizeof(long long) * cnt, cudaMemcpyDeviceToHost, c->stream));
            CK(cudaMemcpyAsync(hm.data(), dmoves.p, sizeof(long long) * cnt, cudaMemcpyDeviceToHost, c->stream));
            CK(cudaMemcpyAsync(hs.data(), dstatus.p, sizeof(int) * cnt, cudaMemcpyDeviceToHost, c->stream));
            CK(cudaStreamSynchronize(c->stream));
            clk.lap(clk.descents);
            c->stats.d2h_bytes += (int64_t)(sizeof(long long) * 2 + sizeof(int)) * cnt + 12 * g + 8;
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
            if (multi_path) { c->stats.pruned_kernel_ms += ms; c->stats.candidate_evals += (int64_t)hcand; }
            else c->stats.scan_kernel_ms += ms;
            c->stats.last_kernel_ms = ms;
            int bi = -1, taken = 0;
            for (int k = 0; k < cnt; k++) {
                if (nn_costs_out) nn_costs_out[done + k] = (double)hc[k];
                if (hs[k] == MS_NOT_STARTED) {                           // starts are handed out in ascending order: a suffix of the batch
                    cut_short = true;
                    if (final_costs_out) final_costs_out[done + k] = -1.0;
                    if (moves_out) moves_out[done + k] = -1;
                    continue;
                }
                ++taken;
                if (hs[k] == MS_CUT) cut_short = true;
                const long long scans = hm[k] + (hs[k] == MS_COMPLETE ? 1 : 0);
                c->stats.moves += hm[k];
                c->stats.scans += scans;
                if (multi_path) c->stats.pruned_scans += scans;
                else c->stats.pair_evals += scans * ((long long)n * (n - 1) / 2);
                if (final_costs_out) final_costs_out[done + k] = (double)hf[k];
                if (moves_out) moves_out[done + k] = hm[k];
                if (cta_path && (best_start < 0 || hf[k] < best)) { best = hf[k]; best_start = fs + k * stride; bi = k; }   // strict '<', tsp.c:671
            }
            if (multi_path) {                                            // (cost, start)-minimum over the teams' own bests
                int bslot = -1;
                for (int q = 0; q < g; q++) {
                    if (hbs[q] < 0) continue;
                    if (best_start < 0 || hbc[q] < best || (hbc[q] == best && hbs[q] < best_start)) { best = hbc[q]; best_start = hbs[q]; bslot = q; }
                }
                if (bslot >= 0)
                    CK(cudaMemcpyAsync(best_path.p, MS.best_path + (size_t)bslot * n, sizeof(int) * n, cudaMemcpyDeviceToDevice, c->stream));
            } else if (bi >= 0) {
                CK(cudaMemcpyAsync(best_path.p, dpaths.as<int>() + (size_t)bi * n, sizeof(int) * n, cudaMemcpyDeviceToDevice, c->stream));
            }
            if (cut_short) { done += taken; break; }
            continue;
        }
        int k = 0;
        for (; k < cnt; k++) {
            if (k > 0 && dl.passed()) { cut_short = true; break; }
            if (nn_costs_out) nn_costs_out[done + k] = (double)hc[k];
            // the NN visit order IS the tour-position order: no host round trip
            CK(cudaMemcpyAsync(c->d_order, S.order + (size_t)k * n, sizeof(int) * n, cudaMemcpyDeviceToDevice, c->stream));
            e = tour_from_device_order(c);
            if (e) return e;
            c->tour_loaded = false;
            bool dcut = false;
            e = descend(c, -1, 0, dl.left(), &dcut);
            if (e) return e;
            const long long fc = c->h_state->cost;
            if (final_costs_out) final_costs_out[done + k] = (double)fc;
            if (moves_out) moves_out[done + k] = c->h_state->moves;
            if (best_start < 0 || fc < best) {                           // strict '<', tsp.c:671
                best = fc; best_start = fs + k * stride;
                tour_export_kernel<<<grid_for(n, 4 * c->num_sms), 256, 0, c->stream>>>(c->tour, best_path.as<int>());
                c->stats.kernel_launches += 1;
                CK(cudaGetLastError());
            }
            if (dcut) { cut_short = true; ++k; break; }
        }
        if (cut_short) { done += k; break; }
    }
    done = std::min(done, nstarts);
    if (best_path_out && best_start >= 0) {
        CK(cudaMemcpyAsync(best_path_out, best_path.p, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
        c->stats.d2h_bytes += (int64_t)sizeof(int) * n;
    }
    CK(cudaStreamSynchronize(c->stream));
    if (raised) CK(cudaStreamSynchronize(c->aux_stream));
    clk.lap(clk.rest);
#ifdef TSP_NB_STATS
    if (multi_path) {
        unsigned long long h[8];
        cudaMemcpyFromSymbol(h, g_nbstat, sizeof(h));
        const double sc = (double)std::max<long long>(1, c->stats.pruned_scans);
        fprintf(stderr, "[tspcuda] K8 delta cache per scan: rebuilds %.1f (box nodes %.0f)  giants/no-room %.2f  heap slots %.0f (live %.0f)  dirty %.1f  evaluated %.1f  attend calls %.1f\n",
                h[0] / sc, h[7] / sc, h[1] / sc, h[2] / sc, h[4] / sc, h[3] / sc, h[5] / sc, h[6] / sc);
        unsigned long long z[8] = {0};
        cudaMemcpyToSymbol(g_nbstat, z, sizeof(z));
    }
#endif
    if (starts_done_out) *starts_done_out = done;
    if (best_start_out) *best_start_out = best_start;
    if (best_cost_out) *best_cost_out = (double)best;
    return (cut_short || done < nstarts) ? TSPCUDA_DEADLINE_EXCEEDED : TSPCUDA_OK;
}

// ---- K1 cost matrix ---------------------------------------------------------------------------
extern "C" int tspcuda_cost_matrix_device(tspcuda_ctx *c, double *ms_out) {
    int e = use(c);
    if (e) return e;
    const size_t n = (size_t)c->n;
    if (!c->d_matrix) CK(cudaMalloc(&c->d_matrix, sizeof(double) * n * n));
    // columns per work unit: 4 block-wide trips (1024 cells = 8 KB per row); one unit per CTA, handed out in index order by the
    // hardware (kernels.cuh).  For measurements: TSPCUDA_K1_CHUNK (trips, 0 = whole rows), TSPCUDA_K1_CTAS (> 0: that many
    // CTAs per SM dealing the units out grid-stride, the round-1/2 launch shape).
    int trips = 4, ctas = 0;
    if (const char *e2 = getenv("TSPCUDA_K1_CHUNK")) trips = atoi(e2);
    if (const char *e2 = getenv("TSPCUDA_K1_CTAS")) ctas = atoi(e2);
    const int chunk_pairs = trips > 0 ? trips * TSP_K1_BLOCK : (int)((n / 2 + TSP_K1_BLOCK) / TSP_K1_BLOCK) * TSP_K1_BLOCK;
    const size_t k1_s = (n & 1) ? 2 : 1, k1_span = TSP_K1_ROWS * k1_s;
    const size_t nunits = ((n + k1_span - 1) / k1_span) * k1_s * std::max<size_t>(1, (n / 2 + chunk_pairs - 1) / chunk_pairs);   // as the kernel counts them
    const int grid = (int)std::max<size_t>(1, std::min<size_t>(nunits, ctas > 0 ? (size_t)c->num_sms * ctas : (size_t)0x7fffffff));
    CK(cudaEventRecord(c->ev0, c->stream));
    DISPATCH_METRIC(c->metric, (cost_matrix_kernel<M><<<grid, TSP_K1_BLOCK, 0, c->stream>>>(c->d_xy, c->n, c->d_matrix, chunk_pairs)));
    CK(cudaEventRecord(c->ev1, c->stream));
    c->stats.kernel_launches += 1;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->stats.last_kernel_ms = ms;
    if (ms_out) *ms_out = ms;
    return TSPCUDA_OK;
}

extern "C" int tspcuda_cost_matrix(tspcuda_ctx *c, double *host_out) {
    if (!host_out) return fail(TSPCUDA_INVALID_ARGUMENT, "null output");
    int e = tspcuda_cost_matrix_device(c, nullptr);
    if (e) return e;
    const size_t n = (size_t)c->n;
    CK(cudaMemcpyAsync(host_out, c->d_matrix, sizeof(double) * n * n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->stats.d2h_bytes += (int64_t)(sizeof(double) * n * n);
    return TSPCUDA_OK;
}

extern "C" int tspcuda_cost_matrix_checksum(tspcuda_ctx *c, int64_t *sum_out) {
    int e = use(c);
    if (e) return e;
    if (!c->d_matrix) return fail(TSPCUDA_INVALID_ARGUMENT, "matrix not built: call tspcuda_cost_matrix_device first");
    const unsigned long long cells = (unsigned long long)c->n * (unsigned long long)c->n;
    CK(cudaMemsetAsync(c->d_sum, 0, sizeof(long long), c->stream));
    matrix_checksum_kernel<<<c->num_sms * 8, 256, 0, c->stream>>>(c->d_matrix, cells, c->d_sum);
    c->stats.kernel_launches += 1;
    CK(cudaGetLastError());
    long long h = 0;
    CK(cudaMemcpyAsync(&h, c->d_sum, sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (sum_out) *sum_out = h;
    return TSPCUDA_OK;
}

// ---- measurement helpers ----------------------------------------------------------------------
template <int KIND>
static cudaError_t run_pipe(tspcuda_ctx *c, int iters, double *d_sink) {
    pipe_rate_kernel<KIND><<<c->num_sms * 8, 256, 0, c->stream>>>(iters, 1.5, d_sink);
    return cudaGetLastError();
}

extern "C" int tspcuda_pipe_rate(tspcuda_ctx *c, int kind, double *ops_per_s_out) {
    int e = use(c);
    if (e) return e;
    if (kind == 100 && ops_per_s_out) {
        // what a pure writer reaches on this device: the driver's own fill of the HBM-resident matrix buffer, in bytes/s
        if (!c->d_matrix) return fail(TSPCUDA_INVALID_ARGUMENT, "matrix not built: call tspcuda_cost_matrix_device first");
        const size_t bytes = sizeof(double) * (size_t)c->n * (size_t)c->n;
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            CK(cudaEventRecord(c->ev0, c->stream));
            CK(cudaMemsetAsync(c->d_matrix, 0, bytes, c->stream));
            CK(cudaEventRecord(c->ev1, c->stream));
            CK(cudaStreamSynchronize(c->stream));
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
            if (rep > 0) best = std::min(best, ms);
        }
        *ops_per_s_out = (double)bytes / (best * 1e-3);
        return TSPCUDA_OK;
    }
    if (kind == 101 && ops_per_s_out) {
        // mismatches of the FP32-only cost sequence (kind 10) against the shipped one over 1.5e9 integer coordinate triples
        unsigned long long *d_bad = nullptr, h_bad = 0;
        CK(cudaMalloc(&d_bad, sizeof(unsigned long long)));
        CK(cudaMemsetAsync(d_bad, 0, sizeof(unsigned long long), c->stream));
        f32cost_selfcheck_kernel<<<c->num_sms * 8, 256, 0, c->stream>>>(12345u, 5000, d_bad);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(&h_bad, d_bad, sizeof(h_bad), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        cudaFree(d_bad);
        *ops_per_s_out = (double)h_bad;
        return TSPCUDA_OK;
    }
    if (kind < 0 || kind > 10 || !ops_per_s_out) return fail(TSPCUDA_INVALID_ARGUMENT, "kind");
    double *d_sink = nullptr;
    CK(cudaMalloc(&d_sink, sizeof(double)));
    const int iters = 4096;
    float best_ms = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        CK(cudaEventRecord(c->ev0, c->stream));
        cudaError_t le;
        switch (kind) {
            case 0: le = run_pipe<0>(c, iters, d_sink); break;
            case 1: le = run_pipe<1>(c, iters, d_sink); break;
            case 2: le = run_pipe<2>(c, iters, d_sink); break;
            case 3: le = run_pipe<3>(c, iters, d_sink); break;
            case 4: le = run_pipe<4>(c, iters, d_sink); break;
            case 5: le = run_pipe<5>(c, iters, d_sink); break;
            case 6: le = run_pipe<6>(c, iters, d_sink); break;
            case 7: le = run_pipe<7>(c, iters, d_sink); break;
            case 8: le = run_pipe<8>(c, iters, d_sink); break;
            case 10: le = run_pipe<10>(c, iters, d_sink); break;
            default: le = run_pipe<9>(c, iters, d_sink); break;
        }
        CK(le);
        CK(cudaEventRecord(c->ev1, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        c->stats.kernel_launches += 1;
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        if (rep > 0) best_ms = std::min(best_ms, ms);
    }
    cudaFree(d_sink);
    const double per_iter = (kind == 3 || kind == 9 || kind == 10) ? 16.0 : 8.0;   // kind 3: two conversions, kind 9: two edge costs per chain step       // kind 3 = F2F.F32.F64 + F2F.F64.F32 per chain step
    const double ops = (double)c->num_sms * 8 * 256 * (double)iters * per_iter;
    *ops_per_s_out = ops / (best_ms * 1e-3);
    return TSPCUDA_OK;
}

extern "C" int tspcuda_selfcheck_sqrt(tspcuda_ctx *c, int64_t *mismatches_out) {
    int e = use(c);
    if (e) return e;
    unsigned long long *d_bad = nullptr;
    CK(cudaMalloc(&d_bad, sizeof(unsigned long long)));
    CK(cudaMemsetAsync(d_bad, 0, sizeof(unsigned long long), c->stream));
    // floats in [2^-60, 2^64): bit patterns [ (127-60)<<23, (127+64)<<23 )
    const unsigned lo = (unsigned)(127 - 60) << 23, hi = (unsigned)(127 + 64) << 23;
    sqrt_selfcheck_kernel<<<c->num_sms * 16, 256, 0, c->stream>>>(lo, hi, d_bad);
    c->stats.kernel_launches += 1;
    CK(cudaGetLastError());
    unsigned long long h = 0;
    CK(cudaMemcpyAsync(&h, d_bad, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    cudaFree(d_bad);
    if (mismatches_out) *mismatches_out = (int64_t)h;
    return TSPCUDA_OK;
}

extern "C" int tspcuda_timer_start(tspcuda_ctx *c) {
    int e = use(c);
    if (e) return e;
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaEventRecord(c->tev0, c->stream));
    return TSPCUDA_OK;
}

extern "C" int tspcuda_timer_stop(tspcuda_ctx *c, double *ms_out) {
    int e = use(c);
    if (e) return e;
    CK(cudaEventRecord(c->tev1, c->stream));
    CK(cudaEventSynchronize(c->tev1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, c->tev0, c->tev1));
    if (ms_out) *ms_out = ms;
    return TSPCUDA_OK;
}

extern "C" int tspcuda_flush_l2(tspcuda_ctx *c) {
    int e = use(c);
    if (e) return e;
    const size_t bytes = 192u << 20;
    if (!c->d_flush) CK(cudaMalloc(&c->d_flush, bytes));
    CK(cudaMemsetAsync(c->d_flush, 0x5A, bytes, c->stream));
    return TSPCUDA_OK;
}


// ---- Extra Mileage (heuristics.c:156-210, :290-367) ---------------------------------------------
constexpr int EM_BLOCK = 256;

extern "C" int tspcuda_extra_mileage(tspcuda_ctx *c, int node_a, int node_b, int *path_out, double *cost_out,
                                     int *node_a_out, int *node_b_out) {
    int e = use(c);
    if (e) return e;
    const int n = c->n;
    if (n < 2) return fail(TSPCUDA_INVALID_ARGUMENT, "extra mileage needs at least two nodes");
    if (node_a >= 0 && (node_a >= n || node_b < 0 || node_b >= n || node_a == node_b))
        return fail(TSPCUDA_INVALID_ARGUMENT, "initial edge (%d,%d)", node_a, node_b);
    DevBuf vis, eu, ev, best, bedge, list, key, lc, cost, pair, rk, done;
    CK(cudaMalloc(&vis.p, sizeof(int) * n)); CK(cudaMalloc(&eu.p, sizeof(int) * n)); CK(cudaMalloc(&ev.p, sizeof(int) * n));
    CK(cudaMalloc(&best.p, sizeof(int) * n)); CK(cudaMalloc(&bedge.p, sizeof(int) * n)); CK(cudaMalloc(&list.p, sizeof(int) * n));
    CK(cudaMalloc(&key.p, sizeof(unsigned long long) * 2)); CK(cudaMalloc(&lc.p, sizeof(int) * 2));
    CK(cudaMalloc(&cost.p, sizeof(long long))); CK(cudaMalloc(&pair.p, sizeof(int) * 2));
    CK(cudaMalloc(&rk.p, sizeof(unsigned long long) * n)); CK(cudaMalloc(&done.p, sizeof(int)));
    CK(cudaMemsetAsync(rk.p, 0xFF, sizeof(unsigned long long) * n, c->stream));
    CK(cudaMemsetAsync(done.p, 0, sizeof(int), c->stream));
    int h_pair[2] = {node_a, node_b};
    CK(cudaEventRecord(c->ev0, c->stream));
    if (node_a < 0) {                                   // EM_MAX, heuristics.c:169-178
        CK(cudaMemsetAsync(key.p, 0, sizeof(unsigned long long), c->stream));
        DISPATCH_METRIC(c->metric, (em_max_kernel<M><<<c->num_sms * 8, 256, 0, c->stream>>>(c->d_xy, n, key.as<unsigned long long>())));
        c->stats.kernel_launches += 1;
        CK(cudaGetLastError());
        unsigned long long k = 0;
        CK(cudaMemcpyAsync(&k, key.p, sizeof(k), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        const unsigned long long ij = 0x3FFFFFFFFull - (k & 0x3FFFFFFFFull);
        h_pair[0] = k ? (int)(ij >> 17) : 0;
        h_pair[1] = k ? (int)(ij & 0x1FFFF) : 1;
    }
    CK(cudaMemcpyAsync(pair.p, h_pair, sizeof(h_pair), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemsetAsync(key.p, 0xFF, sizeof(unsigned long long) * 2, c->stream));
    EmState S;
    S.visited = vis.as<int>(); S.eu = eu.as<int>(); S.ev = ev.as<int>(); S.best = best.as<int>(); S.bedge = bedge.as<int>();
    S.path = c->d_path; S.list = list.as<int>(); S.key = key.as<unsigned long long>(); S.list_count = lc.as<int>();
    S.cost = cost.as<long long>(); S.init_pair = pair.as<int>();
    S.rk = rk.as<unsigned long long>(); S.done = done.as<int>();
    int bps = 0;
    DISPATCH_METRIC(c->metric, (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, extra_mileage_kernel<M, EM_BLOCK>, EM_BLOCK, 0)));
    if (bps <= 0) return fail(TSPCUDA_INTERNAL, "extra mileage kernel cannot be resident");
    const int grid = std::max(1, std::min(std::min(bps, 2) * c->num_sms, (n + EM_BLOCK - 1) / EM_BLOCK));
    const double2 *xy = c->d_xy;
    int nn = n;
    void *args[] = {&xy, &nn, &S};
    cudaError_t le = cudaSuccess;
    DISPATCH_METRIC(c->metric, (le = cudaLaunchCooperativeKernel((void *)extra_mileage_kernel<M, EM_BLOCK>, dim3(grid), dim3(EM_BLOCK), args, 0, c->stream)));
    CK(le);
    CK(cudaEventRecord(c->ev1, c->stream));
    c->stats.kernel_launches += 1;
    long long h_cost = 0;
    CK(cudaMemcpyAsync(&h_cost, cost.p, sizeof(h_cost), cudaMemcpyDeviceToHost, c->stream));
    if (path_out) CK(cudaMemcpyAsync(c->h_buf, c->d_path, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->stats.last_kernel_ms = ms;
    c->stats.d2h_bytes += 8 + (path_out ? (int64_t)sizeof(int) * n : 0);
    c->tour_loaded = false;
    if (path_out) std::memcpy(path_out, c->h_buf, sizeof(int) * n);
    if (cost_out) *cost_out = (double)h_cost;
    if (node_a_out) *node_a_out = h_pair[0];
    if (node_b_out) *node_b_out = h_pair[1];
    return TSPCUDA_OK;
}

// ---- VNS loop (metaheuristic.c:277-321 of mh_VNS; vns_kick :344-409, tabu_make_move case 7 :490-500) ----------
// The kick, on the tour ORDER from node 0 (what vns_kick rebuilds from the successor array at every call): three
// random positions i0 < i1 < i2 drawn from the process's rand() stream with the reference's rejection test --
// which compares a drawn POSITION with the NODES next to the positions accepted before (:372) -- then
// path[A] = succ B, path[C] = succ A, path[B] = succ C, i.e. the blocks (i0, i1] and (i1, i2] swap places.
// The reference reads tour[-1] / tour[n] out of bounds in that test; under glibc both read 0 unless n = 2 (mod 4)
// (checked against the reference objects on small n, where the case is frequent), and 0 is
// what is used here.
// Where the random numbers of a VNS run come from.  The reference draws from the process-wide glibc stream (rand(),
// never seeded on file instances), and so does tspcuda_vns_search.  The restarts entry points give every restart its OWN
// generator instead -- glibc's reentrant random_r with the 128-byte state srand()/rand() use, i.e. the SAME sequence as
// `srand(seed); rand(); ...` -- so that restarts running on several GPUs (worker threads) at once stay reproducible.
struct RandSource {
    bool own = false;
    struct random_data rd;
    char state[128];
    RandSource() { std::memset(&rd, 0, sizeof(rd)); std::memset(state, 0, sizeof(state)); }
    void seed(unsigned s) { own = true; std::memset(&rd, 0, sizeof(rd)); initstate_r(s, state, sizeof(state), &rd); }
    int next() {
        if (!own) return rand();
        int32_t r = 0;
        random_r(&rd, &r);
        return (int)r;
    }
};

static void vns_kick_order(int n, int *ord, int *tmp, RandSource &rng) {
    int idx[3];
    for (int i = 0; i < 3; i++) {
        int r;
        do {
            r = rng.next() % n;
            for (int j = 0; j < i; j++) {
                const int lo = idx[j] - 1, hi = idx[j] + 1;
                const int tlo = lo >= 0 ? ord[lo] : 0, thi = hi < n ? ord[hi] : 0;
                if (r == idx[j] || r == tlo || r == thi) { r = -1; break; }
            }
        } while (r == -1);
        idx[i] = r;
        for (int j = i; j > 0; j--)
            if (idx[j] < idx[j - 1]) std::swap(idx[j], idx[j - 1]);
    }
    const int i0 = idx[0], i1 = idx[1], i2 = idx[2];
    const int len1 = i1 - i0, len2 = i2 - i1;                  // blocks (i0, i1] and (i1, i2]
    std::memcpy(tmp, ord + i0 + 1, sizeof(int) * len1);
    std::memmove(ord + i0 + 1, ord + i1 + 1, sizeof(int) * len2);
    std::memcpy(ord + i0 + 1 + len2, tmp, sizeof(int) * len1);
}

static int vns_search_impl(tspcuda_ctx *c, int *path, double *cost_inout, int n_iters, double time_limit_s,
                           double *cost_trace_out, int *best_path_out, double *best_cost_inout, int *best_iter_out,
                           int *iters_done_out, RandSource &rng);

extern "C" int tspcuda_vns_search(tspcuda_ctx *c, int *path, double *cost_inout, int n_iters, double time_limit_s,
                                  double *cost_trace_out, int *best_path_out, double *best_cost_inout, int *best_iter_out,
                                  int *iters_done_out) {
    RandSource process_stream;                                 // rand(): the caller's (the reference driver's) own stream
    return vns_search_impl(c, path, cost_inout, n_iters, time_limit_s, cost_trace_out, best_path_out, best_cost_inout, best_iter_out,
                           iters_done_out, process_stream);
}

static int vns_search_impl(tspcuda_ctx *c, int *path, double *cost_inout, int n_iters, double time_limit_s,
                           double *cost_trace_out, int *best_path_out, double *best_cost_inout, int *best_iter_out,
                           int *iters_done_out, RandSource &rng) {
    int e = use(c);
    if (e) return e;
    if (!path || !best_cost_inout) return fail(TSPCUDA_INVALID_ARGUMENT, "null argument");
    const int n = c->n;
    if (n < 5) return fail(TSPCUDA_INVALID_ARGUMENT, "the kick needs three non-adjacent positions");
    if (best_iter_out) *best_iter_out = -1;
    if (iters_done_out) *iters_done_out = 0;
    c->tour_loaded = false;
    e = tspcuda_tour_upload(c, path);                          // validates the tour; device order starts at node 0
    if (e) return e;
    // the descents of this loop start from kicked local optima: giant search boxes are evaluated, not cached, the first few
    // times (pruned.cuh, NbCache::giant_age; TSPCUDA_NB_GIANT_AGE for measurements)
    struct AgeGuard { tspcuda_ctx *c; ~AgeGuard() { c->nb_giant_age = 0; } } age_guard{c};
    c->nb_giant_age = 5;
    if (const char *ga = getenv("TSPCUDA_NB_GIANT_AGE")) c->nb_giant_age = std::max(0, std::min(1000, atoi(ga)));
    std::vector<int> ord(n), tmp(n), best_ord, before(n), where(n), changed;
    const auto t_begin = std::chrono::steady_clock::now();
    auto elapsed = [&]() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count(); };
    bool deadline = false;
    int done = 0, best_iter = -1;
    double cost = cost_inout ? *cost_inout : 0.0;
    for (int i = 0; i < n_iters; i++) {
        if (time_limit_s >= 0 && elapsed() > time_limit_s) { deadline = true; break; }     // :279-285
        bool dl = false;
        e = descend(c, -1, 0, time_limit_s >= 0 ? std::max(0.0, time_limit_s - elapsed()) : -1.0, &dl);   // ref_2opt, :288
        if (e) return e;
        CK(cudaMemcpyAsync(c->h_buf, c->tour.tid, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
        e = read_state(c);                                     // synchronises
        if (e) return e;
        c->stats.d2h_bytes += (int64_t)sizeof(int) * n;
        cost = (double)c->h_state->cost;
        int p0 = 0;
        while (p0 < n && c->h_buf[p0] != 0) ++p0;              // rotate: the reference's tour[] starts at node 0 (:356-363)
        if (p0 == n) return fail(TSPCUDA_INTERNAL, "node 0 missing from the device tour");
        std::memcpy(ord.data(), c->h_buf + p0, sizeof(int) * (n - p0));
        std::memcpy(ord.data() + (n - p0), c->h_buf, sizeof(int) * p0);
        if (cost < *best_cost_inout) {                         // :296-300, strict '<'
            *best_cost_inout = cost;
            best_ord = ord;
            best_iter = i;
        }
        if (cost_trace_out) cost_trace_out[i] = cost;          // :303
        done = i + 1;
        const int r = rng.next() % (10 - 2 + 1) - 2;           // :306, UPPER 10 / LOWER 2 (metaheuristic.h:7-8)
        if (r > 0) before = ord;
        for (int j = 0; j < r; j++) vns_kick_order(n, ord.data(), tmp.data(), rng);
        std::memcpy(c->h_buf, ord.data(), sizeof(int) * n);
        CK(cudaMemcpyAsync(c->d_order, c->h_buf, sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
        c->stats.h2d_bytes += (int64_t)sizeof(int) * n;
        // A kick exchanges two blocks without turning either round: it changes three tour edges, i.e. the neighbours of at
        // most six nodes, and no node's direction.  The device's delta cache (pruned.cuh) depends, per record (u,w), on the
        // neighbours of u and w only, so only the nodes whose neighbour pair differs between the tour before and after the
        // kicks are marked (their lists stale, the records pointing at them dirty) -- the descent after a kick then costs
        // a handful of searches instead of one per node.
        const bool keep = c->nb_valid;
        changed.clear();
        if (keep && r > 0) {
            for (int k = 0; k < n; k++) where[before[k]] = k;
            for (int k = 0; k < n; k++) {
                const int u = ord[k], kb = where[u];
                const int p1 = ord[k == 0 ? n - 1 : k - 1], s1 = ord[k + 1 == n ? 0 : k + 1];
                const int p0 = before[kb == 0 ? n - 1 : kb - 1], s0 = before[kb + 1 == n ? 0 : kb + 1];
                if (!(p1 == p0 && s1 == s0)) changed.push_back(u);     // same neighbours in the same direction (the cache's o8 bits stay valid)
            }
        }
        e = tour_from_device_order(c, keep);                   // cost recomputed from the kicked tour (refinment.c:6-9)
        if (e) return e;
        if (!changed.empty()) {
            DevBuf d_list;
            CK(scratch_get(c, SC_NB_LIST, sizeof(int) * (size_t)n, d_list));
            CK(cudaMemcpyAsync(d_list.p, changed.data(), sizeof(int) * changed.size(), cudaMemcpyHostToDevice, c->stream));
            NbCache nb;
            nb.lines = c->d_nb; nb.heap = c->d_nb_heap; nb.heap_next = c->d_nb_next; nb.cnt8 = c->d_cnt8; nb.heap_cap = nb_heap_cap(n, c->nb_heap_per_node); nb.giant_age = 0; nb.restart_pct = c->nb_restart_pct;
            const int cn = (int)changed.size();
            nb_mark_nodes_kernel<<<1, 256, 0, c->stream>>>(nb, d_list.as<int>(), cn, c->d_nbflag, 1);
            nb_mark_records_kernel<<<grid_for(n, 4 * c->num_sms), 256, 0, c->stream>>>(nb, c->d_nbflag, n);
            nb_mark_nodes_kernel<<<1, 256, 0, c->stream>>>(nb, d_list.as<int>(), cn, c->d_nbflag, 0);
            c->stats.kernel_launches += 3;
            c->stats.h2d_bytes += (int64_t)sizeof(int) * changed.size();
            CK(cudaGetLastError());
        }
        CK(cudaStreamSynchronize(c->stream));                  // h_buf and `changed` are reused
    }
    // the solution as the reference leaves it: the tour after the last kicks, cost = after the last descent
    if (done > 0)
        for (int k = 0; k < n; k++) path[ord[k]] = ord[(k + 1) % n];
    c->tour_loaded = false;
    if (cost_inout) *cost_inout = cost;
    if (best_iter >= 0 && best_path_out)
        for (int k = 0; k < n; k++) best_path_out[best_ord[k]] = best_ord[(k + 1) % n];
    if (best_iter_out) *best_iter_out = best_iter;
    if (iters_done_out) *iters_done_out = done;
    return deadline ? TSPCUDA_DEADLINE_EXCEEDED : TSPCUDA_OK;
}

// ---- config 5 "restarts" (BASELINE.json configs[4]) -----------------------------------------------
// A framework addition: the reference runs ONE tabu / VNS trajectory.  Restart r = start node first_start + r*stride, each
// an independent unit (it depends on the instance, its start node and the seed only), so restarts shard over GPUs exactly
// like the multi-start drivers and merge under the same incumbent rule (strict '<' in ascending start order, tsp.c:671).
//   kind 0, tabu: NN(start) -> 2-opt local optimum (ref_2opt) -> `iters` iterations of the loop of mh_TabuSearch
//                 (metaheuristic.c:128-164) with the reference's default linear tenure schedule (tabu_init :65-84,
//                 tabu_linear_policy :40-59), incumbent = the local optimum
//   kind 1, VNS:  NN(start) -> `iters` iterations of the loop of mh_VNS (:277-321), kicks drawn from a private generator
//                 seeded as srand(seed + start) would seed the process stream; no incumbent to beat (the first local
//                 optimum always enters)
static void linear_tenures(int n, int k, std::vector<int> &out) {
    int tenure = (int)(0.125 * n + 1);                          // tabu_init: MIN_FRACTION * nnodes + 1, truncated by the int field
    const int hi = (int)(0.25 * n), lo = (int)(0.125 * n);
    bool inc = true;
    out.resize(k);
    for (int i = 0; i < k; i++) {
        if (hi == tenure || lo == tenure) inc = !inc;
        tenure += inc ? 1 : -1;
        out[i] = tenure;
    }
}

extern "C" int tspcuda_restarts(tspcuda_ctx *c, int kind, int first_start, int nstarts, int stride, int iters, unsigned seed,
                                double time_limit_s, int *best_start_out, int *best_path_out, double *best_cost_out,
                                int *restarts_done_out) {
    int e = use(c);
    if (e) return e;
    const int n = c->n;
    if (restarts_done_out) *restarts_done_out = 0;
    if (best_start_out) *best_start_out = -1;
    if (kind != TSPCUDA_RESTART_TABU && kind != TSPCUDA_RESTART_VNS) return fail(TSPCUDA_INVALID_ARGUMENT, "restart kind %d", kind);
    if (nstarts <= 0 || stride <= 0 || iters < 0) return fail(TSPCUDA_INVALID_ARGUMENT, "nstarts/stride/iters");
    const long long last = (long long)first_start + (long long)(nstarts - 1) * stride;
    if (first_start < 0 || last >= n) return fail(TSPCUDA_UNAVAILABLE, "starting node not correct");
    const Deadline dl(time_limit_s);
    std::vector<int> path(n), run_best(n), tabu(n), bestp(n), tenures;
    if (kind == TSPCUDA_RESTART_TABU) linear_tenures(n, iters, tenures);
    double best = 0;
    int best_start = -1, done = 0;
    bool late = false;
    for (int r = 0; r < nstarts; r++) {
        if (dl.passed()) { late = true; break; }
        const int start = first_start + r * stride;
        double cost = 0, run_cost = 0;
        e = tspcuda_nn(c, start, path.data(), &cost);
        if (e) return e;
        bool have = false;
        if (kind == TSPCUDA_RESTART_TABU) {
            int64_t mv = 0;
            e = tspcuda_two_opt(c, path.data(), &cost, -1, dl.left(), &mv, nullptr, 0);
            if (e == TSPCUDA_DEADLINE_EXCEEDED) late = true;
            else if (e) return e;
            run_cost = cost;
            run_best = path;
            have = true;
            if (!late && iters > 0) {
                std::fill(tabu.begin(), tabu.end(), -1);
                double bc = cost;
                int bi = -1;
                e = tspcuda_tabu_search(c, path.data(), &cost, tabu.data(), tenures.data(), 0, iters, run_cost, nullptr, bestp.data(), &bc, &bi);
                if (e) return e;
                if (bi >= 0) {                                  // some iteration improved on the local optimum
                    run_cost = bc;
                    run_best = bestp;
                }
            }
        } else {
            RandSource rng;
            rng.seed(seed + (unsigned)start);
            double bc = __DBL_MAX__;
            int bi = -1, it_done = 0;
            e = vns_search_impl(c, path.data(), &cost, iters, dl.left(), nullptr, run_best.data(), &bc, &bi, &it_done, rng);
            if (e == TSPCUDA_DEADLINE_EXCEEDED) late = true;
            else if (e) return e;
            if (bi >= 0) { run_cost = bc; have = true; }
        }
        if (have && (best_start < 0 || run_cost < best)) {      // strict '<' in ascending start order
            best = run_cost;
            best_start = start;
            if (best_path_out) std::memcpy(best_path_out, run_best.data(), sizeof(int) * n);
        }
        ++done;
        if (late) break;
    }
    if (restarts_done_out) *restarts_done_out = done;
    if (best_start_out) *best_start_out = best_start;
    if (best_cost_out && best_start >= 0) *best_cost_out = best;
    return (late || done < nstarts) ? TSPCUDA_DEADLINE_EXCEEDED : TSPCUDA_OK;
}

// ---- patching (cplex_model.c:693-755 cx_patching, the Benders loop's while (ncomp > 1), :205-212) -----
extern "C" int tspcuda_patching(tspcuda_ctx *c, int *path, int *comp, int *ncomp_inout, double *cost_inout,
                                int max_patches, tspcuda_move *log, int64_t log_cap, int *patches_out) {
    int e = use(c);
    if (e) return e;
    const int n = c->n;
    if (!path || !comp || !ncomp_inout) return fail(TSPCUDA_INVALID_ARGUMENT, "null argument");
    if (patches_out) *patches_out = 0;
    // the successor array must be a permutation whose cycles carry one label each, distinct between cycles
    // (cx_build_sol's output); the relabelling below relies on it
    {
        std::vector<char> seen(n, 0);
        std::vector<int> labels;
        int cycles = 0;
        for (int v = 0; v < n; ++v) {
            if (seen[v]) continue;
            int len = 0;
            for (int u = v; !seen[u]; u = path[u]) {
                seen[u] = 1;
                ++len;
                if (comp[u] != comp[v]) return fail(TSPCUDA_INVALID_ARGUMENT, "comp[] is not constant on the cycle of node %d", v);
                if (path[u] < 0 || path[u] >= n) return fail(TSPCUDA_INVALID_ARGUMENT, "successor of node %d out of range", u);
                if (seen[path[u]] && path[u] != v) return fail(TSPCUDA_INVALID_ARGUMENT, "path[] is not a permutation (node %d)", u);
            }
            if (len < 2) return fail(TSPCUDA_INVALID_ARGUMENT, "node %d is its own successor", v);
            labels.push_back(comp[v]);
            ++cycles;
        }
        std::sort(labels.begin(), labels.end());
        if (std::adjacent_find(labels.begin(), labels.end()) != labels.end())
            return fail(TSPCUDA_INVALID_ARGUMENT, "two cycles share a component label");
        if (cycles != *ncomp_inout) return fail(TSPCUDA_INVALID_ARGUMENT, "ncomp = %d but path[] has %d cycles", *ncomp_inout, cycles);
    }
    int todo = *ncomp_inout - 1;
    if (max_patches >= 0) todo = std::min(todo, max_patches);
    if (todo <= 0) return TSPCUDA_OK;
    DevBuf dpath, dcomp, sxy, dn, key, cost, info, dlog;
    CK(cudaMalloc(&dpath.p, sizeof(int) * n)); CK(cudaMalloc(&dcomp.p, sizeof(int) * n));
    CK(cudaMalloc(&sxy.p, sizeof(double2) * n)); CK(cudaMalloc(&dn.p, sizeof(int) * n));
    CK(cudaMalloc(&key.p, sizeof(unsigned long long))); CK(cudaMalloc(&cost.p, sizeof(long long)));
    CK(cudaMalloc(&info.p, sizeof(int) * 2)); CK(cudaMalloc(&dlog.p, sizeof(int) * 3 * (size_t)todo));
    CK(cudaMemcpyAsync(dpath.p, path, sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(dcomp.p, comp, sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemsetAsync(key.p, 0xFF, sizeof(unsigned long long), c->stream));
    CK(cudaMemsetAsync(cost.p, 0, sizeof(long long), c->stream));
    CK(cudaMemsetAsync(info.p, 0, sizeof(int) * 2, c->stream));
    c->stats.h2d_bytes += 2 * (int64_t)sizeof(int) * n;
    PatchState S;
    S.path = dpath.as<int>(); S.comp = dcomp.as<int>(); S.sxy = sxy.as<double2>(); S.dn = dn.as<int>();
    S.key = key.as<unsigned long long>(); S.cost = cost.as<long long>(); S.info = info.as<int>();
    S.log = dlog.as<int>(); S.log_cap = todo;
    // small instances: 64-node tiles, so that the triangle still yields several hundred tile pairs to spread over the SMs
    const int tile = n < 12000 ? 64 : 256;
    const long long ntiles = (n + tile - 1) / tile;
    const int scan_grid = (int)std::min<long long>(ntiles * (ntiles + 1) / 2, (long long)c->num_sms * (tile == 64 ? 32 : 8));
    CK(cudaEventRecord(c->ev0, c->stream));
    DISPATCH_METRIC(c->metric, (patch_prep_kernel<M><<<grid_for(n, c->num_sms * 4), 256, 0, c->stream>>>(c->d_xy, n, S)));
    for (int k = 0; k < todo; ++k) {
        if (tile == 64) { DISPATCH_METRIC(c->metric, (patch_scan_kernel<M, 64><<<scan_grid, 64, 0, c->stream>>>(c->d_xy, n, S))); }
        else { DISPATCH_METRIC(c->metric, (patch_scan_kernel<M, 256><<<scan_grid, 256, 0, c->stream>>>(c->d_xy, n, S))); }
        DISPATCH_METRIC(c->metric, (patch_apply_kernel<M><<<1, 1024, 0, c->stream>>>(c->d_xy, n, S)));
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(c->ev1, c->stream));
    c->stats.kernel_launches += 1 + 2 * (int64_t)todo;
    int h_info[2] = {0, 0};
    std::vector<int> h_log(3 * (size_t)todo);
    CK(cudaMemcpyAsync(h_info, info.p, sizeof(h_info), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(h_log.data(), dlog.p, sizeof(int) * h_log.size(), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(path, dpath.p, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(comp, dcomp.p, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->stats.last_kernel_ms = ms;
    c->stats.d2h_bytes += 2 * (int64_t)sizeof(int) * n + (int64_t)sizeof(int) * h_log.size() + 8;
    c->tour_loaded = false;
    const int done = h_info[0];
    for (int k = 0; k < done; ++k) {
        if (cost_inout) *cost_inout += (double)h_log[3 * k + 2];          // cplex_model.c:738, patch by patch
        if (log && k < log_cap) { log[k].a = h_log[3 * k]; log[k].b = h_log[3 * k + 1]; log[k].delta = h_log[3 * k + 2]; log[k].reserved = 0; }
    }
    *ncomp_inout -= done;
    if (patches_out) *patches_out = done;
    if (h_info[1])
        return fail(TSPCUDA_FAILED_PRECONDITION, "no pair of nodes in different components has delta > 0 after %d patches "
                    "(the reference indexes path[-1] here, cplex_model.c:730-735)", done);
    return TSPCUDA_OK;
}

// ===== the `costs`-argument paths: a caller-supplied dense matrix (matrix.cuh) ========================
struct tspcuda_mx {
    int n = 0, device = 0, num_sms = 0, coop_blocks = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double *d_C = nullptr, *d_dnp = nullptr, *d_nncost = nullptr;
    int *d_tid = nullptr, *d_pos = nullptr, *d_order = nullptr, *d_path = nullptr, *d_len = nullptr;
    unsigned char *d_visited = nullptr;
    unsigned long long *d_slots = nullptr;
    MxState *d_state = nullptr, *h_state = nullptr;
    MxMove *d_log = nullptr;
    long long log_cap = 0;
    int *h_buf = nullptr;
    tspcuda_stats stats{};
};

extern "C" void tspcuda_mx_destroy(tspcuda_mx *m) {
    if (!m) return;
    cudaSetDevice(m->device);
    if (m->stream) cudaStreamSynchronize(m->stream);
    cudaFree(m->d_C); cudaFree(m->d_dnp); cudaFree(m->d_nncost); cudaFree(m->d_tid); cudaFree(m->d_pos);
    cudaFree(m->d_order); cudaFree(m->d_path); cudaFree(m->d_len); cudaFree(m->d_visited); cudaFree(m->d_slots);
    cudaFree(m->d_state); cudaFree(m->d_log);
    if (m->h_state) cudaFreeHost(m->h_state);
    if (m->h_buf) cudaFreeHost(m->h_buf);
    if (m->ev0) cudaEventDestroy(m->ev0);
    if (m->ev1) cudaEventDestroy(m->ev1);
    if (m->stream) cudaStreamDestroy(m->stream);
    delete m;
}

extern "C" int tspcuda_mx_create(tspcuda_mx **out, const double *costs, int n, int device) {
    if (!out || !costs) return fail(TSPCUDA_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    if (n <= 0) return fail(TSPCUDA_INVALID_ARGUMENT, "n = %d", n);
    int ndev = tspcuda_device_count();
    if (ndev <= 0) return fail(TSPCUDA_UNAVAILABLE, "no CUDA device: libtspcuda has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(TSPCUDA_INVALID_ARGUMENT, "device %d of %d", device, ndev);
    tspcuda_mx *m = new tspcuda_mx();
    m->n = n; m->device = device;
#define CKM(call)                                                                                     \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            int code_ = fail(e_ == cudaErrorMemoryAllocation ? TSPCUDA_RESOURCE_EXHAUSTED : TSPCUDA_INTERNAL, \
                             "%s failed: %s", #call, cudaGetErrorString(e_));                          \
            tspcuda_mx_destroy(m);                                                                    \
            return code_;                                                                             \
        }                                                                                             \
    } while (0)
    CKM(cudaSetDevice(device));
    cudaDeviceProp prop;
    CKM(cudaGetDeviceProperties(&prop, device));
    m->num_sms = prop.multiProcessorCount;
    int bps = 0;
    CKM(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, descend_matrix_kernel, MX_BLOCK, 0));
    if (bps <= 0) { tspcuda_mx_destroy(m); return fail(TSPCUDA_INTERNAL, "matrix descent kernel cannot be resident"); }
    m->coop_blocks = std::min(bps, 4) * m->num_sms;
    CKM(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
    CKM(cudaEventCreate(&m->ev0));
    CKM(cudaEventCreate(&m->ev1));
    const size_t N = (size_t)n;
    CKM(cudaMalloc(&m->d_C, sizeof(double) * N * N));
    CKM(cudaMalloc(&m->d_dnp, sizeof(double) * (N + 1)));
    CKM(cudaMalloc(&m->d_nncost, sizeof(double)));
    CKM(cudaMalloc(&m->d_tid, sizeof(int) * (N + 1)));
    CKM(cudaMalloc(&m->d_pos, sizeof(int) * N));
    CKM(cudaMalloc(&m->d_order, sizeof(int) * N));
    CKM(cudaMalloc(&m->d_path, sizeof(int) * N));
    CKM(cudaMalloc(&m->d_len, sizeof(int)));
    CKM(cudaMalloc(&m->d_visited, N));
    CKM(cudaMalloc(&m->d_slots, sizeof(unsigned long long) * 4));
    CKM(cudaMalloc(&m->d_state, sizeof(MxState)));
    m->log_cap = std::max<long long>(4096, 8LL * n);
    CKM(cudaMalloc(&m->d_log, sizeof(MxMove) * m->log_cap));
    CKM(cudaMallocHost(&m->h_state, sizeof(MxState)));
    CKM(cudaMallocHost(&m->h_buf, sizeof(int) * N));
    CKM(cudaMemcpyAsync(m->d_C, costs, sizeof(double) * N * N, cudaMemcpyHostToDevice, m->stream));
    CKM(cudaStreamSynchronize(m->stream));
    m->stats.h2d_bytes += (int64_t)(sizeof(double) * N * N);
    *out = m;
    return TSPCUDA_OK;
}

// Replace the matrix of an existing context (same n): what a caller that is handed a fresh matrix per call -- the CPLEX
// relaxation callback through h_Greedy_2opt_mod_costs -- needs instead of a create / destroy pair per call.
extern "C" int tspcuda_mx_update(tspcuda_mx *m, const double *costs) {
    if (!m || !costs) return fail(TSPCUDA_INVALID_ARGUMENT, "null argument");
    CK(cudaSetDevice(m->device));
    const size_t N = (size_t)m->n;
    CK(cudaMemcpyAsync(m->d_C, costs, sizeof(double) * N * N, cudaMemcpyHostToDevice, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    m->stats.h2d_bytes += (int64_t)(sizeof(double) * N * N);
    return TSPCUDA_OK;
}

extern "C" int tspcuda_mx_get_stats(tspcuda_mx *m, tspcuda_stats *out) {
    if (!m || !out) return fail(TSPCUDA_INVALID_ARGUMENT, "null argument");
    *out = m->stats;
    return TSPCUDA_OK;
}

static int mx_use(tspcuda_mx *m) {
    if (!m) return fail(TSPCUDA_INVALID_ARGUMENT, "null context");
    CK(cudaSetDevice(m->device));
    return TSPCUDA_OK;
}

// h_greedyutil(start, solution, costs), heuristics.c:216-288.  *len_out (may be NULL) = nodes reached
// (< n when NOT_CONNECTED entries close the path early, as in the reference).
extern "C" int tspcuda_mx_nn(tspcuda_mx *m, int start, int *path_out, double *cost_out, int *len_out) {
    int e = mx_use(m);
    if (e) return e;
    if (start < 0 || start >= m->n) return fail(TSPCUDA_UNAVAILABLE, "starting node not correct");     // :223-226
    const int n = m->n;
    nn_matrix_kernel<<<1, MX_NN_BLOCK, 0, m->stream>>>(m->d_C, n, start, 1, m->d_visited, m->d_path, m->d_order, m->d_nncost, m->d_len);
    m->stats.kernel_launches += 1;
    CK(cudaGetLastError());
    double c = 0;
    int len = 0;
    CK(cudaMemcpyAsync(&c, m->d_nncost, sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    CK(cudaMemcpyAsync(&len, m->d_len, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    if (path_out) CK(cudaMemcpyAsync(m->h_buf, m->d_path, sizeof(int) * n, cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    m->stats.d2h_bytes += 12 + (path_out ? (int64_t)sizeof(int) * n : 0);
    if (path_out) std::memcpy(path_out, m->h_buf, sizeof(int) * n);
    if (cost_out) *cost_out = c;
    if (len_out) *len_out = len;
    return TSPCUDA_OK;
}

static int mx_descend(tspcuda_mx *m, long long max_moves, bool recompute, double time_limit_s, bool *deadline) {
    const int n = m->n;
    if (deadline) *deadline = false;
    MxTour T;
    T.tid = m->d_tid; T.pos = m->d_pos; T.dnp = m->d_dnp; T.n = n;
    const int grid = std::max(1, std::min(m->coop_blocks, (n + 7) / 8));
    const auto t0 = std::chrono::steady_clock::now();
    long long done = 0;
    bool first = true;
    for (;;) {
        if (max_moves >= 0 && done >= max_moves && !(first && recompute)) break;
        long long chunk = n < 4 ? 0 : 2048;
        if (max_moves >= 0) chunk = std::min(chunk, max_moves - done);
        int rc = (first && recompute) ? 1 : 0;
        CK(cudaMemsetAsync(m->d_slots, 0xFF, sizeof(unsigned long long) * 4, m->stream));
        CK(cudaEventRecord(m->ev0, m->stream));
        const double *C = m->d_C;
        unsigned long long *slots = m->d_slots;
        MxState *st = m->d_state;
        MxMove *log = m->d_log;
        long long log_cap = m->log_cap;
        void *args[] = {&C, &T, &slots, &st, &log, &log_cap, &chunk, &rc};
        CK(cudaLaunchCooperativeKernel((void *)descend_matrix_kernel, dim3(grid), dim3(MX_BLOCK), args, 0, m->stream));
        CK(cudaEventRecord(m->ev1, m->stream));
        m->stats.kernel_launches += 1;
        CK(cudaMemcpyAsync(m->h_state, m->d_state, sizeof(MxState), cudaMemcpyDeviceToHost, m->stream));
        CK(cudaStreamSynchronize(m->stream));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, m->ev0, m->ev1));
        m->stats.scan_kernel_ms += ms;
        m->stats.last_kernel_ms = ms;
        first = false;
        done = m->h_state->moves;
        if (m->h_state->status == 1 || n < 4 || chunk == 0) break;
        if (time_limit_s >= 0) {
            const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (el > time_limit_s) { if (deadline) *deadline = true; break; }
        }
    }
    return TSPCUDA_OK;
}

static int mx_upload_tour(tspcuda_mx *m, const int *path) {
    const int n = m->n;
    int *order = m->h_buf;
    int cur = 0;
    std::vector<unsigned char> seen(n, 0);
    for (int k = 0; k < n; k++) {
        if (cur < 0 || cur >= n || seen[cur]) return fail(TSPCUDA_INVALID_ARGUMENT, "path is not a single cycle through all nodes");
        seen[cur] = 1;
        order[k] = cur;
        cur = path[cur];
    }
    if (cur != 0) return fail(TSPCUDA_INVALID_ARGUMENT, "path is not a single cycle through all nodes");
    CK(cudaMemcpyAsync(m->d_order, order, sizeof(int) * n, cudaMemcpyHostToDevice, m->stream));
    m->stats.h2d_bytes += (int64_t)sizeof(int) * n;
    MxTour T;
    T.tid = m->d_tid; T.pos = m->d_pos; T.dnp = m->d_dnp; T.n = n;
    CK(cudaMemsetAsync(m->d_state, 0, sizeof(MxState), m->stream));
    mx_tour_init_kernel<<<std::max(1, std::min(4 * m->num_sms, (n + 256) / 256)), 256, 0, m->stream>>>(T, m->d_order);
    m->stats.kernel_launches += 1;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(m->stream));
    return TSPCUDA_OK;
}

__global__ void mx_export_kernel(MxTour T, int *path) {
    const int n = T.n;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) path[T.tid[p]] = T.tid[p + 1];
}

static int mx_download(tspcuda_mx *m, int *path) {
    const int n = m->n;
    MxTour T;
    T.tid = m->d_tid; T.pos = m->d_pos; T.dnp = m->d_dnp; T.n = n;
    mx_export_kernel<<<std::max(1, std::min(4 * m->num_sms, (n + 255) / 256)), 256, 0, m->stream>>>(T, m->d_path);
    m->stats.kernel_launches += 1;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(m->h_buf, m->d_path, sizeof(int) * n, cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    m->stats.d2h_bytes += (int64_t)sizeof(int) * n;
    std::memcpy(path, m->h_buf, sizeof(int) * n);
    return TSPCUDA_OK;
}

// ref_2opt(solution, costs, false), refinment.c:3-37, on the uploaded matrix.
extern "C" int tspcuda_mx_two_opt(tspcuda_mx *m, int *path, double *cost_out, int64_t max_moves, double time_limit_s,
                                  int64_t *moves_out, tspcuda_move_f *log, int64_t log_cap) {
    int e = mx_use(m);
    if (e) return e;
    if (!path) return fail(TSPCUDA_INVALID_ARGUMENT, "null path");
    e = mx_upload_tour(m, path);
    if (e) return e;
    bool deadline = false;
    e = mx_descend(m, max_moves, true, time_limit_s, &deadline);
    if (e) return e;
    m->stats.moves += m->h_state->moves;
    m->stats.scans += m->h_state->moves + (m->h_state->status == 1 ? 1 : 0);
    m->stats.pair_evals += (m->h_state->moves + (m->h_state->status == 1 ? 1 : 0)) * ((long long)m->n * (m->n - 1) / 2);
    if (moves_out) *moves_out = m->h_state->moves;
    if (cost_out) *cost_out = m->h_state->cost;
    if (log && log_cap > 0) {
        static_assert(sizeof(tspcuda_move_f) == sizeof(MxMove), "move record layout");
        const long long cnt = std::min<long long>(std::min<long long>(m->h_state->log_count, m->log_cap), log_cap);
        if (cnt > 0) {
            CK(cudaMemcpyAsync(log, m->d_log, sizeof(MxMove) * cnt, cudaMemcpyDeviceToHost, m->stream));
            CK(cudaStreamSynchronize(m->stream));
        }
    }
    e = mx_download(m, path);
    if (e) return e;
    return deadline ? TSPCUDA_DEADLINE_EXCEEDED : TSPCUDA_OK;
}

// ref_2opt_once(solution, costs), refinment.c:39-93: one scan; the move is applied iff best_delta < -1e-7.
extern "C" int tspcuda_mx_two_opt_once(tspcuda_mx *m, int *path, double *cost_inout, double *best_delta_out) {
    int e = mx_use(m);
    if (e) return e;
    if (!path || !cost_inout) return fail(TSPCUDA_INVALID_ARGUMENT, "null argument");
    e = mx_upload_tour(m, path);
    if (e) return e;
    e = mx_descend(m, 1, false, -1.0, nullptr);
    if (e) return e;
    const bool moved = m->h_state->moves > 0;
    // best_delta: the applied move's delta, or the (non-applicable) minimum found, or 0 when no pair improves
    const double d = moved ? m->h_state->last_delta : (m->h_state->status == 1 ? m->h_state->last_delta : 0.0);
    if (best_delta_out) *best_delta_out = d;
    m->stats.scans += 1;
    m->stats.moves += moved ? 1 : 0;
    if (moved) {
        *cost_inout += d;                                // refinment.c:85
        e = mx_download(m, path);
        if (e) return e;
    }
    return TSPCUDA_OK;
}

// ===== multi-GPU multi-start inside the library (one worker thread + context per GPU, NCCL incumbent reduction) =====
#include "mgpu.cuh"
