// mgpu.cuh -- multi-start work sharded over the GPUs of one box from INSIDE the library (host code, included
// at the end of tspcuda.cu): what the drop-in CLI uses when TSPCUDA_GPUS > 1.  bench.py / multistart.py do the
// same with one process per GPU over torch.distributed; the C driver is one process, so here it is one worker
// thread and one tspcuda_ctx per GPU, and NCCL communicators created with ncclCommInitAll.
//
// Sharding: starts are independent units (heuristics.c:42,82 loop over them); GPU g of G takes
// starts first + (g + k G) stride (block-cyclic: descent lengths vary by start) and there is no data-path
// collective.  The only exchange is the incumbent: ONE ncclAllReduce(ncclMin, ncclUint64) of
// key = cost << 20 | start node -- integer order of the key = (cost, start) lexicographic = the reference's
// strict-'<' incumbent rule evaluated in ascending start order (tsp.c:671, heuristics.c:58/82): the LOWEST
// start wins cost ties -- followed by ONE ncclBroadcast of the winner's successor array from its owner.
//
// NCCL is resolved at run time (dlopen "libnccl.so.2"): libtspcuda.so has no link-time dependency on it, a
// single-GPU user never loads it, and inside a torch process the already-loaded NCCL is reused.  No fallback:
// without NCCL the multi-GPU entry points fail with TSPCUDA_UNAVAILABLE.
#pragma once
#include <dlfcn.h>
#include <nccl.h>
#include <pthread.h>
#include <unistd.h>

#include <functional>
#include <thread>

namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi *nccl_api(std::string &why) {
    static NcclApi api;
    static bool tried = false;
    static std::string err;
    if (!tried) {
        tried = true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *nm : names)
            if ((api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL))) break;
        if (!api.handle) {
            err = std::string("NCCL not found: ") + dlerror();
        } else {
            api.CommInitAll = (decltype(api.CommInitAll))dlsym(api.handle, "ncclCommInitAll");
            api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
            api.AllReduce = (decltype(api.AllReduce))dlsym(api.handle, "ncclAllReduce");
            api.Broadcast = (decltype(api.Broadcast))dlsym(api.handle, "ncclBroadcast");
            api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
            if (!api.CommInitAll || !api.CommDestroy || !api.AllReduce || !api.Broadcast || !api.GetErrorString) {
                err = "NCCL library lacks a required entry point";
                api.handle = nullptr;
            }
        }
    }
    why = err;
    return api.handle ? &api : nullptr;
}

constexpr int MGPU_START_BITS = 20;
constexpr unsigned long long MGPU_NO_RESULT = ~0ull;

}  // namespace

struct tspcuda_mgpu {
    int n = 0, ndev = 0;
    std::vector<int> devices;
    std::vector<tspcuda_ctx *> ctx;
    std::vector<ncclComm_t> comm;
    std::vector<unsigned long long *> d_key;
    std::vector<int *> d_path;
    double last_reduce_ms = 0, last_compute_ms = 0;
};

extern "C" void tspcuda_mgpu_destroy(tspcuda_mgpu *m) {
    if (!m) return;
    std::string why;
    NcclApi *N = nccl_api(why);
    for (int g = 0; g < (int)m->ctx.size(); ++g) {
        if (g < (int)m->devices.size()) cudaSetDevice(m->devices[g]);
        if (g < (int)m->d_key.size()) cudaFree(m->d_key[g]);
        if (g < (int)m->d_path.size()) cudaFree(m->d_path[g]);
        if (N && g < (int)m->comm.size() && m->comm[g]) N->CommDestroy(m->comm[g]);
        tspcuda_destroy(m->ctx[g]);
    }
    delete m;
}

extern "C" int tspcuda_mgpu_create(tspcuda_mgpu **out, const double *xy, int n, int weight_type, int ndev, const int *devices) {
    if (!out || !xy) return fail(TSPCUDA_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    const int have = tspcuda_device_count();
    if (ndev <= 0) ndev = have;
    if (ndev <= 0 || ndev > have) return fail(TSPCUDA_UNAVAILABLE, "%d GPUs requested, %d visible (no CPU fallback)", ndev, have);
    tspcuda_mgpu *m = new tspcuda_mgpu;
    m->n = n;
    m->ndev = ndev;
    for (int g = 0; g < ndev; ++g) m->devices.push_back(devices ? devices[g] : g);
    m->ctx.assign(ndev, nullptr);
    m->comm.assign(ndev, nullptr);
    m->d_key.assign(ndev, nullptr);
    m->d_path.assign(ndev, nullptr);
    for (int g = 0; g < ndev; ++g) {
        int e = tspcuda_create(&m->ctx[g], xy, n, weight_type, m->devices[g]);
        if (e) { tspcuda_mgpu_destroy(m); return e; }
        if (cudaMalloc(&m->d_key[g], sizeof(unsigned long long)) != cudaSuccess ||
            cudaMalloc(&m->d_path[g], sizeof(int) * (size_t)n) != cudaSuccess) {
            tspcuda_mgpu_destroy(m);
            return fail(TSPCUDA_RESOURCE_EXHAUSTED, "device %d: cudaMalloc failed", m->devices[g]);
        }
    }
    if (ndev > 1) {
        std::string why;
        NcclApi *N = nccl_api(why);
        if (!N) { tspcuda_mgpu_destroy(m); return fail(TSPCUDA_UNAVAILABLE, "%s", why.c_str()); }
        // NCCL announces itself on STDOUT when NCCL_DEBUG asks for it ("NCCL version ..."); the reference's quiet mode
        // promises a stdout that holds nothing but the result line (errors.c:159-164), so stdout points at stderr
        // while the communicators are created.
        fflush(stdout);
        const int saved = dup(STDOUT_FILENO);
        if (saved >= 0) dup2(STDERR_FILENO, STDOUT_FILENO);
        ncclResult_t r = N->CommInitAll(m->comm.data(), ndev, m->devices.data());
        if (saved >= 0) { fflush(stdout); dup2(saved, STDOUT_FILENO); close(saved); }
        if (r != ncclSuccess) {
            const std::string msg = N->GetErrorString(r);
            for (auto &cm : m->comm) cm = nullptr;
            tspcuda_mgpu_destroy(m);
            return fail(TSPCUDA_INTERNAL, "ncclCommInitAll: %s", msg.c_str());
        }
    }
    *out = m;
    return TSPCUDA_OK;
}

extern "C" int tspcuda_mgpu_ndev(const tspcuda_mgpu *m) { return m ? m->ndev : 0; }

// One shard's work: starts f, f + st, ... (count of them) on context c; its best under the incumbent rule in
// (*bs, path, *bc), *bs < 0 when it got to none; *taken = starts it got to.  Returns an ERROR_CODE.
typedef std::function<int(tspcuda_ctx *c, int f, int count, int st, int *bs, int *path, double *bc, int *taken)> ShardJob;

static int mgpu_run(tspcuda_mgpu *m, const ShardJob &job, int first_start, int nstarts, int stride,
                    int *best_start_out, int *best_path_out, double *best_cost_out, int *starts_done_out,
                    double *compute_ms_out, double *reduce_ms_out) {
    if (!m) return fail(TSPCUDA_INVALID_ARGUMENT, "null handle");
    const int n = m->n, G = m->ndev;
    if (nstarts <= 0 || stride <= 0) return fail(TSPCUDA_INVALID_ARGUMENT, "nstarts/stride");
    const long long last = (long long)first_start + (long long)(nstarts - 1) * stride;
    if (first_start < 0 || last >= n) return fail(TSPCUDA_UNAVAILABLE, "starting node not correct");
    if (n > (1 << MGPU_START_BITS)) return fail(TSPCUDA_OUT_OF_RANGE, "n beyond the %d-bit start field of the reduction key", MGPU_START_BITS);
    std::string why;
    NcclApi *N = G > 1 ? nccl_api(why) : nullptr;
    if (G > 1 && !N) return fail(TSPCUDA_UNAVAILABLE, "%s", why.c_str());

    if (starts_done_out) *starts_done_out = 0;
    std::vector<int> code(G, TSPCUDA_OK), taken(G, 0);
    std::vector<char> late(G, 0);
    std::vector<std::string> msg(G);
    std::vector<unsigned long long> key(G, MGPU_NO_RESULT), winner(G, MGPU_NO_RESULT);
    std::vector<double> t_compute(G, 0.0), t_reduce(G, 0.0);
    std::vector<int> h_path((size_t)n);
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, nullptr, (unsigned)G);

    auto worker = [&](int g) {
        tspcuda_ctx *c = m->ctx[g];
        const int count = nstarts > g ? (nstarts - g + G - 1) / G : 0;       // starts g, g + G, ... of the caller's sequence
        std::vector<int> path((size_t)n);
        auto fail_here = [&](int e, const char *what) { code[g] = e; msg[g] = what; };
        cudaSetDevice(m->devices[g]);
        const auto t0 = std::chrono::steady_clock::now();
        if (count > 0) {
            int bs = -1;
            double bc = 0;
            const int f = first_start + g * stride, st = stride * G;
            // every GPU runs against the same limit and keeps the best of the prefix of ITS shard it got to (the union of
            // those prefixes is what a time-limited run of the reference's ascending loop corresponds to, heuristics.c:83-90)
            int e = job(c, f, count, st, &bs, path.data(), &bc, &taken[g]);
            if (e == TSPCUDA_DEADLINE_EXCEEDED) { late[g] = 1; e = TSPCUDA_OK; }
            if (e) fail_here(e, tspcuda_last_error());
            else if (bs < 0) key[g] = MGPU_NO_RESULT;                            // the limit passed before this shard's first start
            else if (bc < 0 || bc >= (double)(1ll << 42)) fail_here(TSPCUDA_OUT_OF_RANGE, "tour cost beyond the reduction key");
            else key[g] = ((unsigned long long)(long long)bc << MGPU_START_BITS) | (unsigned long long)bs;
        }
        t_compute[g] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        pthread_barrier_wait(&bar);                                          // every shard is done (or has failed)
        bool any_failed = false;
        for (int k = 0; k < G; ++k) any_failed = any_failed || code[k] != TSPCUDA_OK;
        if (any_failed) return;                                              // nobody enters a collective
        const auto t1 = std::chrono::steady_clock::now();
        unsigned long long w = key[g];
        if (G > 1) {
            cudaStream_t s = c->stream;
            bool ok = cudaMemcpyAsync(m->d_key[g], &key[g], sizeof(w), cudaMemcpyHostToDevice, s) == cudaSuccess;
            ok = ok && N->AllReduce(m->d_key[g], m->d_key[g], 1, ncclUint64, ncclMin, m->comm[g], s) == ncclSuccess;
            ok = ok && cudaMemcpyAsync(&w, m->d_key[g], sizeof(w), cudaMemcpyDeviceToHost, s) == cudaSuccess;
            ok = ok && cudaStreamSynchronize(s) == cudaSuccess;
            // the winner's successor array travels from the GPU that owns its start
            const int wstart = (int)(w & ((1ull << MGPU_START_BITS) - 1));
            const int owner = ((wstart - first_start) / stride) % G;
            if (ok && w != MGPU_NO_RESULT) {
                if (g == owner) ok = cudaMemcpyAsync(m->d_path[g], path.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, s) == cudaSuccess;
                ok = ok && N->Broadcast(m->d_path[g], m->d_path[g], (size_t)n, ncclInt32, owner, m->comm[g], s) == ncclSuccess;
                if (g == 0) ok = ok && cudaMemcpyAsync(h_path.data(), m->d_path[g], sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, s) == cudaSuccess;
                ok = ok && cudaStreamSynchronize(s) == cudaSuccess;
            }
            if (!ok) fail_here(TSPCUDA_INTERNAL, "NCCL best-tour reduction failed");
        } else {
            h_path = path;
        }
        winner[g] = w;
        t_reduce[g] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count();
    };
    std::vector<std::thread> th;
    for (int g = 1; g < G; ++g) th.emplace_back(worker, g);
    worker(0);
    for (auto &t : th) t.join();
    pthread_barrier_destroy(&bar);
    for (int g = 0; g < G; ++g)
        if (code[g] != TSPCUDA_OK) return fail(code[g], "GPU %d: %s", m->devices[g], msg[g].c_str());
    for (int g = 1; g < G; ++g)
        if (winner[g] != winner[0]) return fail(TSPCUDA_INTERNAL, "all-reduce(min) disagrees between GPUs");
    m->last_compute_ms = *std::max_element(t_compute.begin(), t_compute.end());
    m->last_reduce_ms = *std::max_element(t_reduce.begin(), t_reduce.end());
    if (compute_ms_out) *compute_ms_out = m->last_compute_ms;
    if (reduce_ms_out) *reduce_ms_out = m->last_reduce_ms;
    bool any_late = false;
    int total_taken = 0;
    for (int g = 0; g < G; ++g) { any_late = any_late || late[g]; total_taken += taken[g]; }
    if (starts_done_out) *starts_done_out = total_taken;
    if (winner[0] == MGPU_NO_RESULT) {
        if (any_late) { if (best_start_out) *best_start_out = -1; return TSPCUDA_DEADLINE_EXCEEDED; }
        return fail(TSPCUDA_INTERNAL, "no shard produced a tour");
    }
    if (best_start_out) *best_start_out = (int)(winner[0] & ((1ull << MGPU_START_BITS) - 1));
    if (best_cost_out) *best_cost_out = (double)(long long)(winner[0] >> MGPU_START_BITS);
    if (best_path_out) std::memcpy(best_path_out, h_path.data(), sizeof(int) * (size_t)n);
    return any_late ? TSPCUDA_DEADLINE_EXCEEDED : TSPCUDA_OK;
}

extern "C" int tspcuda_mgpu_multistart(tspcuda_mgpu *m, int two_opt, int first_start, int nstarts, int stride, double time_limit_s,
                                       int *best_start_out, int *best_path_out, double *best_cost_out, int *starts_done_out,
                                       double *compute_ms_out, double *reduce_ms_out) {
    const ShardJob job = [=](tspcuda_ctx *c, int f, int count, int st, int *bs, int *path, double *bc, int *taken) {
        return two_opt ? tspcuda_multistart_two_opt(c, f, count, st, time_limit_s, nullptr, nullptr, nullptr, bs, path, bc, taken)
                       : tspcuda_nn_multi(c, f, count, st, time_limit_s, nullptr, bs, path, bc, taken);
    };
    return mgpu_run(m, job, first_start, nstarts, stride, best_start_out, best_path_out, best_cost_out, starts_done_out, compute_ms_out,
                    reduce_ms_out);
}

// Config 5: tabu / VNS restarts sharded over the GPUs (tspcuda_restarts on every shard, same reduction as the multi-start).
extern "C" int tspcuda_mgpu_restarts(tspcuda_mgpu *m, int kind, int first_start, int nstarts, int stride, int iters, unsigned seed,
                                     double time_limit_s, int *best_start_out, int *best_path_out, double *best_cost_out,
                                     int *restarts_done_out, double *compute_ms_out, double *reduce_ms_out) {
    const ShardJob job = [=](tspcuda_ctx *c, int f, int count, int st, int *bs, int *path, double *bc, int *taken) {
        return tspcuda_restarts(c, kind, f, count, st, iters, seed, time_limit_s, bs, path, bc, taken);
    };
    return mgpu_run(m, job, first_start, nstarts, stride, best_start_out, best_path_out, best_cost_out, restarts_done_out, compute_ms_out,
                    reduce_ms_out);
}
