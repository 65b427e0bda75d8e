// k1_bench.cu -- stand-alone timing + self-check of K1 (cost_matrix_kernel) build variants (seconds to compile instead of the
// minutes of the whole library).  nvcc -DTSP_K1_ROWS=.. -DTSP_K1_MINB=.. -DTSP_K1_BLOCK=.. ; args: n integer(0/1) trips ctas_per_sm
// Check: every cell against the scalar cost functor M::K (diagonal -1.0) on the device.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../travellingsalesmanoptimization_b200/csrc/kernels.cuh"
using namespace tspcuda;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

template <class M>
__global__ void check_kernel(const double2 *xy, int n, const double *m, unsigned long long *bad) {
    const size_t cells = (size_t)n * n, stride = (size_t)gridDim.x * blockDim.x;
    unsigned long long local = 0;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < cells; t += stride) {
        const int i = (int)(t / n), j = (int)(t - (size_t)i * n);
        const double want = (i == j) ? -1.0 : (double)(M::K(xy[i].x, xy[i].y, xy[j].x, xy[j].y) - KOFF);
        local += (m[t] != want);
    }
    if (local) atomicAdd(bad, local);
}

template <class M>
static void run(const char *name, int n, bool integer, int trips, int ctas) {
    std::vector<double2> h(n);
    unsigned long long s = 88172645463325252ull;
    for (int i = 0; i < n; ++i) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17; double x = (double)(s % 2000000) ;
        s ^= s << 13; s ^= s >> 7; s ^= s << 17; double y = (double)(s % 1000000);
        if (!integer) { x += (double)(s % 997) / 997.0; y += (double)(s % 991) / 1024.0; }
        if (i % 1000 == 999) { x = h[i - 1].x; y = h[i - 1].y; }       // coincident nodes
        h[i] = make_double2(x, y);
    }
    double2 *d_xy; double *d_m; unsigned long long *d_bad;
    CK(cudaMalloc(&d_xy, sizeof(double2) * n));
    CK(cudaMalloc(&d_m, sizeof(double) * (size_t)n * n));
    CK(cudaMalloc(&d_bad, 8));
    CK(cudaMemcpy(d_xy, h.data(), sizeof(double2) * n, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_bad, 0, 8));
    const int chunk_pairs = trips > 0 ? trips * TSP_K1_BLOCK : (int)((n / 2 + TSP_K1_BLOCK) / TSP_K1_BLOCK) * TSP_K1_BLOCK;
    const size_t nunits = ((size_t)(n + TSP_K1_ROWS - 1) / TSP_K1_ROWS + 1) * std::max<size_t>(1, (n / 2 + chunk_pairs - 1) / chunk_pairs);
    const int grid = (int)std::max<size_t>(1, std::min<size_t>(nunits, ctas > 0 ? (size_t)148 * ctas : (size_t)0x7fffffff));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f, sum = 0;
    const int reps = 7;
    for (int r = 0; r < reps + 2; ++r) {
        CK(cudaEventRecord(e0));
        cost_matrix_kernel<M><<<grid, TSP_K1_BLOCK>>>(d_xy, n, d_m, chunk_pairs);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r >= 2) { best = std::min(best, ms); sum += ms; }
    }
    CK(cudaGetLastError());
    check_kernel<M><<<148 * 8, 256>>>(d_xy, n, d_m, d_bad);
    unsigned long long bad; CK(cudaMemcpy(&bad, d_bad, 8, cudaMemcpyDeviceToHost));
    // pure writer over the same buffer
    CK(cudaEventRecord(e0)); CK(cudaMemsetAsync(d_m, 0, sizeof(double) * (size_t)n * n)); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float fill; CK(cudaEventElapsedTime(&fill, e0, e1));
    const double gb = 8.0 * n * n / 1e9;
    printf("%-10s n=%6d %s rows=%d block=%d minb=%d rot=%d trips=%d ctas=%d  avg %.3f ms best %.3f ms  %.0f GB/s (best %.0f) = %.3f of 6551   memset %.0f GB/s  mismatches %llu\n",
           name, n, integer ? "int " : "frac", TSP_K1_ROWS, TSP_K1_BLOCK, TSP_K1_MINB, TSP_K1_ROT, trips, ctas, sum / reps, best, gb / (sum / reps) * 1e3, gb / best * 1e3,
           gb / (sum / reps) * 1e3 / 6551.0, gb / fill * 1e3, bad);
    cudaFree(d_xy); cudaFree(d_m); cudaFree(d_bad);
}

int main(int argc, char **argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 85900;
    const int integer = argc > 2 ? atoi(argv[2]) : 1;
    const int trips = argc > 3 ? atoi(argv[3]) : 4;
    const int ctas = argc > 4 ? atoi(argv[4]) : 0;          // 0 = one unit per CTA (the shipped launch shape)
    if (integer) run<EucInt>("EucInt", n, true, trips, ctas);
    else run<EucGeneral>("EucGeneral", n, false, trips, ctas);
    if (argc > 5) { run<Att>("Att", n, true, trips, ctas); run<Ceil2D>("Ceil2D", n, true, trips, ctas); }
    return 0;
}
