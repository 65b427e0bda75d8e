// Probe: cluster launch via attribute, DSMEM 64-bit atomics, visibility of GLOBAL writes across a cluster barrier.
#include <cooperative_groups.h>
#include <cstdio>
namespace cg = cooperative_groups;
template <int CL, bool FENCE>
__global__ void __launch_bounds__(1024, 1) probe(int *buf, int n, int iters, unsigned long long *out) {
    __shared__ unsigned long long s_min_, s_sum_;
    cg::cluster_group cl = cg::this_cluster();
    const unsigned rank = cl.block_rank();
    unsigned long long *s_min = cl.map_shared_rank(&s_min_, 0), *s_sum = cl.map_shared_rank(&s_sum_, 0);
    const int t = rank * 1024 + threadIdx.x, T = CL * 1024;
    int *mine = buf + (size_t)(blockIdx.x / CL) * n;
    unsigned long long bad = 0;
    if (t == 0) { *s_min = ~0ull; *s_sum = 0; }
    cl.sync();
    for (int it = 0; it < iters; ++it) {
        // every thread writes a strip, then reads the strip written by the "opposite" thread
        for (int p = t; p < n; p += T) mine[p] = it * 7 + p;
        if (FENCE) __threadfence();
        cl.sync();
        if (FENCE) __threadfence();
        for (int p = t; p < n; p += T) { const int q = n - 1 - p; if (mine[q] != it * 7 + q) ++bad; }
        atomicMin(s_min, (unsigned long long)(t + it));
        atomicAdd(s_sum, 1ull);
        if (FENCE) __threadfence();
        cl.sync();
        if (FENCE) __threadfence();
    }
    if (bad) atomicAdd(out + 0, bad);
    cl.sync();
    if (t == 0) { out[1 + 3 * (blockIdx.x / CL)] = *s_min; out[2 + 3 * (blockIdx.x / CL)] = *s_sum; out[3 + 3 * (blockIdx.x / CL)] = cl.num_blocks(); }
}
template <int CL, bool FENCE>
void run(int clusters) {
    int n = 13509, iters = 200;
    int *buf; unsigned long long *out, h[64] = {0};
    cudaMalloc(&buf, sizeof(int) * n * clusters); cudaMalloc(&out, sizeof(h)); cudaMemset(out, 0, sizeof(h));
    cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(clusters * CL); cfg.blockDim = dim3(1024);
    cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = CL; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
    cfg.attrs = a; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, probe<CL, FENCE>, buf, n, iters, out);
    cudaError_t e2 = cudaDeviceSynchronize();
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("CL=%d fence=%d launch=%s sync=%s stale_reads=%llu  cluster0: min=%llu sum=%llu (expect %llu) blocks=%llu\n", CL, (int)FENCE,
           cudaGetErrorString(e), cudaGetErrorString(e2), h[0], h[1], h[2], (unsigned long long)CL * 1024 * iters, h[3]);
    cudaFree(buf); cudaFree(out);
}
int main() { run<1, false>(4); run<2, false>(4); run<2, true>(4); run<4, false>(4); run<8, false>(4); return 0; }
