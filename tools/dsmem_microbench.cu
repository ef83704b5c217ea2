// dsmem_microbench.cu -- would a CLUSTER-resident trainer beat the L2-resident one?  
//
// One replicate's two embedding tables (2 x 38000 rows x 32 B = 2.43 MB at C2) fit the distributed
// shared memory of a 16-CTA thread-block cluster (16 x 152 KB).  This benchmark runs the SGNS step's
// access mix -- per step a group of 4 lanes gathers 13 random 32-byte rows (8 B per lane) and then
// issues 14 row reductions whose operands depend on the gathered values -- against
//   (a) a table in global memory / L2  (ld.global.cg + red.global.add.v2.f32; what the shipped kernel does)
//   (b) the same table spread over the shared memory of a 16-CTA cluster (rows interleaved over
//       the CTAs; mapa + ld.shared::cluster.v2.f32 / red.shared::cluster.add.f32 -- the latter has no
//       vector form and, like every fp32 add on shared memory, is lowered by ptxas to a
//       compare-and-swap loop (SASS: ATOMS.CAST.SPIN), i.e. a round trip, not a fire-and-forget)
// with rows drawn uniformly or with a power-law popularity like the noise distribution, and prints
// one JSON line per case.  9 clusters of 16 CTAs cover 144 of the 148 SMs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/dsmem_microbench tools/dsmem_microbench.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t lcg(uint32_t &s) { s = s * 1664525u + 1013904223u; return s; }

constexpr int NL = 13, NR = 14, CL = 16;

__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ float2 ld_cluster_v2(uint32_t addr)
{
    float2 v;
    asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void red_cluster(uint32_t addr, float x)
{
    asm volatile("red.shared::cluster.add.f32 [%0], %1;" ::"r"(addr), "f"(x) : "memory");
}

__global__ void __launch_bounds__(512) l2_kernel(float *table, const uint32_t *pick, uint32_t npick, int iters, float *sink)
{
    const int sub = threadIdx.x & 3;
    const unsigned gmask = 0xFu << (threadIdx.x & 28);
    uint32_t s = ((blockIdx.x * blockDim.x + threadIdx.x) >> 2) * 2654435761u + 12345u;
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        uint32_t rows[NL];
        float2 v[NL];
#pragma unroll
        for (int u = 0; u < NL; ++u) rows[u] = pick[(lcg(s) >> 8) % npick];
#pragma unroll
        for (int u = 0; u < NL; ++u) v[u] = __ldcg(reinterpret_cast<const float2 *>(table + (size_t)rows[u] * 8 + sub * 2));
        float f = 0.f;
#pragma unroll
        for (int u = 0; u < NL; ++u) f += v[u].x * v[0].y;
        f += __shfl_xor_sync(gmask, f, 1);
        f += __shfl_xor_sync(gmask, f, 2);
        acc += f;
        const float g = f * 1e-30f;
#pragma unroll
        for (int u = 0; u < NR; ++u) {
            float *p = table + (size_t)rows[u % NL] * 8 + sub * 2;
            asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(g), "f"(g) : "memory");
        }
    }
    if (acc == 123.456f) sink[0] = acc;
}

// table rows interleaved over the CL CTAs of the cluster: row r lives in CTA r % CL at local row r / CL
__global__ void __launch_bounds__(512) dsmem_kernel(uint32_t nrows, const uint32_t *pick, uint32_t npick, int iters, float *sink)
{
    extern __shared__ __align__(16) float slice[];
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t local_rows = (nrows + CL - 1) / CL;
    for (uint32_t i = threadIdx.x; i < local_rows * 8; i += blockDim.x) slice[i] = 0.f;
    cluster.sync();
    const int sub = threadIdx.x & 3;
    const unsigned gmask = 0xFu << (threadIdx.x & 28);
    const uint32_t base0 = (uint32_t)__cvta_generic_to_shared(slice);
    uint32_t s = ((blockIdx.x * blockDim.x + threadIdx.x) >> 2) * 2654435761u + 12345u;
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        uint32_t rows[NL];
        float2 v[NL];
#pragma unroll
        for (int u = 0; u < NL; ++u) rows[u] = pick[(lcg(s) >> 8) % npick];
#pragma unroll
        for (int u = 0; u < NL; ++u)
            v[u] = ld_cluster_v2(mapa(base0 + (rows[u] / CL) * 32 + sub * 8, rows[u] % CL));
        float f = 0.f;
#pragma unroll
        for (int u = 0; u < NL; ++u) f += v[u].x * v[0].y;
        f += __shfl_xor_sync(gmask, f, 1);
        f += __shfl_xor_sync(gmask, f, 2);
        acc += f;
        const float g = f * 1e-30f;
#pragma unroll
        for (int u = 0; u < NR; ++u) {
            const uint32_t a = mapa(base0 + (rows[u % NL] / CL) * 32 + sub * 8, rows[u % NL] % CL);
            red_cluster(a, g);
            red_cluster(a + 4, g);
        }
    }
    cluster.sync();   // no CTA may exit while others still address its shared memory
    if (acc == 123.456f) sink[0] = acc;
}

static double time_l2(float *table, const uint32_t *pick, uint32_t np, int blocks, int threads, float *sink, int iters)
{
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    l2_kernel<<<blocks, threads>>>(table, pick, np, 20, sink);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    l2_kernel<<<blocks, threads>>>(table, pick, np, iters, sink);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    return (double)blocks * threads / 4 * iters / (ms * 1e-3);
}

static double time_dsmem(uint32_t nrows, const uint32_t *pick, uint32_t np, int clusters, int threads, float *sink, int iters)
{
    const size_t smem = (size_t)((nrows + CL - 1) / CL) * 32;
    CK(cudaFuncSetAttribute(dsmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(dsmem_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(clusters * CL);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    int warm = 20;
    CK(cudaLaunchKernelEx(&cfg, dsmem_kernel, nrows, pick, np, warm, sink));
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    CK(cudaLaunchKernelEx(&cfg, dsmem_kernel, nrows, pick, np, iters, sink));
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    return (double)clusters * CL * threads / 4 * iters / (ms * 1e-3);
}

int main()
{
    const uint32_t n = 2 * 38000;   // both tables of one C2 replicate
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    float *table, *sink;
    CK(cudaMalloc(&table, (size_t)n * 32));
    CK(cudaMemset(table, 0, (size_t)n * 32));
    CK(cudaMalloc(&sink, 4));
    const uint32_t np = 1 << 22;
    uint32_t *h = (uint32_t *)malloc(np * 4), *d_uni, *d_pl;
    for (uint32_t i = 0; i < np; ++i) h[i] = (uint32_t)(((uint64_t)i * 2654435761u) % n);
    CK(cudaMalloc(&d_uni, np * 4)); CK(cudaMemcpy(d_uni, h, np * 4, cudaMemcpyHostToDevice));
    // power-law popularity p(k) ~ (k+10)^(-0.5)  (degree ~ k^-2/3, noise ~ degree^0.75)
    double Z = 0; for (uint32_t k = 0; k < n; ++k) Z += pow(k + 10.0, -0.5);
    uint32_t pos = 0;
    for (uint32_t k = 0; k < n && pos < np; ++k) { uint32_t c = (uint32_t)(np * pow(k + 10.0, -0.5) / Z + 0.5); for (uint32_t j = 0; j < c && pos < np; ++j) h[pos++] = k; }
    while (pos < np) { h[pos] = pos % n; ++pos; }
    for (uint32_t i = np - 1; i > 0; --i) { uint32_t j = (uint32_t)(((uint64_t)i * 2654435761u + 12345u) % (i + 1)); uint32_t t = h[i]; h[i] = h[j]; h[j] = t; }
    CK(cudaMalloc(&d_pl, np * 4)); CK(cudaMemcpy(d_pl, h, np * 4, cudaMemcpyHostToDevice));
    const int clusters = sms / CL;   // 9 on B200
    const char *names[2] = {"uniform", "powerlaw"};
    const uint32_t *picks[2] = {d_uni, d_pl};
    for (int p = 0; p < 2; ++p) {
        for (int threads = 128; threads <= 512; threads *= 2) {
            // (a) ONE table pair in L2, the same number of workers as the cluster variant
            const double l2 = time_l2(table, picks[p], np, clusters * CL, threads, sink, 2000);
            printf("{\"where\": \"l2\", \"pattern\": \"%s\", \"groups\": %d, \"steps_per_s\": %.4g}\n", names[p],
                   clusters * CL * threads / 4, l2);
            // (b) 9 clusters, each with its OWN table pair in distributed shared memory
            const double ds = time_dsmem(n, picks[p], np, clusters, threads, sink, 2000);
            printf("{\"where\": \"dsmem\", \"pattern\": \"%s\", \"clusters\": %d, \"groups\": %d, \"steps_per_s\": %.4g, "
                   "\"steps_per_s_per_cluster\": %.4g, \"vs_l2\": %.3f}\n", names[p], clusters,
                   clusters * CL * threads / 4, ds, ds / clusters, ds / l2);
            fflush(stdout);
        }
    }
    return 0;
}
