// l2_mixbench.cu -- achievable rate of the SGNS step's memory pattern in isolation: groups of 4
// lanes; per "step" a group gathers NL random 32-byte rows (8 B per lane), then issues NR row
// reductions (red.v2 per lane) whose operands depend on the gathered values.  Rows are drawn
// uniformly or with a power-law popularity like the noise distribution.  Prints JSON lines.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/l2_mixbench tools/l2_mixbench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <string.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t lcg(uint32_t &s) { s = s * 1664525u + 1013904223u; return s; }

template <int NL, int NR>
__global__ void __launch_bounds__(256) mix_kernel(float *table, const uint32_t *pick, uint32_t npick, int iters, float *sink)
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

// variant: reductions to the H most popular rows are accumulated in a CTA-local shared buffer
// (shared atomics) and flushed round-robin, one row per group every F steps (atomicExch + red)
template <int NL, int NR, int H, int F>
__global__ void __launch_bounds__(256) mix_hub_kernel(float *table, const uint32_t *pick, uint32_t npick, int iters, float *sink)
{
    __shared__ float acc_s[H * 8];
    for (int i = threadIdx.x; i < H * 8; i += blockDim.x) acc_s[i] = 0.f;
    __syncthreads();
    const int sub = threadIdx.x & 3;
    const unsigned gmask = 0xFu << (threadIdx.x & 28);
    const int group = threadIdx.x >> 2;
    uint32_t s = ((blockIdx.x * blockDim.x + threadIdx.x) >> 2) * 2654435761u + 12345u;
    float acc = 0.f;
    int flush_row = group % H;
    for (int it = 0; it < iters; ++it) {
        uint32_t rows[NL];
        float2 v[NL];
#pragma unroll
        for (int u = 0; u < NL; ++u) rows[u] = pick[(lcg(s) >> 8) % npick];
#pragma unroll
        for (int u = 0; u < NL; ++u) {
            v[u] = __ldcg(reinterpret_cast<const float2 *>(table + (size_t)rows[u] * 8 + sub * 2));
            if (rows[u] < H) { v[u].x += acc_s[rows[u] * 8 + sub * 2]; v[u].y += acc_s[rows[u] * 8 + sub * 2 + 1]; }
        }
        float f = 0.f;
#pragma unroll
        for (int u = 0; u < NL; ++u) f += v[u].x * v[0].y;
        f += __shfl_xor_sync(gmask, f, 1);
        f += __shfl_xor_sync(gmask, f, 2);
        acc += f;
        const float g = f * 1e-30f;
#pragma unroll
        for (int u = 0; u < NR; ++u) {
            const uint32_t r = rows[u % NL];
            if (r < H) {
                atomicAdd(&acc_s[r * 8 + sub * 2], g);
                atomicAdd(&acc_s[r * 8 + sub * 2 + 1], g);
            } else {
                float *p = table + (size_t)r * 8 + sub * 2;
                asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(g), "f"(g) : "memory");
            }
        }
        if (it % F == 0) {
            const float a = atomicExch(&acc_s[flush_row * 8 + sub * 2], 0.f);
            const float b = atomicExch(&acc_s[flush_row * 8 + sub * 2 + 1], 0.f);
            if (a != 0.f || b != 0.f) {
                float *p = table + (size_t)flush_row * 8 + sub * 2;
                asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
            }
            flush_row = (flush_row + 64) % H;
        }
    }
    if (acc == 123.456f) sink[0] = acc;
}

template <int NL, int NR, int H, int F>
static void run_hub(const char *name, float *table, const uint32_t *pick, uint32_t npick, int blocks, int threads, float *sink)
{
    const int iters = 3000;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    mix_hub_kernel<NL, NR, H, F><<<blocks, threads>>>(table, pick, npick, 20, sink);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    mix_hub_kernel<NL, NR, H, F><<<blocks, threads>>>(table, pick, npick, iters, sink);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    const double steps = (double)blocks * threads / 4 * iters;
    printf("{\"pattern\": \"%s\", \"gathers\": %d, \"reds\": %d, \"hub_rows\": %d, \"flush_every\": %d, \"groups\": %d, \"steps_per_s\": %.4g, \"us_per_step\": %.3f}\n",
           name, NL, NR, H, F, blocks * threads / 4, steps / (ms * 1e-3), ms * 1e3 / iters);
}

template <int NL, int NR>
static void run(const char *name, float *table, const uint32_t *pick, uint32_t npick, int blocks, int threads, float *sink)
{
    const int iters = 3000;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    mix_kernel<NL, NR><<<blocks, threads>>>(table, pick, npick, 20, sink);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    mix_kernel<NL, NR><<<blocks, threads>>>(table, pick, npick, iters, sink);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    const double steps = (double)blocks * threads / 4 * iters;
    printf("{\"pattern\": \"%s\", \"gathers\": %d, \"reds\": %d, \"groups\": %d, \"steps_per_s\": %.4g, \"us_per_step\": %.3f}\n",
           name, NL, NR, blocks * threads / 4, steps / (ms * 1e-3), ms * 1e3 / iters);
}

int main(int argc, char **argv)
{
    // --peak ROWS: only the uniform-row patterns over a table of ROWS rows (bench.py's live roofline
    // denominator: the L2 sector-request rate of the step's gather + reduction mix without any
    // same-row collisions to speak of)
    const bool peak_only = argc > 2 && strcmp(argv[1], "--peak") == 0;
    const uint32_t n = peak_only ? (uint32_t)strtoul(argv[2], nullptr, 10) : 38000;
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
    if (peak_only) {
        for (int per_sm = 128; per_sm <= 512; per_sm *= 2) {
            run<13, 14>("uniform", table, d_uni, np, sms * per_sm / 64, 256, sink);
            run<13, 0>("uniform", table, d_uni, np, sms * per_sm / 64, 256, sink);
        }
        return 0;
    }
    if (argc > 1) {
        // row popularity of a real workload: a binary file of uint32 row ids drawn by the caller
        // (tools/make_pick.py: centre/context rows ~ degree, negatives ~ degree^0.75 of the C2 network)
        FILE *fp = fopen(argv[1], "rb");
        if (!fp) { printf("cannot open %s\n", argv[1]); return 1; }
        size_t got = fread(h, 4, np, fp);
        fclose(fp);
        if (got != np) { printf("short read %zu\n", got); return 1; }
        uint32_t *d_real;
        CK(cudaMalloc(&d_real, np * 4)); CK(cudaMemcpy(d_real, h, np * 4, cudaMemcpyHostToDevice));
        for (int per_sm = 64; per_sm <= 256; per_sm *= 2)
            run<13, 14>("c2_popularity", table, d_real, np, sms * per_sm / 64, 256, sink);
    }
    for (int per_sm = 64; per_sm <= 512; per_sm *= 2) {
        const int threads = 256, blocks = sms * per_sm / 64;
        run<13, 0>("uniform", table, d_uni, np, blocks, threads, sink);
        run<13, 7>("uniform", table, d_uni, np, blocks, threads, sink);
        run<13, 14>("uniform", table, d_uni, np, blocks, threads, sink);
        run<13, 14>("powerlaw", table, d_pl, np, blocks, threads, sink);
        if (per_sm == 64) {
            run_hub<13, 14, 128, 1>("powerlaw", table, d_pl, np, blocks, threads, sink);
            run_hub<13, 14, 128, 4>("powerlaw", table, d_pl, np, blocks, threads, sink);
            run_hub<13, 14, 128, 16>("powerlaw", table, d_pl, np, blocks, threads, sink);
            run_hub<13, 14, 512, 4>("powerlaw", table, d_pl, np, blocks, threads, sink);
            run_hub<13, 14, 1024, 16>("powerlaw", table, d_pl, np, blocks, threads, sink);
        }
    }
    return 0;
}
