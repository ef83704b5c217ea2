// l2_microbench.cu -- achievable L2 rates for the access patterns of the SGNS kernel, measured on
// the box ("the builder must first measure achievable L2 random-32 B gather,
// scatter and red.global.add.v4.f32 rates").  Prints one JSON object.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/l2_microbench tools/l2_microbench.cu
//
// Patterns (table = n rows of 32 B, L2-resident; every thread draws rows with an LCG):
//   gather16x2   2 x ld.global.cg.v4.f32 per row     (what sgns_train_kernel issues today)
//   gather32     1 x ld.global.cg.v8.f32 per row     (256-bit load, sm_100)
//   red16x2      2 x red.global.add.v4.f32 per row, uniform rows
//   red_zipf     same, rows drawn from a Zipf(1.0)-like distribution (hub contention)
//   red_single   every thread hits row 0
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t lcg(uint32_t &s) { s = s * 1664525u + 1013904223u; return s; }

__device__ __forceinline__ void red_add_v4(float *p, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int MODE, int UNROLL>
__global__ void __launch_bounds__(256) bench_kernel(float *table, uint32_t n, int iters, const uint32_t *zipf, uint32_t zipf_n, float *sink)
{
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        uint32_t rows[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const uint32_t r = lcg(s) >> 8;
            if (MODE == 3) rows[u] = zipf[r % zipf_n];
            else if (MODE == 4 || MODE >= 10) rows[u] = 0;
            else rows[u] = r % n;
        }
        if (MODE == 0) {
            float4 v[UNROLL][2];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const float4 *p = reinterpret_cast<const float4 *>(table + (size_t)rows[u] * 8);
                v[u][0] = __ldcg(p);
                v[u][1] = __ldcg(p + 1);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) acc += v[u][0].x + v[u][1].w;
        } else if (MODE == 1) {
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                float a, b, c, d, e, f, g, h;
                asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=f"(a), "=f"(b), "=f"(c), "=f"(d), "=f"(e), "=f"(f), "=f"(g), "=f"(h)
                             : "l"(table + (size_t)rows[u] * 8));
                acc += a + h;
            }
        } else if (MODE >= 5) {
            // sub-warp cooperative access: G adjacent lanes share one row (one 32 B sector)
            constexpr int G = (MODE == 5 || MODE == 6 || MODE == 10) ? 4 : (MODE == 7 || MODE == 8 || MODE == 12) ? 8 : 2;
            const int sub = threadIdx.x & (G - 1);
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const uint32_t row = __shfl_sync(0xffffffffu, rows[u], threadIdx.x & 31 & ~(G - 1));
                float *p = table + (size_t)row * 8 + sub * (8 / G);
                if (MODE == 5) { const float2 v = __ldcg(reinterpret_cast<const float2 *>(p)); acc += v.x + v.y; }
                else if (MODE == 7) { acc += __ldcg(p); }
                else if (MODE == 6 || MODE == 10) asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(1e-9f), "f"(1e-9f) : "memory");
                else if (MODE == 8 || MODE == 12) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(1e-9f) : "memory");
                else red_add_v4(p, 1e-9f, 1e-9f, 1e-9f, 1e-9f);
            }
        } else {
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                float *p = table + (size_t)rows[u] * 8;
                red_add_v4(p, 1e-9f, 1e-9f, 1e-9f, 1e-9f);
                red_add_v4(p + 4, 1e-9f, 1e-9f, 1e-9f, 1e-9f);
            }
        }
    }
    if (acc == 123.456f) sink[0] = acc;
}

template <int MODE>
static double run(float *table, uint32_t n, int blocks, int iters, const uint32_t *zipf, uint32_t zn, float *sink)
{
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    bench_kernel<MODE, 6><<<blocks, 256>>>(table, n, 10, zipf, zn, sink);   // warm-up
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    bench_kernel<MODE, 6><<<blocks, 256>>>(table, n, iters, zipf, zn, sink);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    const int G = (MODE == 5 || MODE == 6 || MODE == 10) ? 4 : (MODE == 7 || MODE == 8 || MODE == 12) ? 8 : (MODE == 9 || MODE == 11) ? 2 : 1;
    const double rows = (double)blocks * 256 * iters * 6 / G;
    return rows / (ms * 1e-3);   // rows per second
}

int main(int argc, char **argv)
{
    const uint32_t n = argc > 1 ? atoi(argv[1]) : 38000;
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    float *table, *sink;
    CK(cudaMalloc(&table, (size_t)n * 32));
    CK(cudaMemset(table, 0, (size_t)n * 32));
    CK(cudaMalloc(&sink, 4));
    // Zipf(1.0) sample of rows: P(row k) ~ 1/(k+1)
    const uint32_t zn = 1 << 20;
    uint32_t *hz = (uint32_t *)malloc(zn * 4), *dz;
    double H = 0;
    for (uint32_t k = 0; k < n; ++k) H += 1.0 / (k + 1);
    uint32_t pos = 0;
    for (uint32_t k = 0; k < n && pos < zn; ++k) {
        uint32_t c = (uint32_t)(zn / H / (k + 1) + 0.5);
        for (uint32_t j = 0; j < c && pos < zn; ++j) hz[pos++] = k;
    }
    while (pos < zn) { hz[pos] = pos % n; ++pos; }
    CK(cudaMalloc(&dz, zn * 4));
    CK(cudaMemcpy(dz, hz, zn * 4, cudaMemcpyHostToDevice));
    printf("{\"n_rows\": %u, \"sms\": %d", n, sms);
    const int occ[] = {1, 2, 4, 8};
    for (int oi = 0; oi < 4; ++oi) {
        const int blocks = sms * occ[oi];
        const int iters = 4000;
        double g0 = run<0>(table, n, blocks, iters, dz, zn, sink);
        double g1 = run<1>(table, n, blocks, iters, dz, zn, sink);
        double r2 = run<2>(table, n, blocks, iters, dz, zn, sink);
        double r3 = run<3>(table, n, blocks, iters, dz, zn, sink);
        double r4 = run<4>(table, n, blocks, iters / 8, dz, zn, sink);
        double g5 = run<5>(table, n, blocks, iters, dz, zn, sink);
        double r6 = run<6>(table, n, blocks, iters, dz, zn, sink);
        double g7 = run<7>(table, n, blocks, iters, dz, zn, sink);
        double r8 = run<8>(table, n, blocks, iters, dz, zn, sink);
        double r9 = run<9>(table, n, blocks, iters, dz, zn, sink);
        double r10 = run<10>(table, n, blocks, iters / 8, dz, zn, sink);
        double r11 = run<11>(table, n, blocks, iters / 8, dz, zn, sink);
        double r12 = run<12>(table, n, blocks, iters / 8, dz, zn, sink);
        printf(", \"ctas_per_sm_%d\": {\"gather16x2_rows_per_s\": %.4g, \"gather16x2_GBps\": %.1f, "
               "\"gather32_rows_per_s\": %.4g, \"gather32_GBps\": %.1f, \"red16x2_rows_per_s\": %.4g, "
               "\"red16x2_GBps\": %.1f, \"red_zipf_rows_per_s\": %.4g, \"red_single_rows_per_s\": %.4g, "
               "\"gather_4lanes_x8B_rows_per_s\": %.4g, \"red_4lanes_v2_rows_per_s\": %.4g, "
               "\"gather_8lanes_x4B_rows_per_s\": %.4g, \"red_8lanes_f32_rows_per_s\": %.4g, "
               "\"red_2lanes_v4_rows_per_s\": %.4g, \"red_single_4lanes_v2_rows_per_s\": %.4g, "
               "\"red_single_2lanes_v4_rows_per_s\": %.4g, \"red_single_8lanes_f32_rows_per_s\": %.4g}",
               occ[oi], g0, g0 * 32 / 1e9, g1, g1 * 32 / 1e9, r2, r2 * 32 / 1e9, r3, r4, g5, r6, g7, r8, r9, r10, r11, r12);
    }
    printf("}\n");
    return 0;
}
