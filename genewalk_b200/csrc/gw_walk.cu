// gw_walk.cu -- uniform first-order random walks (replaces deepwalk.py:50-96,145-200).
//
// One thread per walk.  Per step the thread reads row_ptr[v], row_ptr[v+1] (same 32 B sector 7
// times out of 8) and one col_idx entry -- dependent, L2-resident random gathers (CSR is 8-40 MB at
// the BASELINE configs, L2 is 126 MB) -- and writes one token.  Tokens are position-major
// (tokens[i*W + p]) so the 32 lanes of a warp write 128 contiguous bytes per position: the only HBM
// stream of the kernel is fully coalesced.  Algorithmic bytes per step: 8 (row offsets) + 4
// (neighbour) + 4 (token) = 16 B.
#include "gw_common.cuh"

namespace gw {

__global__ void expand_rows_kernel(int32_t n, const int32_t *__restrict__ row_ptr,
                                   int32_t *__restrict__ row_of_entry)
{
    // one warp per row: row_of_entry[row_ptr[v] .. row_ptr[v+1]) = v
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t v = warp; v < n; v += nwarps) {
        const int32_t b = row_ptr[v], e = row_ptr[v + 1];
        for (int32_t k = b + lane; k < e; k += 32) row_of_entry[k] = (int32_t)v;
    }
}

constexpr int kWalkThreads = 256;

__global__ void __launch_bounds__(kWalkThreads)
walk_uniform_kernel(const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col_idx,
                    const int32_t *__restrict__ row_of_entry, int64_t A, int32_t niter, int32_t L,
                    uint32_t key0, uint32_t key1, uint64_t stride, int64_t slot_begin,
                    int64_t nslots, int32_t *__restrict__ tokens)
{
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < nslots; o += nthreads) {
        const int64_t p = slot_begin + o;
        int64_t e;
        uint64_t w;
        if (stride == 0) {  // canonical (reference corpus) order: slot == walk id
            w = (uint64_t)p;
            e = p / niter;
        } else {
            const int64_t it = p / A, r = p - it * A;
            // (r * stride) % A with r, stride < A < 2^31: the product fits 64 bits
            e = (int64_t)(((uint64_t)r * stride) % (uint64_t)A);
            w = (uint64_t)e * (uint64_t)niter + (uint64_t)it;
        }
        int32_t v = __ldg(row_of_entry + e);
        __stcs(tokens + o, v);
        u32x4 rnd = {0, 0, 0, 0};
        for (int32_t s = 1; s < L; ++s) {
            const int sel = (s - 1) & 3;
            if (sel == 0)
                rnd = philox4x32_10((uint32_t)w, (uint32_t)(w >> 32), (uint32_t)((s - 1) >> 2),
                                    kStreamWalk, key0, key1);
            const uint32_t word = sel == 0 ? rnd.x : sel == 1 ? rnd.y : sel == 2 ? rnd.z : rnd.w;
            const int32_t b = __ldg(row_ptr + v);
            const uint32_t deg = (uint32_t)(__ldg(row_ptr + v + 1) - b);
            v = __ldg(col_idx + b + (int32_t)__umulhi(word, deg));
            __stcs(tokens + (int64_t)s * nslots + o, v);  // streaming store: never re-read here
        }
    }
}


// ---- walks counted, not stored ---------------------------------------------------------------
// The trainer regenerates every walk from its Philox counter (gw_sgns_train with tokens == NULL),
// so the corpus never exists in memory; the vocabulary (gensim build_vocab's raw counts) comes
// from one counting pass over the same walks.  n <= 49152: every CTA keeps a private histogram in
// shared memory (192 KB) and flushes its non-zero bins once; larger graphs: one warp-aggregated
// global atomic per distinct node of a warp's 32 walks at each position.
constexpr int kWalkCountThreads = 1024;
constexpr int kWalkCountSmemBins = 49152;

__device__ __forceinline__ int32_t walk_step(const int32_t *__restrict__ row_ptr,
                                             const int32_t *__restrict__ col_idx, int32_t v, uint64_t w,
                                             int32_t s, u32x4 &rnd, uint32_t key0, uint32_t key1)
{
    const int sel = (s - 1) & 3;
    if (sel == 0)
        rnd = philox4x32_10((uint32_t)w, (uint32_t)(w >> 32), (uint32_t)((s - 1) >> 2), kStreamWalk, key0, key1);
    const uint32_t word = sel == 0 ? rnd.x : sel == 1 ? rnd.y : sel == 2 ? rnd.z : rnd.w;
    const int32_t b = __ldg(row_ptr + v);
    const uint32_t deg = (uint32_t)(__ldg(row_ptr + v + 1) - b);
    return __ldg(col_idx + b + (int32_t)__umulhi(word, deg));
}

__global__ void __launch_bounds__(kWalkCountThreads)
walk_count_smem_kernel(const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col_idx,
                       const int32_t *__restrict__ row_of_entry, int32_t n, int32_t niter, int32_t L,
                       uint32_t key0, uint32_t key1, int64_t W, unsigned long long *__restrict__ counts)
{
    extern __shared__ uint32_t bins[];
    for (int i = threadIdx.x; i < n; i += blockDim.x) bins[i] = 0;
    __syncthreads();
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < W; w += nthreads) {
        int32_t v = __ldg(row_of_entry + w / niter);
        atomicAdd(&bins[v], 1u);
        u32x4 rnd = {0, 0, 0, 0};
        for (int32_t s = 1; s < L; ++s) {
            v = walk_step(row_ptr, col_idx, v, (uint64_t)w, s, rnd, key0, key1);
            atomicAdd(&bins[v], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t c = bins[i];
        if (c) atomicAdd(counts + i, (unsigned long long)c);
    }
}

__global__ void __launch_bounds__(kWalkThreads)
walk_count_global_kernel(const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col_idx,
                         const int32_t *__restrict__ row_of_entry, int32_t niter, int32_t L,
                         uint32_t key0, uint32_t key1, int64_t W, unsigned long long *__restrict__ counts)
{
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    for (int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) - lane; w0 < W; w0 += nthreads) {
        const int64_t w = w0 + lane;
        const bool valid = w < W;
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
        if (!valid) continue;
        int32_t v = __ldg(row_of_entry + w / niter);
        u32x4 rnd = {0, 0, 0, 0};
        for (int32_t s = 0; s < L; ++s) {
            if (s > 0) v = walk_step(row_ptr, col_idx, v, (uint64_t)w, s, rnd, key0, key1);
            const unsigned peers = __match_any_sync(vmask, v);
            if ((peers & ((1u << lane) - 1u)) == 0) atomicAdd(counts + v, (unsigned long long)__popc(peers));
        }
    }
}

}  // namespace gw

extern "C" int gw_walk_uniform(int32_t n, int64_t nnz, const int32_t *row_ptr,
                               const int32_t *col_idx, int32_t niter, int32_t L, uint64_t seed, uint64_t stride,
                               int64_t slot_begin, int64_t slot_end, int32_t *tokens,
                               gw_stream_t stream)
{
    using namespace gw;
    cudaStream_t s = (cudaStream_t)stream;
    GW_REQUIRE(n > 0 && row_ptr && col_idx && tokens, "gw_walk_uniform: null/empty graph");
    GW_REQUIRE(niter > 0 && L > 0 && L <= 4096, "gw_walk_uniform: niter=%d L=%d out of range", niter, L);
    const int64_t A = nnz;
    GW_REQUIRE(A > 0 && A < ((int64_t)1 << 31), "gw_walk_uniform: nnz=%lld out of range", (long long)A);
    const int64_t W = A * (int64_t)niter;
    GW_REQUIRE(slot_begin >= 0 && slot_begin <= slot_end && slot_end <= W,
               "gw_walk_uniform: slots [%lld,%lld) outside [0,%lld)", (long long)slot_begin,
               (long long)slot_end, (long long)W);
    GW_REQUIRE(stride < (uint64_t)A + (A == 1), "gw_walk_uniform: stride must be in [0, A)");
    if (stride != 0) {
        uint64_t a = (uint64_t)A, b = stride;
        while (b) { const uint64_t t = a % b; a = b; b = t; }
        GW_REQUIRE(a == 1, "gw_walk_uniform: stride %llu is not coprime to A=%lld",
                   (unsigned long long)stride, (long long)A);
    }
    const int64_t nslots = slot_end - slot_begin;
    if (nslots == 0) return GW_OK;
    Temp rows;
    GW_CUDA_OK(rows.alloc(sizeof(int32_t) * (size_t)A, s));
    const int sms = num_sms();
    expand_rows_kernel<<<sms * 8, 256, 0, s>>>(n, row_ptr, rows.as<int32_t>());
    GW_LAUNCH_OK();
    // grid: a multiple of the SM count, 8 resident CTAs of 256 threads per SM (2048 threads/SM)
    int64_t blocks = ceil_div<int64_t>(nslots, kWalkThreads);
    const int64_t cap = (int64_t)sms * 8 * 4;
    if (blocks > cap) blocks = cap;
    walk_uniform_kernel<<<(unsigned)blocks, kWalkThreads, 0, s>>>(
        row_ptr, col_idx, rows.as<int32_t>(), A, niter, L, (uint32_t)seed, (uint32_t)(seed >> 32),
        stride, slot_begin, nslots, tokens);
    GW_LAUNCH_OK();
    return GW_OK;
}

extern "C" int gw_walk_count(int32_t n, int64_t nnz, const int32_t *row_ptr, const int32_t *col_idx,
                             int32_t niter, int32_t L, uint64_t seed, int64_t *counts,
                             gw_stream_t stream)
{
    using namespace gw;
    cudaStream_t s = (cudaStream_t)stream;
    GW_REQUIRE(n > 0 && row_ptr && col_idx && counts, "gw_walk_count: null/empty graph");
    GW_REQUIRE(niter > 0 && L > 0 && L <= 4096, "gw_walk_count: niter=%d L=%d out of range", niter, L);
    const int64_t A = nnz;
    GW_REQUIRE(A >= 0 && A < ((int64_t)1 << 31), "gw_walk_count: nnz=%lld out of range", (long long)A);
    GW_CUDA_OK(cudaMemsetAsync(counts, 0, sizeof(int64_t) * (size_t)n, s));
    const int64_t W = A * (int64_t)niter;
    if (W == 0) return GW_OK;
    Temp rows;
    GW_CUDA_OK(rows.alloc(sizeof(int32_t) * (size_t)A, s));
    const int sms = num_sms();
    expand_rows_kernel<<<sms * 8, 256, 0, s>>>(n, row_ptr, rows.as<int32_t>());
    GW_LAUNCH_OK();
    auto *c = reinterpret_cast<unsigned long long *>(counts);
    if (n <= kWalkCountSmemBins) {
        const size_t smem = sizeof(uint32_t) * (size_t)n;
        GW_CUDA_OK(cudaFuncSetAttribute(walk_count_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
        // every CTA of 1024 walkers owns a histogram: one per SM, two when the bins leave room
        int64_t blocks = (int64_t)sms * (smem <= 100 * 1024 ? 2 : 1);
        const int64_t need = ceil_div<int64_t>(W, kWalkCountThreads);
        if (blocks > need) blocks = need;
        walk_count_smem_kernel<<<(unsigned)blocks, kWalkCountThreads, smem, s>>>(
            row_ptr, col_idx, rows.as<int32_t>(), n, niter, L, (uint32_t)seed, (uint32_t)(seed >> 32), W, c);
    } else {
        int64_t blocks = ceil_div<int64_t>(W, kWalkThreads);
        const int64_t cap = (int64_t)sms * 8;
        if (blocks > cap) blocks = cap;
        walk_count_global_kernel<<<(unsigned)blocks, kWalkThreads, 0, s>>>(
            row_ptr, col_idx, rows.as<int32_t>(), niter, L, (uint32_t)seed, (uint32_t)(seed >> 32), W, c);
    }
    GW_LAUNCH_OK();
    return GW_OK;
}
