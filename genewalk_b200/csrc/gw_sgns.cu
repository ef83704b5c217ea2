// gw_sgns.cu -- vocabulary, noise (alias) table and weight init of the SGNS trainer; the training
// kernel itself is in gw_sgns_train.cu.
// Replaces the host-side steps of gensim.models.Word2Vec(sg=1, min_count=1, negative=K, sample=0)
// as reached from /root/reference/genewalk/deepwalk.py:135-139 (gensim is third-party and absent
// from the reference tree): build_vocab's counts and descending-count order, the count^0.75 noise
// distribution (gensim: make_cum_table + bisect; here an alias table) and init_weights.
// All of it is a few streaming passes over the token array and O(n) work on the vocabulary:
// 0.05 % of a replicate.
#include <math.h>

#include "gw_common.cuh"

namespace gw {

// =============================================================================================
// token counts
// =============================================================================================
constexpr int kCountThreads = 1024;
constexpr int kCountSmemBins = 49152;  // 192 KB of uint32 bins: covers n <= 49152 (C1-C3)

// privatised histogram: one CTA per SM keeps all bins in shared memory, flushes non-zero bins once
__global__ void __launch_bounds__(kCountThreads)
count_tokens_smem_kernel(const int32_t *__restrict__ tokens, int64_t count, int32_t n,
                         unsigned long long *__restrict__ counts)
{
    extern __shared__ uint32_t bins[];
    for (int i = threadIdx.x; i < n; i += blockDim.x) bins[i] = 0;
    __syncthreads();
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nvec = count >> 2;
    const int4 *t4 = reinterpret_cast<const int4 *>(tokens);
    for (int64_t i = tid; i < nvec; i += nthreads) {
        const int4 t = __ldcs(t4 + i);
        if (t.x >= 0) atomicAdd(&bins[t.x], 1u);
        if (t.y >= 0) atomicAdd(&bins[t.y], 1u);
        if (t.z >= 0) atomicAdd(&bins[t.z], 1u);
        if (t.w >= 0) atomicAdd(&bins[t.w], 1u);
    }
    for (int64_t i = (nvec << 2) + tid; i < count; i += nthreads)
        if (tokens[i] >= 0) atomicAdd(&bins[tokens[i]], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t c = bins[i];
        if (c) atomicAdd(counts + i, (unsigned long long)c);
    }
}

__global__ void __launch_bounds__(256)
count_tokens_global_kernel(const int32_t *__restrict__ tokens, int64_t count,
                           unsigned long long *__restrict__ counts)
{
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    for (int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) - lane; i0 < count;
         i0 += nthreads) {
        const int64_t i = i0 + lane;
        const bool valid = i < count;
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            const int32_t t = __ldcs(tokens + i);
            // warp-aggregated: one atomic per distinct token in the warp
            const unsigned peers = __match_any_sync(vmask, t);
            if (t >= 0 && (peers & ((1u << lane) - 1u)) == 0)
                atomicAdd(counts + t, (unsigned long long)__popc(peers));
        }
    }
}

// =============================================================================================
// vocabulary rank (descending count, ties by node id)
// =============================================================================================
constexpr uint64_t kCountCap = ((uint64_t)1 << 40) - 1;

__global__ void vocab_keys_kernel(int32_t n, const int64_t *__restrict__ counts, uint64_t *keys,
                                  uint32_t *vals)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        uint64_t c = (uint64_t)counts[i];
        if (c > kCountCap) c = kCountCap;
        keys[i] = kCountCap - c;  // ascending key == descending count
        vals[i] = (uint32_t)i;
    }
}

__global__ void vocab_finish_kernel(int32_t n, const uint64_t *__restrict__ keys,
                                    const uint32_t *__restrict__ vals, int32_t *rank_of_node,
                                    int32_t *node_of_rank, int32_t *n_vocab)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) {
        const bool present = keys[r] != kCountCap;
        const int32_t node = (int32_t)vals[r];
        node_of_rank[r] = node;
        rank_of_node[node] = present ? r : -1;
        const bool next_present = (r + 1 < n) && keys[r + 1] != kCountCap;
        if (present && !next_present) *n_vocab = r + 1;
    }
}

// =============================================================================================
// alias table, built in parallel.
//
// Vose's method is a sequential sweep: "light" slots (q = w*n/sum < 1) are topped up by the current
// "heavy" item; a heavy item whose residual drops below 1 becomes a slot topped up by the next
// heavy item.  With lights l_0..l_{nl-1} and heavies h_0..h_{nh-1} in index order the residual of
// heavy b after a lights and b heavies have been filled is a closed form,
//     S(a, b) = LP[a] + HP[b],  LP[a] = sum_{k<a} q(l_k) - a  (decreasing in a),
//                               HP[b] = sum_{k<=b} q(h_k) - b  (non-decreasing in b),
// so every slot can find its partner with one binary search instead of replaying the sweep:
//     light a:  donor  j = min{ b : S(a, b) >= 1 }          -> prob = q(l_a),  alias = h_j
//     heavy b:  a* = min{ a in [0, nl] : S(a, b) < 1 }       -> prob = S(a*, b), alias = h_{b+1}
// (no such j / a*, or b+1 == nh: the slot keeps itself with prob 1 -- the last heavy item, up to
// rounding).  Sums are fp64; the implied distribution reproduces w/sum to ~1e-12 (tests).
// =============================================================================================
__device__ __forceinline__ double noise_weight(int64_t c, double power, int fast075)
{
    if (c <= 0) return 0.0;
    const double x = (double)c;
    if (fast075) return sqrt(sqrt(x * x * x));  // x^0.75 from correctly rounded ops only
    return pow(x, power);
}

__global__ void alias_weights_kernel(int32_t n, const int64_t *__restrict__ counts, double power,
                                     int fast075, double *w)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) w[i] = noise_weight(counts[i], power, fast075);
}

// q = w*n/total; light flags and the two masked q arrays for the prefix sums
__global__ void alias_split_kernel(int32_t n, const double *__restrict__ w,
                                   const double *__restrict__ total, double *q, int64_t *is_light,
                                   double *q_light, double *q_heavy)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const double qi = w[i] * ((double)n / *total);
        const bool light = qi < 1.0;
        q[i] = qi;
        is_light[i] = light ? 1 : 0;
        q_light[i] = light ? qi : 0.0;
        q_heavy[i] = light ? 0.0 : qi;
    }
}

// compaction: light a -> (id, LP[a]); heavy b -> (id, HP[b]); also LP[nl]
__global__ void alias_compact_kernel(int32_t n, const double *__restrict__ q,
                                     const int64_t *__restrict__ light_before,
                                     const int64_t *__restrict__ n_light,
                                     const double *__restrict__ lsum_before,
                                     const double *__restrict__ lsum_total,
                                     const double *__restrict__ hsum_before, int32_t *light_id,
                                     double *LP, int32_t *heavy_id, double *HP)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int64_t a = light_before[i];
        const int64_t b = (int64_t)i - a;
        if (q[i] < 1.0) {
            light_id[a] = i;
            LP[a] = lsum_before[i] - (double)a;
        } else {
            heavy_id[b] = i;
            HP[b] = (hsum_before[i] + q[i]) - (double)b;
        }
    }
    if (i == 0) LP[*n_light] = *lsum_total - (double)(*n_light);
}

__global__ void alias_assign_kernel(int32_t n, const double *__restrict__ q,
                                    const int64_t *__restrict__ light_before,
                                    const int64_t *__restrict__ n_light,
                                    const int32_t *__restrict__ heavy_id,
                                    const double *__restrict__ LP, const double *__restrict__ HP,
                                    gw_alias_entry *table)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t nl = *n_light, nh = (int64_t)n - nl;
    const int64_t a = light_before[i];
    const int64_t b = (int64_t)i - a;
    float prob = 1.0f;
    int32_t alias = i;
    if (q[i] < 1.0) {
        const double lp = LP[a];
        int64_t lo = 0, hi = nh;  // first b with lp + HP[b] >= 1
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (lp + HP[mid] >= 1.0) hi = mid; else lo = mid + 1;
        }
        if (lo < nh) {
            prob = (float)q[i];
            alias = heavy_id[lo];
        }
    } else {
        const double hp = HP[b];
        int64_t lo = 0, hi = nl + 1;  // first a in [0, nl] with LP[a] + hp < 1
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (LP[mid] + hp < 1.0) hi = mid; else lo = mid + 1;
        }
        if (lo <= nl && b + 1 < nh) {
            double r = LP[lo] + hp;
            r = r < 0.0 ? 0.0 : (r > 1.0 ? 1.0 : r);
            prob = (float)r;
            alias = heavy_id[b + 1];
        }
    }
    table[i].prob = prob;
    table[i].alias = alias;
}

// =============================================================================================
// init
// =============================================================================================
__global__ void sgns_init_kernel(int32_t n, int32_t d, const int32_t *__restrict__ rank_of_node,
                                 const float *__restrict__ rng_rows, float *syn0, float *syn1neg)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (int64_t)n * d) {
        const int32_t v = (int32_t)(i / d), k = (int32_t)(i - (int64_t)v * d);
        const int32_t r = rank_of_node[v];
        syn0[i] = r >= 0 ? rng_rows[(int64_t)r * d + k] : 0.0f;
        syn1neg[i] = 0.0f;
    }
}

}  // namespace gw

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" int gw_count_tokens(const int32_t *tokens, int64_t count, int32_t n, int64_t *counts,
                               gw_stream_t stream)
{
    using namespace gw;
    cudaStream_t s = (cudaStream_t)stream;
    GW_REQUIRE(n > 0 && counts && (tokens || count == 0) && count >= 0, "gw_count_tokens: bad arguments");
    GW_CUDA_OK(cudaMemsetAsync(counts, 0, sizeof(int64_t) * (size_t)n, s));
    if (count == 0) return GW_OK;
    const int sms = num_sms();
    if (n <= kCountSmemBins && (reinterpret_cast<uintptr_t>(tokens) & 15) == 0) {
        const size_t smem = sizeof(uint32_t) * (size_t)n;
        GW_CUDA_OK(cudaFuncSetAttribute(count_tokens_smem_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // each CTA's private bins are uint32: bound the tokens one CTA can see below 2^32
        int64_t blocks = sms;
        while (ceil_div<int64_t>(count, blocks) >= ((int64_t)1 << 32)) blocks *= 2;
        count_tokens_smem_kernel<<<(unsigned)blocks, kCountThreads, smem, s>>>(
            tokens, count, n, reinterpret_cast<unsigned long long *>(counts));
    } else {
        count_tokens_global_kernel<<<sms * 8, 256, 0, s>>>(
            tokens, count, reinterpret_cast<unsigned long long *>(counts));
    }
    GW_LAUNCH_OK();
    return GW_OK;
}

extern "C" int gw_vocab_rank(int32_t n, const int64_t *counts, int32_t *rank_of_node,
                             int32_t *node_of_rank, int32_t *n_vocab, gw_stream_t stream)
{
    using namespace gw;
    cudaStream_t s = (cudaStream_t)stream;
    GW_REQUIRE(n > 0 && counts && rank_of_node && node_of_rank && n_vocab, "gw_vocab_rank: bad arguments");
    Temp tk, tv;
    GW_CUDA_OK(tk.alloc(sizeof(uint64_t) * (size_t)n, s));
    GW_CUDA_OK(tv.alloc(sizeof(uint32_t) * (size_t)n, s));
    const int blocks = ceil_div(n, 256);
    vocab_keys_kernel<<<blocks, 256, 0, s>>>(n, counts, tk.as<uint64_t>(), tv.as<uint32_t>());
    GW_LAUNCH_OK();
    GW_TRY(radix_sort_u64(tk.as<uint64_t>(), tv.as<uint32_t>(), n, 40, s));
    GW_CUDA_OK(cudaMemsetAsync(n_vocab, 0, sizeof(int32_t), s));
    vocab_finish_kernel<<<blocks, 256, 0, s>>>(n, tk.as<uint64_t>(), tv.as<uint32_t>(), rank_of_node,
                                               node_of_rank, n_vocab);
    GW_LAUNCH_OK();
    return GW_OK;
}

extern "C" int gw_alias_build(int32_t n, const int64_t *counts, double power, gw_alias_entry *table,
                              gw_stream_t stream)
{
    using namespace gw;
    cudaStream_t s = (cudaStream_t)stream;
    GW_REQUIRE(n > 0 && counts && table, "gw_alias_build: bad arguments");
    GW_REQUIRE(power >= 0.0 && power <= 4.0, "gw_alias_build: power %g out of range", power);
    const size_t N = (size_t)n;
    // one slab: w, q, q_light, q_heavy, lsum, hsum (double[n]) | LP (n+1), HP (n) | is_light,
    // light_before (int64[n]) | light_id, heavy_id (int32[n]) | totals
    Temp slab;
    const size_t bytes = sizeof(double) * (8 * N + 1) + sizeof(int64_t) * 2 * N +
                         sizeof(int32_t) * 2 * N + 64;
    GW_CUDA_OK(slab.alloc(bytes, s));
    double *w = slab.as<double>();
    double *q = w + N, *ql = q + N, *qh = ql + N, *lsum = qh + N, *hsum = lsum + N;
    double *LP = hsum + N, *HP = LP + N + 1;
    int64_t *is_light = reinterpret_cast<int64_t *>(HP + N);
    int64_t *light_before = is_light + N;
    int32_t *light_id = reinterpret_cast<int32_t *>(light_before + N);
    int32_t *heavy_id = light_id + N;
    char *tail = reinterpret_cast<char *>(heavy_id + N);
    tail += (8 - (reinterpret_cast<uintptr_t>(tail) & 7)) & 7;
    double *w_total = reinterpret_cast<double *>(tail);
    double *l_total = w_total + 1, *h_total = w_total + 2;
    int64_t *n_light = reinterpret_cast<int64_t *>(w_total + 3);
    const int blocks = ceil_div(n, 256);
    alias_weights_kernel<<<blocks, 256, 0, s>>>(n, counts, power, power == 0.75 ? 1 : 0, w);
    GW_LAUNCH_OK();
    GW_TRY(scan_exclusive_f64(w, q /*scratch*/, n, w_total, s));
    alias_split_kernel<<<blocks, 256, 0, s>>>(n, w, w_total, q, is_light, ql, qh);
    GW_LAUNCH_OK();
    GW_TRY(scan_exclusive_i64(is_light, light_before, n, n_light, s));
    GW_TRY(scan_exclusive_f64(ql, lsum, n, l_total, s));
    GW_TRY(scan_exclusive_f64(qh, hsum, n, h_total, s));
    alias_compact_kernel<<<blocks, 256, 0, s>>>(n, q, light_before, n_light, lsum, l_total, hsum,
                                                light_id, LP, heavy_id, HP);
    GW_LAUNCH_OK();
    alias_assign_kernel<<<blocks, 256, 0, s>>>(n, q, light_before, n_light, heavy_id, LP, HP, table);
    GW_LAUNCH_OK();
    return GW_OK;
}

extern "C" int gw_sgns_init(int32_t n, int32_t d, const int32_t *rank_of_node, const float *rng_rows,
                            float *syn0, float *syn1neg, gw_stream_t stream)
{
    using namespace gw;
    cudaStream_t s = (cudaStream_t)stream;
    GW_REQUIRE(n > 0 && d > 0 && rank_of_node && rng_rows && syn0 && syn1neg, "gw_sgns_init: bad arguments");
    const int64_t total = (int64_t)n * d;
    sgns_init_kernel<<<(unsigned)ceil_div<int64_t>(total, 256), 256, 0, s>>>(n, d, rank_of_node,
                                                                            rng_rows, syn0, syn1neg);
    GW_LAUNCH_OK();
    return GW_OK;
}
