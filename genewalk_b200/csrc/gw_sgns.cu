// gw_sgns.cu -- skip-gram negative sampling on L2-resident embedding tables.
// Replaces gensim.models.Word2Vec(sg=1, window=1, negative=K, sample=0) as called at
// /root/reference/genewalk/deepwalk.py:135-139 (gensim itself is third-party and absent from the
// reference tree; the update rule below is its published w2v_fast_sentence_sg_neg).
//
// Layout / roofline.  syn0 and syn1neg are [n, d] fp32 row-major; at d = 8 a row is exactly one
// 32-byte sector, and both tables (2*n*32 B = 2.4 MB at n = 38k, 64 MB at n = 1M) live in the
// 126 MB L2 for the whole run, so the governing bound is L2 sector / reduction throughput, not HBM.
// One THREAD owns one walk (18 (centre, context) pairs at L = 10) and, per pair, keeps the whole
// context row, the K+1 target rows and the accumulated `work` row in registers: 2 + 2(K+1)
// ld.global.cg.v4 (L1-bypassing: other SMs update these rows) and as many red.global.add.v4.f32
// (one REDG.E.ADD.F32x4 per 16 bytes, fire-and-forget, no lost updates under Hogwild).
// Algorithmic bytes per pair: 4d(2K+4) table + 8 token + 8K alias = 496 B at d=8, K=5.
// The only HBM stream is the position-major token array (coalesced, read once per epoch).
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

// =============================================================================================
// training
// =============================================================================================
constexpr int kExpTableSize = 1000;
constexpr float kMaxExp = 6.0f;
constexpr int kTrainThreads = 256;

struct TrainParams {
    const int32_t *tokens;
    int64_t W;
    int32_t L;
    int32_t n;
    int32_t K;
    int32_t epochs;
    int32_t epoch;
    double alpha0;
    double alpha_span;  // alpha0 - alpha_min
    double alpha_min;
    int64_t job_walks;
    const gw_alias_entry *alias;
    const float *exp_table;  // [1000] device copy of gensim's EXP_TABLE
    float *syn0;
    float *syn1neg;
    uint32_t key0, key1;
};

__device__ __forceinline__ void red_add_v4(float *p, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c),
                 "f"(d)
                 : "memory");
}

template <int D> __device__ __forceinline__ void load_row(const float *p, float (&r)[D])
{
#pragma unroll
    for (int k = 0; k < D; k += 4) {
        const float4 v = __ldcg(reinterpret_cast<const float4 *>(p + k));  // L2, never stale L1
        r[k] = v.x; r[k + 1] = v.y; r[k + 2] = v.z; r[k + 3] = v.w;
    }
}

// sequential-order, unfused dot product: identical rounding to the CPU oracle used by the tests
template <int D> __device__ __forceinline__ float dot_row(const float (&a)[D], const float (&b)[D])
{
    float f = 0.0f;
#pragma unroll
    for (int k = 0; k < D; ++k) f = __fadd_rn(f, __fmul_rn(a[k], b[k]));
    return f;
}

__device__ __forceinline__ int32_t draw_negative(const TrainParams &P, uint32_t x0, uint32_t x1)
{
    const int32_t slot = (int32_t)__umulhi(x0, (uint32_t)P.n);
    const float u = (float)(x1 >> 8) * (1.0f / 16777216.0f);
    // the alias table is read-only for the whole kernel: non-coherent path, L1-cacheable
    const int2 e = __ldg(reinterpret_cast<const int2 *>(P.alias) + slot);
    return u < __int_as_float(e.x) ? slot : e.y;
}

// one (centre, context) pair; KT > 0: K known at compile time, all K+1 target rows are loaded
// before the first use (memory-level parallelism); KT == 0: runtime K, one target at a time.
template <int D, int KT>
__device__ __forceinline__ void pair_update(const TrainParams &P, const float *s_exp,
                                            int32_t centre, int32_t context, float alpha,
                                            uint32_t p_lo, uint32_t p_hi, uint32_t q)
{
    float *row1_ptr = P.syn0 + (int64_t)context * D;
    float row1[D], work[D];
    load_row<D>(row1_ptr, row1);
#pragma unroll
    for (int k = 0; k < D; ++k) work[k] = 0.0f;
    const uint32_t ctr2 = q << 4, ctr3 = ((uint32_t)P.epoch << 8) | kStreamSgns;

    auto apply = [&](int32_t target, float label, const float (&row2)[D]) {
        const float f = dot_row<D>(row1, row2);
        if (f <= -kMaxExp || f >= kMaxExp) return;
        const int idx = (int)((double)__fadd_rn(f, kMaxExp) * ((double)kExpTableSize / 6.0 / 2.0));
        const float g = __fmul_rn(__fsub_rn(label, s_exp[idx]), alpha);
#pragma unroll
        for (int k = 0; k < D; ++k) work[k] = __fadd_rn(work[k], __fmul_rn(g, row2[k]));
        float *row2_ptr = P.syn1neg + (int64_t)target * D;
#pragma unroll
        for (int k = 0; k < D; k += 4)
            red_add_v4(row2_ptr + k, __fmul_rn(g, row1[k]), __fmul_rn(g, row1[k + 1]),
                       __fmul_rn(g, row1[k + 2]), __fmul_rn(g, row1[k + 3]));
    };

    if constexpr (KT > 0) {
        int32_t target[KT + 1];
        target[0] = centre;
        u32x4 r = {0, 0, 0, 0};
#pragma unroll
        for (int t = 1; t <= KT; ++t) {
            const int w = 2 * (t - 1);
            if ((w & 3) == 0) r = philox4x32_10(p_lo, p_hi, ctr2 | (uint32_t)(w >> 2), ctr3, P.key0, P.key1);
            target[t] = (w & 3) == 0 ? draw_negative(P, r.x, r.y) : draw_negative(P, r.z, r.w);
        }
        float rows[KT + 1][D];
#pragma unroll
        for (int t = 0; t <= KT; ++t) load_row<D>(P.syn1neg + (int64_t)target[t] * D, rows[t]);
#pragma unroll
        for (int t = 0; t <= KT; ++t) {
            if (t > 0 && target[t] == centre) continue;
            if (t > 0) {
                // an earlier target of THIS pair may be the same row: its update must be visible
                // (the sequential oracle re-reads the row); reload only in that rare case
                bool dup = false;
#pragma unroll
                for (int t2 = 0; t2 < t; ++t2) dup |= (target[t2] == target[t]);
                if (dup) load_row<D>(P.syn1neg + (int64_t)target[t] * D, rows[t]);
            }
            apply(target[t], t == 0 ? 1.0f : 0.0f, rows[t]);
        }
    } else {
        u32x4 r = {0, 0, 0, 0};
        for (int t = 0; t <= P.K; ++t) {
            int32_t target = centre;
            if (t > 0) {
                const int w = 2 * (t - 1);
                if ((w & 3) == 0) r = philox4x32_10(p_lo, p_hi, ctr2 | (uint32_t)(w >> 2), ctr3, P.key0, P.key1);
                target = (w & 3) == 0 ? draw_negative(P, r.x, r.y) : draw_negative(P, r.z, r.w);
                if (target == centre) continue;
            }
            float row2[D];
            load_row<D>(P.syn1neg + (int64_t)target * D, row2);
            apply(target, t == 0 ? 1.0f : 0.0f, row2);
        }
    }
#pragma unroll
    for (int k = 0; k < D; k += 4) red_add_v4(row1_ptr + k, work[k], work[k + 1], work[k + 2], work[k + 3]);
}

template <int D, int KT>
__global__ void __launch_bounds__(kTrainThreads)
sgns_train_kernel(const TrainParams P)
{
    __shared__ float s_exp[kExpTableSize];
    for (int i = threadIdx.x; i < kExpTableSize; i += blockDim.x) s_exp[i] = P.exp_table[i];
    __syncthreads();
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P.W; p += nthreads) {
        // gensim _get_next_alpha at pushed = job*job_walks examples (fp64, unfused, as the oracle)
        const int64_t pushed = (p / P.job_walks) * P.job_walks;
        const double epoch_progress = __ddiv_rn((double)pushed, (double)P.W);
        const double progress = __ddiv_rn(__dadd_rn((double)P.epoch, epoch_progress), (double)P.epochs);
        double next_alpha = __dsub_rn(P.alpha0, __dmul_rn(P.alpha_span, progress));
        if (next_alpha < P.alpha_min) next_alpha = P.alpha_min;
        const float alpha = (float)next_alpha;
        const uint32_t p_lo = (uint32_t)(uint64_t)p, p_hi = (uint32_t)((uint64_t)p >> 32);
        int32_t prev = -1, cur = __ldcs(P.tokens + p);
        for (int32_t i = 0; i < P.L && cur >= 0; ++i) {
            const int32_t next = (i + 1 < P.L) ? __ldcs(P.tokens + (int64_t)(i + 1) * P.W + p) : -1;
            if (prev >= 0) pair_update<D, KT>(P, s_exp, cur, prev, alpha, p_lo, p_hi, 2u * i);
            if (next >= 0) pair_update<D, KT>(P, s_exp, cur, next, alpha, p_lo, p_hi, 2u * i + 1u);
            prev = cur;
            cur = next;
        }
    }
}

// gensim word2vec_inner.pyx init(): EXP_TABLE[i] = e/(e+1), e = exp((i/1000*2-1)*6) in REAL_t
static void host_exp_table(float *t)
{
    for (int i = 0; i < kExpTableSize; ++i) {
        const float x = ((float)i / (float)kExpTableSize * 2.0f - 1.0f) * 6.0f;
        const float e = (float)exp((double)x);
        t[i] = e / (e + 1.0f);
    }
}

template <int D, int KT>
static int launch_train(const TrainParams &P, int64_t max_threads, cudaStream_t s)
{
    int per_sm = 0;
    GW_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sgns_train_kernel<D, KT>,
                                                             kTrainThreads, 0));
    if (per_sm < 1) per_sm = 1;
    const int sms = num_sms();
    int64_t threads = (int64_t)sms * per_sm * kTrainThreads;  // one resident wave, persistent
    if (max_threads > 0 && max_threads < threads) threads = max_threads;
    if (threads > P.W) threads = P.W;
    // `threads` walks are in flight at any time (the Hogwild staleness knob).  Below one full
    // CTA per SM the walks are spread over all SMs in 64-thread CTAs so that every SM's LSU/L1
    // path carries part of the gather/reduction traffic.
    int block = kTrainThreads;
    if (threads < (int64_t)sms * kTrainThreads) block = 64;
    if (threads < block) block = (int)threads;
    const int64_t grid = ceil_div<int64_t>(threads, block);
    sgns_train_kernel<D, KT><<<(unsigned)grid, block, 0, s>>>(P);
    GW_LAUNCH_OK();
    return GW_OK;
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

extern "C" int gw_sgns_train(int32_t n, int32_t d, const int32_t *tokens, int64_t W, int32_t L,
                             int32_t window, int32_t K, int32_t epochs, int32_t epoch_begin,
                             int32_t epoch_end, double alpha0, double alpha_min, int64_t job_walks,
                             const gw_alias_entry *alias_table, float *syn0, float *syn1neg,
                             uint64_t seed, int64_t max_threads, gw_stream_t stream)
{
    using namespace gw;
    cudaStream_t s = (cudaStream_t)stream;
    GW_REQUIRE(n > 0 && tokens && alias_table && syn0 && syn1neg, "gw_sgns_train: null argument");
    GW_REQUIRE(W > 0 && L >= 1 && L <= 4096, "gw_sgns_train: W=%lld L=%d out of range", (long long)W, L);
    GW_REQUIRE(K >= 0 && K <= 32, "gw_sgns_train: negative=%d out of range [0,32]", K);
    GW_REQUIRE(epochs >= 1 && epochs <= 65535 && epoch_begin >= 0 && epoch_begin <= epoch_end &&
                   epoch_end <= epochs,
               "gw_sgns_train: epochs=%d [%d,%d) invalid", epochs, epoch_begin, epoch_end);
    GW_REQUIRE(job_walks >= 1 && max_threads >= 0, "gw_sgns_train: job_walks/max_threads invalid");
    if (window != 1) {
        set_error("gw_sgns_train: window=%d is not implemented (GeneWalk uses window=1)", window);
        return GW_ERR_UNSUPPORTED;
    }
    if (d != 8 && d != 16 && d != 32) {
        set_error("gw_sgns_train: vector_size=%d is not implemented (8, 16, 32)", d);
        return GW_ERR_UNSUPPORTED;
    }
    GW_REQUIRE((reinterpret_cast<uintptr_t>(syn0) & 15) == 0 && (reinterpret_cast<uintptr_t>(syn1neg) & 15) == 0,
               "gw_sgns_train: tables must be 16-byte aligned");
    float h_exp[kExpTableSize];
    host_exp_table(h_exp);
    Temp d_exp;
    GW_CUDA_OK(d_exp.alloc(sizeof h_exp, s));
    // pageable source: the copy is staged before the call returns, so the stack buffer is safe
    GW_CUDA_OK(cudaMemcpyAsync(d_exp.ptr, h_exp, sizeof h_exp, cudaMemcpyHostToDevice, s));
    TrainParams P;
    P.tokens = tokens; P.W = W; P.L = L; P.n = n; P.K = K; P.epochs = epochs; P.epoch = 0;
    P.alpha0 = alpha0; P.alpha_span = alpha0 - alpha_min; P.alpha_min = alpha_min;
    P.job_walks = job_walks; P.alias = alias_table; P.exp_table = d_exp.as<float>();
    P.syn0 = syn0; P.syn1neg = syn1neg; P.key0 = (uint32_t)seed; P.key1 = (uint32_t)(seed >> 32);
    for (int ep = epoch_begin; ep < epoch_end; ++ep) {
        P.epoch = ep;
        int rc;
        if (d == 8) rc = (K == 5) ? launch_train<8, 5>(P, max_threads, s) : launch_train<8, 0>(P, max_threads, s);
        else if (d == 16) rc = (K == 5) ? launch_train<16, 5>(P, max_threads, s) : launch_train<16, 0>(P, max_threads, s);
        else rc = launch_train<32, 0>(P, max_threads, s);
        GW_TRY(rc);
    }
    return GW_OK;
}
