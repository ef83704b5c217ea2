// gw_sgns_train.cu -- skip-gram negative-sampling training on L2-resident embedding tables.
// Replaces gensim.models.Word2Vec(sg=1, window=1, negative=K, sample=0).train as reached from
// /root/reference/genewalk/deepwalk.py:135-139 (gensim is third-party and absent from the
// reference tree; the update rule is its published w2v_fast_sentence_sg_neg, the schedule its
// _job_producer/_get_next_alpha, the parallel structure its worker threads).
//
// Parallel structure = gensim's.  The corpus is cut into jobs of `job_walks` consecutive walks
// (gensim: batch_words = 10000 words = 1000 walks of 10) with one learning rate per job; gensim's
// `workers` threads each train one job at a time and race on the tables (Hogwild).  Here a worker
// is a LANE GROUP of G threads (G = 4 at d = 8); `lanes` of them run concurrently and pull jobs from
// a device-side queue in corpus order, so that, like with gensim, the jobs in flight are always
// consecutive and every job is trained sequentially by its worker.  lanes = 1 is the sequential
// statement of the algorithm (bit-exact against the CPU oracle); `lanes` is the staleness knob.
//
// Data layout / roofline.  syn0, syn1neg: [n, d] fp32 row-major, L2-resident for the whole run
// (2*n*32 B = 2.4 MB at n = 38k).  At d = 8 a row is one 32-byte sector; the G = 4 lanes of a group
// each own 2 consecutive floats of every row, so one warp-level LDG.64 / RED.v2 touches exactly one
// sector per group: 1 sector read + 1 sector reduction per row visit, the minimum (measured rates on
// B200: 280 G row gathers/s, 121 G row reductions/s, tools/l2_microbench.cu).  One STEP = the two
// pairs (c, a), (c, b) of a centre c = walk[i] with contexts a = walk[i-1], b = walk[i+1]: 13 row
// gathers issued together (centre, 2 contexts, 2K negatives), 6+6 dot products reduced by xor-
// shuffles inside the group, 14 row reductions.  Per pair: 6.5 sector reads + 7 sector reductions +
// K alias entries (L1-resident, ld.global.nc) + token loads; algorithmic bytes 4d(2K+4)+8+8K = 496 B.
#include <math.h>

#include "gw_common.cuh"

namespace gw {

constexpr int kExpTableSize = 1000;
constexpr float kMaxExp = 6.0f;
constexpr int kTrainThreads = 512;      // upper bound of the CTA size (launch bounds)
constexpr int kSmemAliasMaxN = 56000;   // 4 B/entry packed alias table + EXP_TABLE fit 227 KB

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
    int64_t job_walks;    // walks per learning-rate job (gensim: 1000 walks of 10 words)
    int64_t chunk_walks;  // walks a worker takes from the queue at a time (divides job_walks)
    int64_t streams;      // corpus slices the queue deals chunks from in turn (see gw_sgns_train)
    int64_t stream_chunks;  // chunks per slice
    int32_t shared_sm;      // leave the SMs' shared memory to other kernels (concurrent replicates)
    int32_t merge;          // merged row reductions (Hogwild mode), see step_update_merged
    unsigned long long *job_queue;  // next job to hand out (device counter, zeroed per launch)
    const gw_alias_entry *alias;
    const float *exp_table;  // [1000] device copy of gensim's EXP_TABLE
    float *syn0;
    float *syn1neg;
    uint32_t key0, key1;
};

// ---- per-lane slice of a row: E consecutive floats ---------------------------------------------
template <int E> __device__ __forceinline__ void ld_slice(const float *p, float (&r)[E])
{
    // L2 (never a stale L1 line: other SMs update these rows)
    if constexpr (E == 1) {
        r[0] = __ldcg(p);
    } else if constexpr (E == 2) {
        const float2 v = __ldcg(reinterpret_cast<const float2 *>(p));
        r[0] = v.x; r[1] = v.y;
    } else {
#pragma unroll
        for (int k = 0; k < E; k += 4) {
            const float4 v = __ldcg(reinterpret_cast<const float4 *>(p + k));
            r[k] = v.x; r[k + 1] = v.y; r[k + 2] = v.z; r[k + 3] = v.w;
        }
    }
}

template <int E> __device__ __forceinline__ void red_slice(float *p, const float (&d)[E])
{
    if constexpr (E == 1) {
        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(d[0]) : "memory");
    } else if constexpr (E == 2) {
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(d[0]), "f"(d[1]) : "memory");
    } else {
#pragma unroll
        for (int k = 0; k < E; k += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p + k), "f"(d[k]),
                         "f"(d[k + 1]), "f"(d[k + 2]), "f"(d[k + 3])
                         : "memory");
    }
}

// dot product of two rows spread over the G lanes of a group: sequential over the lane's E
// elements, then a xor-butterfly over the group (every lane ends with the same bits; the CPU oracle
// sums in the same order).  All arithmetic unfused.
template <int E, int G>
__device__ __forceinline__ float group_dot(const float (&a)[E], const float (&b)[E], unsigned gmask)
{
    float s = __fmul_rn(a[0], b[0]);
#pragma unroll
    for (int k = 1; k < E; ++k) s = __fadd_rn(s, __fmul_rn(a[k], b[k]));
#pragma unroll
    for (int m = 1; m < G; m <<= 1) s = __fadd_rn(s, __shfl_xor_sync(gmask, s, m));
    return s;
}

// acceptance threshold of an alias slot for the 16-bit coin x1 >> 16
__device__ __forceinline__ uint32_t coin_threshold(float prob)
{
    const uint32_t t = (uint32_t)floorf(__fmul_rn(prob, 65536.0f));
    return t > 65535u ? 65535u : t;
}

__device__ __forceinline__ int32_t draw_negative(const TrainParams &P, uint32_t x0, uint32_t x1)
{
    const int32_t slot = (int32_t)__umulhi(x0, (uint32_t)P.n);
    // the alias table is read-only for the whole kernel: non-coherent path, L1-cacheable
    const int2 e = __ldg(reinterpret_cast<const int2 *>(P.alias) + slot);
    return (x1 >> 16) < coin_threshold(__int_as_float(e.x)) ? slot : e.y;
}

// the K negative targets of pair q of walk w in this epoch (every lane of the group computes them)
template <int KMAX>
__device__ __forceinline__ void draw_targets(const TrainParams &P, int K, uint32_t w_lo, uint32_t w_hi,
                                             uint32_t q, int32_t (&t)[KMAX])
{
    const uint32_t ctr2 = q << 4, ctr3 = ((uint32_t)P.epoch << 8) | kStreamSgns;
    u32x4 r = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        if (k < K) {
            const int w = 2 * k;
            if ((w & 3) == 0) r = philox4x32_10(w_lo, w_hi, ctr2 | (uint32_t)(w >> 2), ctr3, P.key0, P.key1);
            t[k] = (w & 3) == 0 ? draw_negative(P, r.x, r.y) : draw_negative(P, r.z, r.w);
        }
    }
}

// gensim: EXP_TABLE[(int)((f + MAX_EXP) * (EXP_TABLE_SIZE / MAX_EXP / 2))].  For the largest floats
// below 6 the float sum f + 6 rounds up to 12 and the index becomes 1000, one past the table (gensim
// then reads whatever follows EXP_TABLE in memory); the index is clamped to the last entry here.
__device__ __forceinline__ int sigmoid_index(float f)
{
    const int i = (int)((double)__fadd_rn(f, kMaxExp) * ((double)kExpTableSize / 6.0 / 2.0));
    return i < kExpTableSize ? i : kExpTableSize - 1;
}

// ---- one (centre, context) pair, one target at a time (any K; also the hazard path) -------------
// Rows are (re)loaded right before use, after this thread's earlier reductions in program order,
// so a target that repeats inside the pair sees its own earlier update, like the sequential oracle.
template <int D, int G>
__device__ __forceinline__ void pair_sequential(const TrainParams &P, const float *s_exp, int sub,
                                                unsigned gmask, int32_t centre, int32_t context,
                                                float alpha, const int32_t *neg /*[K]*/)
{
    constexpr int E = D / G;
    float *ctx_ptr = P.syn0 + (int64_t)context * D + sub * E;
    float row1[E], work[E];
    ld_slice<E>(ctx_ptr, row1);
#pragma unroll
    for (int k = 0; k < E; ++k) work[k] = 0.0f;
    for (int t = 0; t <= P.K; ++t) {
        const int32_t target = t == 0 ? centre : neg[t - 1];
        if (t > 0 && target == centre) continue;
        float *tp = P.syn1neg + (int64_t)target * D + sub * E;
        float row2[E];
        ld_slice<E>(tp, row2);
        const float f = group_dot<E, G>(row1, row2, gmask);
        if (f <= -kMaxExp || f >= kMaxExp) continue;
        const float g = __fmul_rn(__fsub_rn(t == 0 ? 1.0f : 0.0f, s_exp[sigmoid_index(f)]), alpha);
        float delta[E];
#pragma unroll
        for (int k = 0; k < E; ++k) {
            work[k] = __fadd_rn(work[k], __fmul_rn(g, row2[k]));
            delta[k] = __fmul_rn(g, row1[k]);
        }
        red_slice<E>(tp, delta);
    }
    red_slice<E>(ctx_ptr, work);
}

// ---- negative draws of one step, shared out over the lanes of the group (K = KT) ----------------
// A step needs 2*KT negatives = 2*ceil(KT/2) Philox calls.  Call c (pair c / CP, call c % CP of
// that pair, CP = ceil(KT/2)) is computed by lane c % G, which also resolves the call's (up to) two
// negatives through the alias table; resolve() then hands every lane all 2*KT targets by shuffles.
template <int G, int KT, bool kSmemAlias> struct StepDraw {
    static constexpr int CP = (KT + 1) / 2;            // Philox calls per pair
    static constexpr int NC = 2 * CP;                  // calls per step
    static constexpr int R = (NC + G - 1) / G;         // calls per lane (rounds)
    int32_t slot[R][2];                                // this lane's draws: alias slot ...
    uint32_t coin[R][2];                               // ... its 16-bit coin ...
    int2 entry[R][2];                                  // ... and the alias entry (load in flight)

    // s_alias: packed (threshold << 16 | alias) copy of the table in shared memory, or unused
    __device__ __forceinline__ void issue(const TrainParams &P, const uint32_t *s_alias, int sub,
                                          uint32_t w_lo, uint32_t w_hi, uint32_t q_first)
    {
        const uint32_t ctr3 = ((uint32_t)P.epoch << 8) | kStreamSgns;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int c = sub + r * G;                 // call id of this lane in round r
            const int pair = c / CP, call = c - pair * CP;
            const bool used = c < NC;
            const bool second = used && (2 * call + 1 < KT);   // the call's second negative exists
            if (used) {
                const u32x4 x = philox4x32_10(w_lo, w_hi, ((q_first + (uint32_t)pair) << 4) | (uint32_t)call,
                                              ctr3, P.key0, P.key1);
                slot[r][0] = (int32_t)__umulhi(x.x, (uint32_t)P.n);
                slot[r][1] = (int32_t)__umulhi(x.z, (uint32_t)P.n);
                coin[r][0] = x.y >> 16;
                coin[r][1] = x.w >> 16;
            }
            if constexpr (kSmemAlias) {
                if (used) entry[r][0].x = (int)s_alias[slot[r][0]];
                if (second) entry[r][1].x = (int)s_alias[slot[r][1]];
            } else {
                // the alias table is read-only for the whole kernel: non-coherent path, L1-cacheable
                if (used) entry[r][0] = __ldg(reinterpret_cast<const int2 *>(P.alias) + slot[r][0]);
                if (second) entry[r][1] = __ldg(reinterpret_cast<const int2 *>(P.alias) + slot[r][1]);
            }
        }
    }

    __device__ __forceinline__ void resolve(int leader, unsigned gmask, int32_t (&tg)[2 * KT]) const
    {
        int32_t mine[R][2];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if constexpr (kSmemAlias) {
                    const uint32_t e = (uint32_t)entry[r][j].x;
                    mine[r][j] = coin[r][j] < (e >> 16) ? slot[r][j] : (int32_t)(e & 0xffffu);
                } else {
                    mine[r][j] = coin[r][j] < coin_threshold(__int_as_float(entry[r][j].x)) ? slot[r][j]
                                                                                         : entry[r][j].y;
                }
            }
#pragma unroll
        for (int x = 0; x < 2 * KT; ++x) {
            const int pair = x / KT, k = x - pair * KT;   // negative k of pair `pair`
            const int c = pair * CP + (k >> 1);           // its call, owner lane and round
            tg[x] = __shfl_sync(gmask, mine[c / G][k & 1], leader + (c % G));
        }
    }
};

// ---- one step = both pairs of a centre, all rows gathered up front (K = KT at compile time) -----
template <int D, int G, int KT> struct StepRows {
    static constexpr int E = D / G;
    float Rc[E], Ra[E], Rb[E], N[2 * KT][E];
};

template <int D, int G, int KT>
__device__ __forceinline__ void step_gather(const TrainParams &P, int sub, int32_t c, int32_t a, int32_t b,
                                            const int32_t (&tg)[2 * KT], StepRows<D, G, KT> &R)
{
    constexpr int E = D / G;
    const int64_t off = sub * E;
    ld_slice<E>(P.syn1neg + (int64_t)c * D + off, R.Rc);
    ld_slice<E>(P.syn0 + (int64_t)(a >= 0 ? a : c) * D + off, R.Ra);
    ld_slice<E>(P.syn0 + (int64_t)(b >= 0 ? b : c) * D + off, R.Rb);
#pragma unroll
    for (int x = 0; x < 2 * KT; ++x) ld_slice<E>(P.syn1neg + (int64_t)tg[x] * D + off, R.N[x]);
}

template <int D, int G, int KT>
__device__ __forceinline__ void step_update(const TrainParams &P, const float *s_exp, int sub,
                                            unsigned gmask, int32_t c, int32_t a, int32_t b, float alpha,
                                            const int32_t (&tg)[2 * KT], StepRows<D, G, KT> &R)
{
    constexpr int E = D / G;
    const bool hasA = a >= 0, hasB = b >= 0;
    const int64_t off = sub * E;
    float *c_ptr = P.syn1neg + (int64_t)c * D + off;
    // a negative that repeats inside the step must see its earlier update (the sequential oracle
    // re-reads the row): such targets are re-gathered after the earlier reduction was issued
    bool dup[2 * KT];
#pragma unroll
    for (int x = 0; x < 2 * KT; ++x) {
        dup[x] = false;
#pragma unroll
        for (int y = 0; y < x; ++y) dup[x] |= (tg[y] == tg[x]);
    }
    // one pair against its K+1 targets; `Rc` is updated in registers so that the second pair of
    // the step sees the first pair's update of the shared centre row (same bits as the memory
    // value when no other worker interferes)
    auto pair = [&](float (&ctx)[E], float *ctx_ptr, const int o) {
        float f[KT + 1];
        f[0] = group_dot<E, G>(ctx, R.Rc, gmask);
#pragma unroll
        for (int k = 0; k < KT; ++k) f[k + 1] = group_dot<E, G>(ctx, R.N[o + k], gmask);
        float work[E];
#pragma unroll
        for (int k = 0; k < E; ++k) work[k] = 0.0f;
#pragma unroll
        for (int t = 0; t <= KT; ++t) {
            const int x = o + (t > 0 ? t - 1 : 0);
            float *tp = t == 0 ? c_ptr : P.syn1neg + (int64_t)tg[x] * D + off;
            float (&row)[E] = t == 0 ? R.Rc : R.N[x];
            float ft = f[t];
            if (t > 0) {
                if (tg[x] == c) continue;
                if (dup[x]) {
                    ld_slice<E>(tp, row);
                    ft = group_dot<E, G>(ctx, row, gmask);
                }
            }
            if (ft <= -kMaxExp || ft >= kMaxExp) continue;
            const float g = __fmul_rn(__fsub_rn(t == 0 ? 1.0f : 0.0f, s_exp[sigmoid_index(ft)]), alpha);
            float delta[E];
#pragma unroll
            for (int k = 0; k < E; ++k) {
                work[k] = __fadd_rn(work[k], __fmul_rn(g, row[k]));
                delta[k] = __fmul_rn(g, ctx[k]);
                if (t == 0) row[k] = __fadd_rn(row[k], delta[k]);
            }
            red_slice<E>(tp, delta);
        }
        red_slice<E>(ctx_ptr, work);
#pragma unroll
        for (int k = 0; k < E; ++k) ctx[k] = __fadd_rn(ctx[k], work[k]);
    };
    if (hasA) pair(R.Ra, P.syn0 + (int64_t)a * D + off, 0);
    if (hasB) {
        if (hasA && b == a) pair(R.Ra, P.syn0 + (int64_t)a * D + off, KT);  // walk stepped back
        else pair(R.Rb, P.syn0 + (int64_t)b * D + off, KT);
    }
}

// ---- merged reductions (Hogwild mode) ------------------------------------------------------------
// Over a walk every context row syn0[w_j] is the input of two pairs, (c = w_{j-1}, w_j) in step j-1
// and (c = w_{j+1}, w_j) in step j+1, and the centre row syn1neg[w_j] is a target of both pairs of
// step j.  These are the rows whose popularity is proportional to the node degree, i.e. the rows on
// which the reductions of different workers meet in L2.  With other workers racing on the tables the
// order in which ONE worker's own updates reach memory is immaterial, so in Hogwild mode a worker
// holds the update of syn0[w_j] from step j-1 back (pending, in registers) and issues ONE reduction
// with the sum of both updates in step j+1, and ONE reduction for the centre row per step: 13
// gathers + 12 reductions per step instead of 13 + 14, the degree-proportional reductions halved.
// Every row is still gathered fresh from L2 right before use (a view carried in registers over two
// steps would miss the other workers' updates: measured, it costs as much fidelity as tripling the
// worker count); the worker adds its own pending update to the fresh row, so its arithmetic sees
// every one of its updates.  Run sequentially the result differs from the per-pair order only by
// float summation order; the oracle restates this order too (gw_oracle.c:walk_merged, bit-exact).
template <int E> struct CtxCarry {
    float pX[E];           // update of syn0[walk[i-1]] not yet reduced into memory (from step i-2)
    float pY[E];           // the same for syn0[walk[i]] (from step i-1, due in step i+1)
    int32_t xn, yn;        // their nodes, -1 = nothing pending
};

template <int D, int G, int KT>
__device__ __forceinline__ void step_update_merged(const TrainParams &P, const float *s_exp, int sub,
                                                   unsigned gmask, int32_t c, int32_t a, int32_t b,
                                                   float alpha, const int32_t (&tg)[2 * KT],
                                                   StepRows<D, G, KT> &R, CtxCarry<D / G> &carry)
{
    constexpr int E = D / G;
    const bool hasA = a >= 0, hasB = b >= 0;
    const int64_t off = sub * E;
    bool dup[2 * KT];
#pragma unroll
    for (int x = 0; x < 2 * KT; ++x) {
        dup[x] = false;
#pragma unroll
        for (int y = 0; y < x; ++y) dup[x] |= (tg[y] == tg[x]);
    }
    float dC[E];
    bool anyC = false;
#pragma unroll
    for (int k = 0; k < E; ++k) dC[k] = 0.0f;
    // one pair against its K+1 targets; negatives are reduced at once, the centre's update is
    // summed into dC (and into the register copy Rc, which the second pair reads)
    auto pair = [&](float (&ctx)[E], float (&work)[E], const int o) {
        float f[KT + 1];
        f[0] = group_dot<E, G>(ctx, R.Rc, gmask);
#pragma unroll
        for (int k = 0; k < KT; ++k) f[k + 1] = group_dot<E, G>(ctx, R.N[o + k], gmask);
#pragma unroll
        for (int k = 0; k < E; ++k) work[k] = 0.0f;
#pragma unroll
        for (int t = 0; t <= KT; ++t) {
            const int x = o + (t > 0 ? t - 1 : 0);
            float *tp = P.syn1neg + (int64_t)(t == 0 ? c : tg[x]) * D + off;
            float (&row)[E] = t == 0 ? R.Rc : R.N[x];
            float ft = f[t];
            if (t > 0) {
                if (tg[x] == c) continue;
                if (dup[x]) {
                    ld_slice<E>(tp, row);
                    ft = group_dot<E, G>(ctx, row, gmask);
                }
            }
            if (ft <= -kMaxExp || ft >= kMaxExp) continue;
            const float g = __fmul_rn(__fsub_rn(t == 0 ? 1.0f : 0.0f, s_exp[sigmoid_index(ft)]), alpha);
            float delta[E];
#pragma unroll
            for (int k = 0; k < E; ++k) {
                work[k] = __fadd_rn(work[k], __fmul_rn(g, row[k]));
                delta[k] = __fmul_rn(g, ctx[k]);
            }
            if (t == 0) {
#pragma unroll
                for (int k = 0; k < E; ++k) {
                    row[k] = __fadd_rn(row[k], delta[k]);
                    dC[k] = __fadd_rn(dC[k], delta[k]);
                }
                anyC = true;
            } else {
                red_slice<E>(tp, delta);
            }
        }
#pragma unroll
        for (int k = 0; k < E; ++k) ctx[k] = __fadd_rn(ctx[k], work[k]);
    };
    float work[E];
    if (hasA) {
        const bool pending = carry.xn == a;
        if (pending) {   // fresh row + this worker's own update still in registers
#pragma unroll
            for (int k = 0; k < E; ++k) R.Ra[k] = __fadd_rn(R.Ra[k], carry.pX[k]);
        }
        pair(R.Ra, work, 0);
        if (pending) {
#pragma unroll
            for (int k = 0; k < E; ++k) work[k] = __fadd_rn(carry.pX[k], work[k]);
        }
        red_slice<E>(P.syn0 + (int64_t)a * D + off, work);
    } else if (carry.xn >= 0) {
        red_slice<E>(P.syn0 + (int64_t)carry.xn * D + off, carry.pX);   // not reachable in a walk
    }
    // the update of syn0[c] held back by the previous step is due in the next step
#pragma unroll
    for (int k = 0; k < E; ++k) carry.pX[k] = carry.pY[k];
    carry.xn = carry.yn;
    carry.yn = -1;
    if (hasB) {
        if (hasA && b == a) {   // the walk stepped back: continue on the updated row of pair A
#pragma unroll
            for (int k = 0; k < E; ++k) R.Rb[k] = R.Ra[k];
        }
        pair(R.Rb, carry.pY, KT);
        carry.yn = b;
    }
    if (anyC) red_slice<E>(P.syn1neg + (int64_t)c * D + off, dC);
}

template <int D, int G>
__device__ __forceinline__ void carry_flush(const TrainParams &P, int sub, CtxCarry<D / G> &carry)
{
    constexpr int E = D / G;
    if (carry.xn >= 0) red_slice<E>(P.syn0 + (int64_t)carry.xn * D + sub * E, carry.pX);
    if (carry.yn >= 0) red_slice<E>(P.syn0 + (int64_t)carry.yn * D + sub * E, carry.pY);
    carry.xn = carry.yn = -1;
}

template <int D, int G, int KT, bool kSmemAlias, bool kMerge>
__global__ void __launch_bounds__(kTrainThreads)
sgns_train_kernel(const TrainParams P)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    float *s_exp = reinterpret_cast<float *>(s_raw);
    uint32_t *s_alias = reinterpret_cast<uint32_t *>(s_raw + sizeof(float) * kExpTableSize);
    for (int i = threadIdx.x; i < kExpTableSize; i += blockDim.x) s_exp[i] = P.exp_table[i];
    if constexpr (kSmemAlias) {
        // packed copy of the noise table: (16-bit acceptance threshold << 16) | alias (n <= 56000)
        for (int i = threadIdx.x; i < P.n; i += blockDim.x) {
            const int2 e = __ldg(reinterpret_cast<const int2 *>(P.alias) + i);
            s_alias[i] = (coin_threshold(__int_as_float(e.x)) << 16) | (uint32_t)e.y;
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int sub = lane & (G - 1);
    const int leader = lane & ~(G - 1);
    const unsigned gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << leader;
    const int64_t nchunks = ceil_div<int64_t>(P.W, P.chunk_walks);
    for (;;) {
        // gensim's job queue: a worker takes the next chunk of the corpus when it finishes one
        unsigned long long chunk = 0;
        if (sub == 0) chunk = atomicAdd(P.job_queue, 1ULL);
        chunk = __shfl_sync(gmask, chunk, leader);
        if (chunk >= (unsigned long long)(P.streams * P.stream_chunks)) break;
        // the queue deals chunks from the `streams` corpus slices in turn, so the chunks in flight
        // are consecutive inside each slice and at most lanes/streams workers share a start node
        const int64_t slice = (int64_t)(chunk % (unsigned long long)P.streams);
        const int64_t pos = slice * P.stream_chunks + (int64_t)(chunk / (unsigned long long)P.streams);
        if (pos >= nchunks) continue;   // ragged last slice
        const int64_t first = pos * P.chunk_walks;
        const int64_t last = first + P.chunk_walks < P.W ? first + P.chunk_walks : P.W;
        // gensim _get_next_alpha at pushed = job*job_walks examples (fp64, unfused, as the oracle);
        // chunk_walks divides job_walks, so a chunk lies inside one job
        const int64_t pushed = (first / P.job_walks) * P.job_walks;
        const double epoch_progress = __ddiv_rn((double)pushed, (double)P.W);
        const double progress = __ddiv_rn(__dadd_rn((double)P.epoch, epoch_progress), (double)P.epochs);
        double next_alpha = __dsub_rn(P.alpha0, __dmul_rn(P.alpha_span, progress));
        if (next_alpha < P.alpha_min) next_alpha = P.alpha_min;
        const float alpha = (float)next_alpha;
        for (int64_t w = first; w < last; ++w) {
            const uint32_t w_lo = (uint32_t)(uint64_t)w, w_hi = (uint32_t)((uint64_t)w >> 32);
            int32_t prev = -1, cur = __ldg(P.tokens + w);
            int32_t next = P.L > 1 ? __ldg(P.tokens + P.W + w) : -1;
            if constexpr (KT > 0) {
                StepDraw<G, KT, kSmemAlias> draw;
                draw.issue(P, s_alias, sub, w_lo, w_hi, 0u);
                CtxCarry<D / G> carry;
                carry.xn = carry.yn = -1;
                for (int32_t i = 0; i < P.L && cur >= 0; ++i) {
                    // token two positions ahead (the next step's right context), loaded early
                    const int32_t next2 = (i + 2 < P.L && next >= 0) ? __ldg(P.tokens + (int64_t)(i + 2) * P.W + w) : -1;
                    int32_t tg[2 * KT];
                    draw.resolve(leader, gmask, tg);
                    if (prev < 0) {
#pragma unroll
                        for (int k = 0; k < KT; ++k) tg[k] = cur;         // absent pair: valid, skipped
                    }
                    if (next < 0) {
#pragma unroll
                        for (int k = 0; k < KT; ++k) tg[KT + k] = cur;
                    }
                    StepRows<D, G, KT> rows;
                    if constexpr (kMerge) {
                        step_gather<D, G, KT>(P, sub, cur, prev, next, tg, rows);
                        draw.issue(P, s_alias, sub, w_lo, w_hi, 2u * (uint32_t)(i + 1));
                        step_update_merged<D, G, KT>(P, s_exp, sub, gmask, cur, prev, next, alpha, tg,
                                                     rows, carry);
                    } else {
                        step_gather<D, G, KT>(P, sub, cur, prev, next, tg, rows);
                        // the next step's draws (Philox + alias loads) overlap this step's row gathers
                        draw.issue(P, s_alias, sub, w_lo, w_hi, 2u * (uint32_t)(i + 1));
                        step_update<D, G, KT>(P, s_exp, sub, gmask, cur, prev, next, alpha, tg, rows);
                    }
                    prev = cur;
                    cur = next;
                    next = next2;
                }
                if constexpr (kMerge) carry_flush<D, G>(P, sub, carry);
            } else {
                for (int32_t i = 0; i < P.L && cur >= 0; ++i) {
                    int32_t neg[32];
                    if (prev >= 0) {
                        draw_targets<32>(P, P.K, w_lo, w_hi, 2u * i, neg);
                        pair_sequential<D, G>(P, s_exp, sub, gmask, cur, prev, alpha, neg);
                    }
                    if (next >= 0) {
                        draw_targets<32>(P, P.K, w_lo, w_hi, 2u * i + 1u, neg);
                        pair_sequential<D, G>(P, s_exp, sub, gmask, cur, next, alpha, neg);
                    }
                    prev = cur;
                    cur = next;
                    next = (i + 2 < P.L && cur >= 0) ? __ldg(P.tokens + (int64_t)(i + 2) * P.W + w) : -1;
                }
            }
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

template <int D, int G, int KT>
static int launch_train(const TrainParams &P, int64_t lanes, cudaStream_t s)
{
    // `lanes` workers of G threads each.  With the noise table in shared memory (n <= 56000) one
    // CTA per SM holds lanes/num_sms workers (<= 512 threads); otherwise 128-thread CTAs.
    const bool smem_alias = KT > 0 && P.n <= kSmemAliasMaxN && !P.shared_sm;
    const int sms = num_sms();
    int block, blocks;
    if (lanes * G <= 32) {
        block = (int)lanes * G;
        blocks = 1;
    } else if (smem_alias) {
        int64_t per_sm = ceil_div<int64_t>(lanes, sms) * G;
        per_sm = ceil_div<int64_t>(per_sm, 32) * 32;
        block = (int)(per_sm > kTrainThreads ? kTrainThreads : per_sm);
        blocks = (int)ceil_div<int64_t>(lanes * G, block);
        if (blocks > sms) blocks = sms;  // more would not be resident next to a 150+ KB table
    } else {
        block = 128;
        blocks = (int)ceil_div<int64_t>(lanes * G, block);
    }
    size_t smem = sizeof(float) * kExpTableSize + (smem_alias ? sizeof(uint32_t) * (size_t)P.n : 0);
    GW_CUDA_OK(cudaMemsetAsync(P.job_queue, 0, sizeof(unsigned long long), s));
    constexpr bool kCanMerge = KT > 0;
    if (smem_alias) {
        auto kern = (kCanMerge && P.merge) ? sgns_train_kernel<D, G, KT, true, kCanMerge>
                                           : sgns_train_kernel<D, G, KT, true, false>;
        GW_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<blocks, block, smem, s>>>(P);
    } else {
        auto kern = (kCanMerge && P.merge) ? sgns_train_kernel<D, G, KT, false, kCanMerge>
                                           : sgns_train_kernel<D, G, KT, false, false>;
        kern<<<blocks, block, smem, s>>>(P);
    }
    GW_LAUNCH_OK();
    return GW_OK;
}

// largest divisor of job_walks that is <= want (>= 1)
static int64_t chunk_divisor(int64_t job_walks, int64_t want)
{
    if (want >= job_walks) return job_walks;
    if (want < 1) want = 1;
    for (int64_t c = want; c > 1; --c)
        if (job_walks % c == 0) return c;
    return 1;
}

}  // namespace gw

// Default parallel shape.  Fidelity constraints (parity study, profiles/):
// the Hogwild staleness per table row stays small when the workers stay below ~n/4 and at most
// ~128 of them train walks that start at the same node; training keeps sweeping the corpus in the
// reference's node-major order when few slices are active and the chunks in flight inside a slice
// cover little more than one start node.  Within that: as many workers as fill the GPU.
extern "C" int gw_sgns_auto_shape(int32_t n, int64_t W, int64_t job_walks, int32_t concurrent,
                                  int64_t *lanes_out, int64_t *chunk_out, int64_t *streams_out)
{
    using namespace gw;
    GW_REQUIRE(n > 0 && W > 0 && job_walks > 0 && concurrent >= 1 && lanes_out && chunk_out && streams_out,
               "gw_sgns_auto_shape: bad arguments");
    // 128 workers (512 threads at d = 8) per SM are resident at the kernel's register budget; the
    // trainings that run at once share them (all co-resident: measured optimum)
    int64_t lanes = (int64_t)num_sms() * 128 / concurrent;
    if (lanes > n / 4) lanes = n / 4;
    if (lanes < 1) lanes = 1;
    const int64_t chunk = chunk_divisor(job_walks, 8);
    const int64_t nchunks = ceil_div<int64_t>(W, chunk);
    if (lanes > nchunks) lanes = nchunks;
    int64_t streams = ceil_div<int64_t>(lanes, 128);
    const int64_t max_streams = n / 50 > 0 ? n / 50 : 1;
    if (streams > max_streams) streams = max_streams;
    *lanes_out = lanes;
    *chunk_out = chunk;
    *streams_out = streams;
    return GW_OK;
}

extern "C" int gw_sgns_train(int32_t n, int32_t d, const int32_t *tokens, int64_t W, int32_t L,
                             int32_t window, int32_t K, int32_t epochs, int32_t epoch_begin,
                             int32_t epoch_end, double alpha0, double alpha_min, int64_t job_walks,
                             int64_t chunk_walks, int64_t streams, const gw_alias_entry *alias_table,
                             float *syn0, float *syn1neg, uint64_t seed, int64_t lanes,
                             int32_t flags, gw_stream_t stream)
{
    using namespace gw;
    cudaStream_t s = (cudaStream_t)stream;
    GW_REQUIRE(n > 0 && tokens && alias_table && syn0 && syn1neg, "gw_sgns_train: null argument");
    GW_REQUIRE(W > 0 && L >= 1 && L <= 4096, "gw_sgns_train: W=%lld L=%d out of range", (long long)W, L);
    GW_REQUIRE(K >= 0 && K <= 32, "gw_sgns_train: negative=%d out of range [0,32]", K);
    GW_REQUIRE(epochs >= 1 && epochs <= 65535 && epoch_begin >= 0 && epoch_begin <= epoch_end &&
                   epoch_end <= epochs,
               "gw_sgns_train: epochs=%d [%d,%d) invalid", epochs, epoch_begin, epoch_end);
    GW_REQUIRE(job_walks >= 1 && lanes >= 0 && chunk_walks >= 0 && streams >= 0,
               "gw_sgns_train: job_walks/chunk_walks/streams/lanes invalid");
    GW_REQUIRE(chunk_walks == 0 || (chunk_walks <= job_walks && job_walks % chunk_walks == 0),
               "gw_sgns_train: chunk_walks=%lld must divide job_walks=%lld", (long long)chunk_walks,
               (long long)job_walks);
    if (window != 1) {
        set_error("gw_sgns_train: window=%d is not implemented (GeneWalk uses window=1)", window);
        return GW_ERR_UNSUPPORTED;
    }
    if (d != 8 && d != 16 && d != 32 && d != 64 && d != 128) {
        set_error("gw_sgns_train: vector_size=%d is not implemented (8, 16, 32, 64, 128)", d);
        return GW_ERR_UNSUPPORTED;
    }
    GW_REQUIRE((reinterpret_cast<uintptr_t>(syn0) & 15) == 0 && (reinterpret_cast<uintptr_t>(syn1neg) & 15) == 0,
               "gw_sgns_train: tables must be 16-byte aligned");
    int64_t auto_lanes = 0, auto_chunk = 0, auto_streams = 0;
    GW_TRY(gw_sgns_auto_shape(n, W, job_walks, 1, &auto_lanes, &auto_chunk, &auto_streams));
    if (lanes == 0) lanes = auto_lanes;
    if (chunk_walks == 0) chunk_walks = lanes == 1 ? job_walks : auto_chunk;
    const int64_t nchunks = ceil_div<int64_t>(W, chunk_walks);
    if (lanes > nchunks) lanes = nchunks;
    if (streams == 0) streams = lanes == 1 ? 1 : (auto_streams < lanes ? auto_streams : lanes);
    if (streams > nchunks) streams = nchunks;
    float h_exp[kExpTableSize];
    host_exp_table(h_exp);
    Temp d_exp, d_queue;
    GW_CUDA_OK(d_exp.alloc(sizeof h_exp, s));
    GW_CUDA_OK(d_queue.alloc(sizeof(unsigned long long), s));
    // pageable source: the copy is staged before the call returns, so the stack buffer is safe
    GW_CUDA_OK(cudaMemcpyAsync(d_exp.ptr, h_exp, sizeof h_exp, cudaMemcpyHostToDevice, s));
    TrainParams P;
    P.tokens = tokens; P.W = W; P.L = L; P.n = n; P.K = K; P.epochs = epochs; P.epoch = 0;
    P.alpha0 = alpha0; P.alpha_span = alpha0 - alpha_min; P.alpha_min = alpha_min;
    P.job_walks = job_walks; P.chunk_walks = chunk_walks;
    P.streams = streams; P.stream_chunks = ceil_div<int64_t>(nchunks, streams);
    P.shared_sm = (flags & GW_SGNS_SHARED_SM) ? 1 : 0;
    // per-pair reductions in gensim's order when sequential or when asked for
    // Holding an update back for two steps adds to the Hogwild staleness of the row in proportion to
    // workers / rows: measured neutral at 3157 workers on 38000 rows (KS 0.004 either way against a
    // 512-worker run), visible at 575 workers on 2300 rows (KS against the sequential gensim-faithful
    // oracle 0.050 vs 0.032).  Default: merged only where it is neutral.
    P.merge = (flags & GW_SGNS_MERGED) ? 1 : (flags & GW_SGNS_PER_PAIR) ? 0
              : (lanes > 1 && lanes * 12 <= (int64_t)n ? 1 : 0);
    P.job_queue = d_queue.as<unsigned long long>();
    P.alias = alias_table; P.exp_table = d_exp.as<float>();
    P.syn0 = syn0; P.syn1neg = syn1neg; P.key0 = (uint32_t)seed; P.key1 = (uint32_t)(seed >> 32);
    for (int ep = epoch_begin; ep < epoch_end; ++ep) {
        P.epoch = ep;
        int rc;
        if (d == 8) rc = (K == 5) ? launch_train<8, 4, 5>(P, lanes, s) : launch_train<8, 4, 0>(P, lanes, s);
        else if (d == 16) rc = (K == 5) ? launch_train<16, 8, 5>(P, lanes, s) : launch_train<16, 8, 0>(P, lanes, s);
        else if (d == 32) rc = (K == 5) ? launch_train<32, 16, 5>(P, lanes, s) : launch_train<32, 16, 0>(P, lanes, s);
        else if (d == 64) rc = (K == 5) ? launch_train<64, 32, 5>(P, lanes, s) : launch_train<64, 32, 0>(P, lanes, s);
        else rc = (K == 5) ? launch_train<128, 32, 5>(P, lanes, s) : launch_train<128, 32, 0>(P, lanes, s);
        GW_TRY(rc);
    }
    return GW_OK;
}
