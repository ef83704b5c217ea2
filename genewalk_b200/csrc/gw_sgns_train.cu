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
constexpr int kTrainThreads = 128;      // CTA size (launch bounds)

struct TrainParams {
    const int32_t *tokens;   // [L][W] position-major, or nullptr: the walks are regenerated
    const int32_t *row_ptr;  // graph of the regenerated walks
    const int32_t *col_idx;
    int32_t niter;
    uint32_t wkey0, wkey1;   // Philox key of the walks (gw_walk_uniform's seed)
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
    const gw_sgns_unit *units;   // the queue's work units, in the order they are handed out
    const int64_t *num_units;
    int64_t lanes;          // number of workers (groups beyond it, padding of the last warp, exit)
    int32_t merge;          // merged row reductions, see step_update_merged
    unsigned long long *job_queue;  // next unit to hand out (device counter, zeroed per launch)
    const gw_alias_entry *alias;
    const float *exp_table;  // [1000] device copy of gensim's EXP_TABLE
    float *syn0;
    float *syn1neg;
    uint32_t key0, key1;
};

// gw_sgns_auto_shape (profiles/parity_study_r02.md): workers ~ n/24, units of
// ~0.36 n^0.75 consecutive walks, i.e. a window of walks in flight of 0.3-1 % of the corpus -- the
// shapes that score best against the sequential oracle on the 2.3k-, 8.5k- and 38k-node study
// networks (95 x 120, 354 x 320, 1536 x 980).  Networks with hubs (a start node owning >= 100k
// walks) keep units of >= 320 walks: shorter units put every worker on the hub's walks at once and
// its rows overshoot (max row norm 20-70 instead of 12-14).  Twice the workers with half the units
// (the same window) train a lone replicate twice as fast and score the same KS, but were seen to
// overshoot on a hub once in six runs (norm 55): not shipped.
constexpr int64_t kAutoLanesMax = 1536;
constexpr int64_t kAutoLanesMin = 16;
constexpr int64_t kAutoRowsPerLane = 24;
constexpr int64_t kAutoLargeN = 200000;
constexpr int64_t kAutoLargeRowsPerLane = 128;
constexpr int64_t kAutoLargeLanesMax = 8192;
constexpr int64_t kAutoUnitMin = 50;
constexpr int64_t kAutoUnitMax = 1000;
constexpr int64_t kAutoHubWalks = 100000;
constexpr int64_t kAutoHubUnitMin = 320;
constexpr int32_t kAutoPolicy = GW_PLAN_CHUNK;

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
template <int G, int KT> struct StepDraw {
    static constexpr int CP = (KT + 1) / 2;            // Philox calls per pair
    static constexpr int NC = 2 * CP;                  // calls per step
    static constexpr int R = (NC + G - 1) / G;         // calls per lane (rounds)
    int32_t slot[R][2];                                // this lane's draws: alias slot ...
    uint32_t coin[R][2];                               // ... its 16-bit coin ...
    int2 entry[R][2];                                  // ... and the alias entry (load in flight)

    __device__ __forceinline__ void issue(const TrainParams &P, int sub, uint32_t w_lo, uint32_t w_hi,
                                          uint32_t q_first)
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
            // the alias table is read-only for the whole kernel: non-coherent path, L1-cacheable
            if (used) entry[r][0] = __ldg(reinterpret_cast<const int2 *>(P.alias) + slot[r][0]);
            if (second) entry[r][1] = __ldg(reinterpret_cast<const int2 *>(P.alias) + slot[r][1]);
        }
    }

    __device__ __forceinline__ void resolve(int leader, unsigned gmask, int32_t (&tg)[2 * KT]) const
    {
        int32_t mine[R][2];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int j = 0; j < 2; ++j)
                mine[r][j] = coin[r][j] < coin_threshold(__int_as_float(entry[r][j].x)) ? slot[r][j]
                                                                                     : entry[r][j].y;
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

// ---- the training kernel --------------------------------------------------------------------
// A worker (lane group of G threads) pulls UNITS -- runs of consecutive walks, planned by
// gw_sgns_plan -- from a device queue in corpus order and trains each unit's walks one after the
// other.  Walks come from the token array or, when there is none (kRegen), are regenerated from
// their Philox counters exactly as gw_walk_uniform draws them: the G threads of the group each
// generate one of the next G walks into shared memory (9 dependent CSR steps in flight per thread),
// then the group trains those G walks in order.
template <int D, int G, int KT, bool kMerge, bool kRegen>
__global__ void __launch_bounds__(kTrainThreads)
sgns_train_kernel(const TrainParams P)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    float *s_exp = reinterpret_cast<float *>(s_raw);
    for (int i = threadIdx.x; i < kExpTableSize; i += blockDim.x) s_exp[i] = P.exp_table[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int sub = lane & (G - 1);
    const int leader = lane & ~(G - 1);
    const unsigned gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << leader;
    if (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G >= P.lanes) return;
    // regenerated walks of this group: [G walks][L tokens]
    int32_t *s_walks = reinterpret_cast<int32_t *>(s_raw + sizeof(float) * kExpTableSize) +
                       (size_t)(threadIdx.x - sub) * (size_t)P.L;
    const unsigned long long num_units = (unsigned long long)__ldg(P.num_units);
    for (;;) {
        // gensim's job queue: a worker takes the next unit of the corpus when it finishes one
        unsigned long long u = 0;
        if (sub == 0) u = atomicAdd(P.job_queue, 1ULL);
        u = __shfl_sync(gmask, u, leader);
        if (u >= num_units) break;
        const int4 raw = __ldg(reinterpret_cast<const int4 *>(P.units) + u);
        int64_t w = (int64_t)(((uint64_t)(uint32_t)raw.y << 32) | (uint32_t)raw.x);
        const int64_t last = w + raw.z;
        int32_t node = raw.w;   // start node of walk w (regenerated walks)
        float alpha = 0.0f;
        int64_t next_job = w;   // the learning rate is set at the first walk and at every job boundary
        while (w < last) {
            int nb = 1;
            if constexpr (kRegen) {
                nb = last - w < G ? (int)(last - w) : G;
                int32_t v = node;
                if (sub < nb) {
                    const uint64_t ww = (uint64_t)(w + sub);
                    const int32_t e = (int32_t)(ww / (uint64_t)P.niter);
                    while (e >= __ldg(P.row_ptr + v + 1)) ++v;   // start node: owner of adjacency entry e
                    const int32_t v0 = v;
                    int32_t *out = s_walks + sub * P.L;
                    out[0] = v;
                    u32x4 rnd = {0, 0, 0, 0};
                    for (int32_t st = 1; st < P.L; ++st) {
                        const int sel = (st - 1) & 3;
                        if (sel == 0)
                            rnd = philox4x32_10((uint32_t)ww, (uint32_t)(ww >> 32), (uint32_t)((st - 1) >> 2),
                                                kStreamWalk, P.wkey0, P.wkey1);
                        const uint32_t word = sel == 0 ? rnd.x : sel == 1 ? rnd.y : sel == 2 ? rnd.z : rnd.w;
                        const int32_t b = __ldg(P.row_ptr + v);
                        const uint32_t deg = (uint32_t)(__ldg(P.row_ptr + v + 1) - b);
                        v = __ldg(P.col_idx + b + (int32_t)__umulhi(word, deg));
                        out[st] = v;
                    }
                    v = v0;
                }
                node = __shfl_sync(gmask, v, leader + nb - 1);
                __syncwarp(gmask);
            }
            for (int j = 0; j < nb; ++j, ++w) {
                if (w >= next_job) {
                    // gensim _get_next_alpha at pushed = job*job_walks examples (fp64, unfused, as
                    // the oracle)
                    const int64_t pushed = (w / P.job_walks) * P.job_walks;
                    const double epoch_progress = __ddiv_rn((double)pushed, (double)P.W);
                    const double progress = __ddiv_rn(__dadd_rn((double)P.epoch, epoch_progress), (double)P.epochs);
                    double next_alpha = __dsub_rn(P.alpha0, __dmul_rn(P.alpha_span, progress));
                    if (next_alpha < P.alpha_min) next_alpha = P.alpha_min;
                    alpha = (float)next_alpha;
                    next_job = pushed + P.job_walks;
                }
                const int32_t *s_walk = s_walks + j * P.L;
                auto token = [&](int32_t i) -> int32_t {
                    if constexpr (kRegen) return s_walk[i];
                    else return __ldg(P.tokens + (int64_t)i * P.W + w);
                };
                const uint32_t w_lo = (uint32_t)(uint64_t)w, w_hi = (uint32_t)((uint64_t)w >> 32);
                int32_t prev = -1, cur = token(0);
                int32_t next = P.L > 1 ? token(1) : -1;
                if constexpr (KT > 0) {
                    StepDraw<G, KT> draw;
                    draw.issue(P, sub, w_lo, w_hi, 0u);
                    CtxCarry<D / G> carry;
                    carry.xn = carry.yn = -1;
                    for (int32_t i = 0; i < P.L && cur >= 0; ++i) {
                        // token two positions ahead (the next step's right context), loaded early
                        const int32_t next2 = (i + 2 < P.L && next >= 0) ? token(i + 2) : -1;
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
                        step_gather<D, G, KT>(P, sub, cur, prev, next, tg, rows);
                        // the next step's draws (Philox + alias loads) overlap this step's row gathers
                        draw.issue(P, sub, w_lo, w_hi, 2u * (uint32_t)(i + 1));
                        if constexpr (kMerge)
                            step_update_merged<D, G, KT>(P, s_exp, sub, gmask, cur, prev, next, alpha, tg,
                                                         rows, carry);
                        else
                            step_update<D, G, KT>(P, s_exp, sub, gmask, cur, prev, next, alpha, tg, rows);
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
                        next = (i + 2 < P.L && cur >= 0) ? token(i + 2) : -1;
                    }
                }
            }
            if constexpr (kRegen) __syncwarp(gmask);   // the group's walks are consumed: slots free
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
    // `lanes` workers of G threads each in CTAs of 128 threads: the trainings of other replicates,
    // launched on other streams, share the SMs (independent replicates have independent tables, so
    // training several at once is what fills the GPU without adding workers per table)
    const bool regen = P.tokens == nullptr;
    int block = 128;
    if (lanes * G < block) block = (int)(ceil_div<int64_t>(lanes * G, 32) * 32);
    const int blocks = (int)ceil_div<int64_t>(lanes * G, block);
    const size_t smem = sizeof(float) * kExpTableSize + (regen ? sizeof(int32_t) * (size_t)block * (size_t)P.L : 0);
    GW_REQUIRE(smem <= 200 * 1024, "gw_sgns_train: walk_length=%d is too long to regenerate walks in the trainer; "
               "pass the token array", P.L);
    GW_CUDA_OK(cudaMemsetAsync(P.job_queue, 0, sizeof(unsigned long long), s));
    constexpr bool kCanMerge = KT > 0;
    const bool merge = kCanMerge && P.merge;
    void (*kern)(const TrainParams) =
        regen ? (merge ? sgns_train_kernel<D, G, KT, kCanMerge, true> : sgns_train_kernel<D, G, KT, false, true>)
              : (merge ? sgns_train_kernel<D, G, KT, kCanMerge, false> : sgns_train_kernel<D, G, KT, false, false>);
    if (smem > 48 * 1024)
        GW_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<blocks, block, smem, s>>>(P);
    GW_LAUNCH_OK();
    return GW_OK;
}

// ---- unit planner ------------------------------------------------------------------------------
__global__ void plan_chunk_kernel(int32_t n, const int32_t *__restrict__ row_ptr, int32_t niter, int64_t W,
                                  int64_t unit_walks, int64_t nunits, gw_sgns_unit *__restrict__ units,
                                  int64_t *__restrict__ num_units)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *num_units = nunits;
    if (i >= nunits) return;
    const int64_t first = i * unit_walks;
    gw_sgns_unit u;
    u.first = first;
    u.count = (int32_t)(first + unit_walks <= W ? unit_walks : W - first);
    u.node = -1;
    if (row_ptr) {   // owner of adjacency entry first / niter: last v with row_ptr[v] <= e
        const int32_t e = (int32_t)(first / niter);
        int32_t lo = 0, hi = n;   // invariant: row_ptr[lo] <= e < row_ptr[hi]
        while (hi - lo > 1) {
            const int32_t mid = (int32_t)(((uint32_t)lo + (uint32_t)hi) >> 1);
            if (row_ptr[mid] <= e) lo = mid; else hi = mid;
        }
        u.node = lo;
    }
    units[i] = u;
}

__global__ void plan_burst_count_kernel(int32_t n, const int32_t *__restrict__ row_ptr, int32_t niter,
                                        int64_t unit_walks, int64_t max_pieces, int64_t *__restrict__ pieces)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const int64_t len = (int64_t)(row_ptr[v + 1] - row_ptr[v]) * niter;
    int64_t k = len == 0 ? 0 : ceil_div<int64_t>(len, unit_walks);
    if (max_pieces > 0 && k > max_pieces) k = max_pieces;   // a hub's walks: at most max_pieces workers
    pieces[v] = k;
}

__global__ void plan_burst_fill_kernel(int32_t n, const int32_t *__restrict__ row_ptr, int32_t niter,
                                       const int64_t *__restrict__ offset, const int64_t *__restrict__ total,
                                       int64_t capacity, gw_sgns_unit *__restrict__ units,
                                       int64_t *__restrict__ num_units)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v == 0) *num_units = *total <= capacity ? *total : capacity;
    if (v >= n) return;
    const int64_t begin = (int64_t)row_ptr[v] * niter, len = (int64_t)(row_ptr[v + 1] - row_ptr[v]) * niter;
    const int64_t o = offset[v];
    const int64_t k = (v + 1 < n ? offset[v + 1] : *total) - o;
    // k near-equal pieces: piece j covers [begin + j*len/k, begin + (j+1)*len/k)
    for (int64_t j = 0; j < k; ++j) {
        if (o + j >= capacity) return;
        const int64_t a = begin + (len * j) / k, b = begin + (len * (j + 1)) / k;
        gw_sgns_unit u;
        u.first = a;
        u.count = (int32_t)(b - a);
        u.node = v;
        units[o + j] = u;
    }
}

}  // namespace gw

extern "C" int64_t gw_sgns_plan_capacity(int32_t n, int64_t W, int64_t unit_walks, int32_t policy)
{
    if (unit_walks < 1) unit_walks = 1;
    const int64_t chunks = (W + unit_walks - 1) / unit_walks;
    return policy == GW_PLAN_BURST ? chunks + (int64_t)n : chunks;
}

extern "C" int gw_sgns_plan(int32_t n, const int32_t *row_ptr, int32_t niter, int64_t W, int64_t unit_walks,
                            int32_t policy, int64_t max_pieces, gw_sgns_unit *units, int64_t capacity,
                            int64_t *num_units, gw_stream_t stream)
{
    using namespace gw;
    cudaStream_t s = (cudaStream_t)stream;
    GW_REQUIRE(units && num_units && W > 0 && unit_walks >= 1 && unit_walks < ((int64_t)1 << 31) && max_pieces >= 0,
               "gw_sgns_plan: bad arguments");
    GW_REQUIRE(max_pieces == 0 || W / max_pieces < ((int64_t)1 << 31), "gw_sgns_plan: max_pieces too small for W");
    GW_REQUIRE(capacity >= gw_sgns_plan_capacity(n, W, unit_walks, policy),
               "gw_sgns_plan: capacity %lld < gw_sgns_plan_capacity()", (long long)capacity);
    if (policy == GW_PLAN_CHUNK) {
        const int64_t nunits = ceil_div<int64_t>(W, unit_walks);
        plan_chunk_kernel<<<(unsigned)ceil_div<int64_t>(nunits, 256), 256, 0, s>>>(n, row_ptr, niter, W, unit_walks,
                                                                                 nunits, units, num_units);
        GW_LAUNCH_OK();
        return GW_OK;
    }
    GW_REQUIRE(policy == GW_PLAN_BURST, "gw_sgns_plan: unknown policy %d", policy);
    GW_REQUIRE(row_ptr && n > 0 && niter > 0, "gw_sgns_plan: the burst policy needs the graph");
    Temp pieces, total;
    GW_CUDA_OK(pieces.alloc(sizeof(int64_t) * (size_t)n, s));
    GW_CUDA_OK(total.alloc(sizeof(int64_t), s));
    const unsigned grid = (unsigned)ceil_div<int64_t>(n, 256);
    plan_burst_count_kernel<<<grid, 256, 0, s>>>(n, row_ptr, niter, unit_walks, max_pieces, pieces.as<int64_t>());
    GW_LAUNCH_OK();
    GW_TRY(scan_exclusive_i64(pieces.as<int64_t>(), pieces.as<int64_t>(), n, total.as<int64_t>(), s));
    plan_burst_fill_kernel<<<grid, 256, 0, s>>>(n, row_ptr, niter, pieces.as<int64_t>(), total.as<int64_t>(),
                                                capacity, units, num_units);
    GW_LAUNCH_OK();
    return GW_OK;
}

// Default parallel shape: a function of the network alone -- never of how
// many trainings share the GPU or of how many GPUs the replicates are spread over, so that a seed
// gives the same statistics on 1 and on 8 GPUs.
extern "C" int gw_sgns_auto_shape(int32_t n, int64_t W, int64_t job_walks, int64_t max_start_walks,
                                  int64_t *lanes_out, int64_t *unit_walks_out, int32_t *policy_out)
{
    using namespace gw;
    GW_REQUIRE(n > 0 && W > 0 && job_walks > 0 && lanes_out && unit_walks_out && policy_out,
               "gw_sgns_auto_shape: bad arguments");
    int64_t lanes = n / kAutoRowsPerLane;
    if (lanes > kAutoLanesMax) lanes = kAutoLanesMax;
    // beyond the validated range (n <= 38k): one worker per 128 rows, so that the trainings that
    // fill the GPU are few enough for their tables (64 B per row and training) to stay in L2.
    // Extrapolated from the window rule (workers x unit ~ 0.75 % of the corpus); no oracle exists
    // at this size (28 CPU-hours per replicate).
    if (n >= kAutoLargeN && (int64_t)n / kAutoLargeRowsPerLane > lanes) {
        lanes = (int64_t)n / kAutoLargeRowsPerLane;
        if (lanes > kAutoLargeLanesMax) lanes = kAutoLargeLanesMax;
    }
    if (lanes < kAutoLanesMin) lanes = kAutoLanesMin;
    if (lanes > n / 2) lanes = n / 2 > 0 ? n / 2 : 1;
    int64_t unit = (int64_t)(0.36 * pow((double)n, 0.75) + 0.5);
    if (unit > kAutoUnitMax) unit = kAutoUnitMax;
    if (unit < kAutoUnitMin) unit = kAutoUnitMin;
    // max_start_walks = niter * (largest number of distinct neighbours); < 0: unknown, assume hubs
    // on any corpus large enough to have them
    const bool hubs = max_start_walks >= 0 ? max_start_walks >= kAutoHubWalks : W >= 100 * kAutoHubWalks;
    if (hubs && unit < kAutoHubUnitMin) unit = kAutoHubUnitMin;
    if (unit > W) unit = W;
    const int64_t nunits = ceil_div<int64_t>(W, unit);
    if (lanes > nunits) lanes = nunits;
    *lanes_out = lanes;
    *unit_walks_out = unit;
    *policy_out = kAutoPolicy;
    return GW_OK;
}

extern "C" int gw_sgns_resident_trainings(int32_t d, int64_t lanes)
{
    using namespace gw;
    if (lanes < 1) lanes = 1;
    const int G = d / 2 > 32 ? 32 : (d / 2 < 1 ? 1 : d / 2);
    // the kernels need <= 128 registers per thread: 512 threads (4 CTAs of 128) per SM
    const int64_t threads = (int64_t)num_sms() * 512;
    const int64_t fit = threads / (lanes * G);
    return (int)(fit < 1 ? 1 : fit > 1024 ? 1024 : fit);
}

extern "C" int gw_sgns_train(int32_t n, int32_t d, const int32_t *tokens, const int32_t *row_ptr,
                             const int32_t *col_idx, int32_t niter, uint64_t walk_seed, int64_t W, int32_t L,
                             int32_t window, int32_t K, int32_t epochs, int32_t epoch_begin,
                             int32_t epoch_end, double alpha0, double alpha_min, int64_t job_walks,
                             const gw_sgns_unit *units, const int64_t *num_units,
                             const gw_alias_entry *alias_table, float *syn0, float *syn1neg,
                             uint64_t seed, int64_t lanes, int32_t flags, gw_stream_t stream)
{
    using namespace gw;
    cudaStream_t s = (cudaStream_t)stream;
    GW_REQUIRE(n > 0 && alias_table && syn0 && syn1neg && units && num_units, "gw_sgns_train: null argument");
    GW_REQUIRE(tokens || (row_ptr && col_idx && niter > 0),
               "gw_sgns_train: pass the token array, or the graph (row_ptr, col_idx, niter) to regenerate the walks");
    GW_REQUIRE(W > 0 && L >= 1 && L <= 4096, "gw_sgns_train: W=%lld L=%d out of range", (long long)W, L);
    GW_REQUIRE(K >= 0 && K <= 32, "gw_sgns_train: negative=%d out of range [0,32]", K);
    GW_REQUIRE(epochs >= 1 && epochs <= 65535 && epoch_begin >= 0 && epoch_begin <= epoch_end &&
                   epoch_end <= epochs,
               "gw_sgns_train: epochs=%d [%d,%d) invalid", epochs, epoch_begin, epoch_end);
    GW_REQUIRE(job_walks >= 1 && lanes >= 1, "gw_sgns_train: job_walks/lanes invalid");
    if (window != 1) {
        set_error("gw_sgns_train: window=%d is not implemented (GeneWalk uses window=1)", window);
        return GW_ERR_UNSUPPORTED;
    }
    if (d != 8 && d != 16 && d != 32 && d != 64 && d != 128) {
        set_error("gw_sgns_train: vector_size=%d is not implemented (8, 16, 32, 64, 128)", d);
        return GW_ERR_UNSUPPORTED;
    }
    if ((flags & GW_SGNS_MERGED) && K != 5) {
        set_error("gw_sgns_train: GW_SGNS_MERGED is implemented for negative=5 only (got %d)", K);
        return GW_ERR_UNSUPPORTED;
    }
    GW_REQUIRE((reinterpret_cast<uintptr_t>(syn0) & 15) == 0 && (reinterpret_cast<uintptr_t>(syn1neg) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(units) & 15) == 0,
               "gw_sgns_train: tables and units must be 16-byte aligned");
    float h_exp[kExpTableSize];
    host_exp_table(h_exp);
    Temp d_exp, d_queue;
    GW_CUDA_OK(d_exp.alloc(sizeof h_exp, s));
    GW_CUDA_OK(d_queue.alloc(sizeof(unsigned long long), s));
    // pageable source: the copy is staged before the call returns, so the stack buffer is safe
    GW_CUDA_OK(cudaMemcpyAsync(d_exp.ptr, h_exp, sizeof h_exp, cudaMemcpyHostToDevice, s));
    TrainParams P;
    P.tokens = tokens; P.row_ptr = row_ptr; P.col_idx = col_idx; P.niter = niter;
    P.wkey0 = (uint32_t)walk_seed; P.wkey1 = (uint32_t)(walk_seed >> 32);
    P.W = W; P.L = L; P.n = n; P.K = K; P.epochs = epochs; P.epoch = 0;
    P.alpha0 = alpha0; P.alpha_span = alpha0 - alpha_min; P.alpha_min = alpha_min;
    P.job_walks = job_walks; P.units = units; P.num_units = num_units; P.lanes = lanes;
    // per-pair reductions in gensim's order unless the merged order is asked for
    P.merge = (flags & GW_SGNS_MERGED) ? 1 : 0;
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
