/*
 * gw_oracle.c -- CPU ORACLE for the GeneWalk representation-learning hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke() check in
 * __graft_entry__.py and the cpu_baseline / --impl reference legs of bench.py may load it.
 * The product (genewalk_b200/) never imports, links or executes anything under oracle/.
 *
 * What it restates (reference = /root/reference, churchmanlab/genewalk v1.6.3):
 *
 *   gwo_walk_uniform     genewalk/deepwalk.py:145-200  (run_single_walk / run_walks_for_node /
 *                        get_start_nodes): node v starts niter*len(graph[v]) walks, every step is a
 *                        uniform choice among the DISTINCT neighbours of the current node.  The
 *                        reference draws with CPython's MT19937 (random.choice); north_star replaces
 *                        that by a counter-based Philox4x32-10 sampler, and this function is the CPU
 *                        statement of exactly that sampler (the GPU kernel must match it bit for bit).
 *
 *   gwo_sgns_train       genewalk/deepwalk.py:135-139 -> gensim.models.Word2Vec(sg=1, negative=K,
 *                        window=1, sample=0, hs=0).  gensim (requirement `gensim>=4.0.0`,
 *                        /root/reference/setup.py:11) is a third-party dependency that is NOT present
 *                        in /root/reference nor installed in the authoring container, so this is a
 *                        restatement of gensim 4.x's published algorithm (models/word2vec.py
 *                        _job_producer/_get_next_alpha/make_cum_table, models/word2vec_inner.pyx
 *                        w2v_fast_sentence_sg_neg/train_batch_sg), see SURVEY.md section 8(c).
 *                        PARITY UNPINNED against real gensim: the reference's tests hold no numeric
 *                        vector for this step (test_deepwalk.py stops before word2vec) and gensim
 *                        cannot be run here.
 *                        Two samplers for the negative draws:
 *                          sampler 0 "gensim":  cum_table bisect_left driven by the 48-bit LCG
 *                          sampler 1 "device":  Philox4x32-10 + alias table -- the sampler the CUDA
 *                                               kernel uses, so that a 1-thread CUDA run is bit-exact
 *                                               against this function.
 *                        Two reduction orders: gensim's per-pair order (merged = 0) and the order
 *                        of the kernel's Hogwild mode (merged = 1, see walk_merged); both are
 *                        pinned bit for bit by the independent pure-Python statement in
 *                        oracle/sgns_py.py (tests/test_oracle.py).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off, so no FMA contraction: every float op
 * rounds separately, which is what the CUDA kernel does with __fmul_rn/__fadd_rn).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define GWO_STREAM_WALK 0x57u /* 'W': ctr[3] of the walk stream */
#define GWO_STREAM_SGNS 0x53u /* 'S': low byte of ctr[3] of the SGNS stream; epoch sits above it */

/* ------------------------------------------------------------------------------------------
 * Philox4x32-10 (Salmon et al., SC'11 "Parallel random numbers: as easy as 1, 2, 3").
 * ------------------------------------------------------------------------------------------ */
void gwo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* ------------------------------------------------------------------------------------------
 * Walks.  Graph = CSR over distinct neighbours in networkx adjacency order:
 *   col_idx[row_ptr[v] + k] == list(graph[v])[k]          (deepwalk.py:165)
 * A = row_ptr[n] adjacency entries, W = niter * A walks     (deepwalk.py:173,197)
 * Canonical walk id  w = e * niter + it   (e = adjacency entry, it = iteration) is the position
 * of the walk in the reference's serial corpus (deepwalk.py:67-70): node-major, niter*deg each.
 * Storage slot p -> (it = p / A, r = p % A, e = (r * stride) % A), stride coprime to A; stride 1
 * with it-major order is the plain interleaving; stride 0 stores walk w in slot w (corpus order).  tokens are position-major: tokens[i*W + p].
 * Step s (1..L-1) of walk w consumes word (s-1)&3 of
 *   Philox(ctr = {lo32(w), hi32(w), (s-1)>>2, 0x57}, key = {lo32(seed), hi32(seed)})
 * and moves to col_idx[row_ptr[v] + mulhi32(word, deg(v))].
 * ------------------------------------------------------------------------------------------ */
int gwo_walk_uniform(int32_t n, const int32_t *row_ptr, const int32_t *col_idx, int32_t niter,
                     int32_t L, uint64_t seed, uint64_t stride, int64_t slot_begin,
                     int64_t slot_end, int64_t W_stride_out, int32_t *tokens_out)
{
    /* writes slot p (slot_begin <= p < slot_end) to tokens_out[i*W_stride_out + (p - slot_begin)] */
    const int64_t A = row_ptr[n];
    if (A <= 0 || niter <= 0 || L <= 0) return -1;
    /* src[e] = row that owns adjacency entry e */
    int32_t *src = (int32_t *)malloc(sizeof(int32_t) * (size_t)A);
    if (!src) return -2;
    for (int32_t v = 0; v < n; ++v)
        for (int32_t k = row_ptr[v]; k < row_ptr[v + 1]; ++k) src[k] = v;
    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    for (int64_t p = slot_begin; p < slot_end; ++p) {
        int64_t e;
        uint64_t w;
        if (stride == 0) { /* canonical order: slot == walk id == position in the reference corpus */
            w = (uint64_t)p;
            e = p / niter;
        } else {
            const int64_t it = p / A, r = p % A;
            e = (int64_t)(((unsigned __int128)(uint64_t)r * stride) % (uint64_t)A);
            w = (uint64_t)e * (uint64_t)niter + (uint64_t)it;
        }
        int32_t v = src[e];
        int64_t o = p - slot_begin;
        tokens_out[o] = v;
        uint32_t rnd[4] = {0, 0, 0, 0};
        for (int32_t s = 1; s < L; ++s) {
            if (((s - 1) & 3) == 0) {
                uint32_t ctr[4] = {(uint32_t)w, (uint32_t)(w >> 32), (uint32_t)((s - 1) >> 2),
                                   GWO_STREAM_WALK};
                gwo_philox4x32_10(ctr, key, rnd);
            }
            const uint32_t deg = (uint32_t)(row_ptr[v + 1] - row_ptr[v]);
            const uint32_t k = (uint32_t)(((uint64_t)rnd[(s - 1) & 3] * deg) >> 32);
            v = col_idx[row_ptr[v] + (int32_t)k];
            tokens_out[(int64_t)s * W_stride_out + o] = v;
        }
    }
    free(src);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * SGNS.
 * ------------------------------------------------------------------------------------------ */
#define GWO_EXP_TABLE_SIZE 1000
#define GWO_MAX_EXP 6

/* gensim word2vec_inner.pyx init(): EXP_TABLE[i] = e/(e+1), e = exp((i/1000*2-1)*6), REAL_t math */
void gwo_exp_table(float *table /* [1000] */)
{
    for (int i = 0; i < GWO_EXP_TABLE_SIZE; ++i) {
        float x = ((float)i / (float)GWO_EXP_TABLE_SIZE * 2.0f - 1.0f) * (float)GWO_MAX_EXP;
        float e = (float)exp((double)x);
        table[i] = e / (e + 1.0f);
    }
}

/* gensim _get_next_alpha (word2vec.py): evaluated with pushed = job*job_walks examples */
static float job_alpha(double alpha0, double alpha_min, int epoch, int epochs, int64_t pushed,
                       int64_t W)
{
    double epoch_progress = 1.0 * (double)pushed / (double)W;
    double progress = ((double)epoch + epoch_progress) / (double)epochs;
    double next_alpha = alpha0 - (alpha0 - alpha_min) * progress;
    if (next_alpha < alpha_min) next_alpha = alpha_min;
    return (float)next_alpha;
}

typedef struct {
    /* sampler 0 */
    const uint32_t *cum_table;
    int32_t cum_len;
    uint64_t next_random;
    /* sampler 1 */
    const float *alias_prob;
    const int32_t *alias_idx;
    uint32_t key[2];
    uint32_t ctr_p_lo, ctr_p_hi, ctr_base, ctr_tag; /* ctr_base = q<<4, ctr_tag = (epoch<<8)|'S' */
    uint32_t rnd[4];
    int n_drawn;
} draw_state;

static inline int32_t bisect_left_u32(const uint32_t *a, int32_t lo, int32_t hi, uint32_t x)
{
    while (hi > lo) {
        int32_t mid = (int32_t)(((uint32_t)lo + (uint32_t)hi) >> 1);
        if (a[mid] >= x) hi = mid; else lo = mid + 1;
    }
    return lo;
}

/* Dot product in the summation order of the CUDA kernel (gensim calls BLAS sdot, whose order is
 * implementation-defined): the row is split into G = min(32, d/2) chunks (d a power of two >= 4;
 * otherwise one chunk), each chunk summed left to right, chunk sums combined by a pairwise tree. */
static float dot_rows(const float *a, const float *b, int d)
{
    int G = 1;
    if (d >= 4 && (d & (d - 1)) == 0) G = d / 2 > 32 ? 32 : d / 2;
    const int E = d / G;
    float c[32];
    for (int g = 0; g < G; ++g) {
        float s = a[g * E] * b[g * E];
        for (int k = 1; k < E; ++k) s = s + a[g * E + k] * b[g * E + k];
        c[g] = s;
    }
    for (int m = 1; m < G; m <<= 1)
        for (int g = 0; g < G; g += 2 * m) c[g] = c[g] + c[g + m];
    return c[0];
}

/* negative t (1..K) of the current pair */
static int32_t draw_target(int sampler, draw_state *ds, int32_t V, int t)
{
    if (sampler == 0) {
        uint32_t x = (uint32_t)((ds->next_random >> 16) % ds->cum_table[ds->cum_len - 1]);
        int32_t target = bisect_left_u32(ds->cum_table, 0, ds->cum_len, x);
        ds->next_random = (ds->next_random * 25214903917ULL + 11ULL) & 281474976710655ULL;
        return target;
    }
    const int w = 2 * (t - 1); /* words w, w+1 of the pair's Philox stream */
    if ((w >> 2) != ds->n_drawn) {
        uint32_t ctr[4] = {ds->ctr_p_lo, ds->ctr_p_hi, ds->ctr_base | (uint32_t)(w >> 2), ds->ctr_tag};
        gwo_philox4x32_10(ctr, ds->key, ds->rnd);
        ds->n_drawn = w >> 2;
    }
    const uint32_t x0 = ds->rnd[w & 3], x1 = ds->rnd[(w & 3) + 1];
    const int32_t slot = (int32_t)(((uint64_t)x0 * (uint32_t)V) >> 32);
    /* 16-bit coin against the slot's acceptance threshold floor(prob * 2^16) (<= 65535) */
    uint32_t thr = (uint32_t)floorf(ds->alias_prob[slot] * 65536.0f);
    if (thr > 65535u) thr = 65535u;
    return ((x1 >> 16) < thr) ? slot : ds->alias_idx[slot];
}

/* sigmoid through gensim's EXP_TABLE; returns 0 when f is saturated (the target is skipped).
 * For the largest floats below 6, f + 6 rounds up to 12.0f and gensim's index becomes 1000, one past
 * its table (it then reads the adjacent static array); clamped to the last entry. */

/* EXP_TABLE index scale: gensim writes (f + MAX_EXP) * (EXP_TABLE_SIZE / MAX_EXP / 2) with Cython DEF
 * constants 1000 and 6.  Whether the compiled extension folds that to 83.333... (true division) or
 * to 83 (C-style integer division, as in the original word2vec.c) depends on how the .pyx was
 * cythonized; it cannot be read off the published source.  Default: true division.  The pin script
 * (tools/pin_gensim.py) runs real gensim against both and reports which one it is. */
static double g_exp_scale = (double)GWO_EXP_TABLE_SIZE / (double)GWO_MAX_EXP / 2.0;
void gwo_set_exp_scale(double scale) { g_exp_scale = scale; }
double gwo_get_exp_scale(void) { return g_exp_scale; }

static int table_sigmoid(float f, const float *exp_table, float *sig)
{
    if (f <= -(float)GWO_MAX_EXP || f >= (float)GWO_MAX_EXP) return 0;
    int ei = (int)((double)(f + (float)GWO_MAX_EXP) * g_exp_scale);
    if (ei >= GWO_EXP_TABLE_SIZE) ei = GWO_EXP_TABLE_SIZE - 1;
    *sig = exp_table[ei];
    return 1;
}

/* one (centre, context) pair: gensim w2v_fast_sentence_sg_neg (word2vec_inner.pyx) */
static void pair_update(int sampler, draw_state *ds, int32_t V, int32_t d, int32_t K,
                        int32_t centre, int32_t context, float alpha, const float *exp_table,
                        float *syn0, float *syn1neg, float *work)
{
    float *row1 = syn0 + (int64_t)context * d;
    for (int k = 0; k < d; ++k) work[k] = 0.0f;
    for (int t = 0; t < K + 1; ++t) {
        int32_t target;
        float label;
        if (t == 0) {
            target = centre;
            label = 1.0f;
        } else {
            target = draw_target(sampler, ds, V, t);
            if (target == centre) continue;
            label = 0.0f;
        }
        float *row2 = syn1neg + (int64_t)target * d;
        const float f = dot_rows(row1, row2, d);
        float sig;
        if (!table_sigmoid(f, exp_table, &sig)) continue;
        const float g = (label - sig) * alpha;
        for (int k = 0; k < d; ++k) work[k] = work[k] + g * row2[k];
        for (int k = 0; k < d; ++k) row2[k] = row2[k] + g * row1[k];
    }
    for (int k = 0; k < d; ++k) row1[k] = row1[k] + work[k]; /* lockf == 1 */
}

/* ---- merged reduction order (the CUDA kernel's Hogwild default, GW_SGNS_MERGED) ------------------
 * Same update rule, same pairs, same negatives; only the moment at which a worker's updates of the
 * rows it visits twice reach the tables differs: the two updates of syn1neg[centre] inside a step
 * (a step = both pairs of a centre) are summed and added once at the end of the step, and the two
 * updates of syn0[walk[j]] -- from step j-1 (as right context) and step j+1 (as left context) -- are
 * summed and added once in step j+1; in step j+1 the worker reads the row from the table and adds
 * its own held-back update, so it computes on an up-to-date row.
 * This restates genewalk_b200/csrc/gw_sgns_train.cu:step_update_merged operation by operation; it
 * is NOT gensim's order (from which it differs by float summation order only). */
static void pair_merged(int sampler, draw_state *ds, int32_t V, int32_t d, int32_t K, int32_t centre,
                        float *ctx, float *rc, float *dC, int *anyC, float alpha,
                        const float *exp_table, float *syn1neg, float *work)
{
    for (int k = 0; k < d; ++k) work[k] = 0.0f;
    for (int t = 0; t < K + 1; ++t) {
        float sig;
        if (t == 0) {
            const float f = dot_rows(ctx, rc, d);
            if (!table_sigmoid(f, exp_table, &sig)) continue;
            const float g = (1.0f - sig) * alpha;
            for (int k = 0; k < d; ++k) {
                work[k] = work[k] + g * rc[k];
                const float delta = g * ctx[k];
                rc[k] = rc[k] + delta;
                dC[k] = dC[k] + delta;
            }
            *anyC = 1;
        } else {
            const int32_t target = draw_target(sampler, ds, V, t);
            if (target == centre) continue;
            float *row2 = syn1neg + (int64_t)target * d;
            const float f = dot_rows(ctx, row2, d);
            if (!table_sigmoid(f, exp_table, &sig)) continue;
            const float g = (0.0f - sig) * alpha;
            for (int k = 0; k < d; ++k) work[k] = work[k] + g * row2[k];
            for (int k = 0; k < d; ++k) row2[k] = row2[k] + g * ctx[k];
        }
    }
    for (int k = 0; k < d; ++k) ctx[k] = ctx[k] + work[k];
}

static void walk_merged(int sampler, draw_state *ds, int32_t V, int32_t d, int32_t K,
                        const int32_t *walk, int len, int64_t p, int ep, float alpha,
                        const float *exp_table, float *syn0, float *syn1neg)
{
    float pX[1024], pY[1024], ctxA[1024], ctxB[1024], rc[1024], dC[1024], work[1024];
    int32_t xn = -1, yn = -1; /* nodes whose update is pending (held back), -1 = none */
    for (int i = 0; i < len; ++i) {
        const int32_t c = walk[i], a = i > 0 ? walk[i - 1] : -1, b = i + 1 < len ? walk[i + 1] : -1;
        int anyC = 0;
        for (int k = 0; k < d; ++k) {
            rc[k] = syn1neg[(int64_t)c * d + k];
            dC[k] = 0.0f;
        }
        ds->ctr_p_lo = (uint32_t)(uint64_t)p;
        ds->ctr_p_hi = (uint32_t)((uint64_t)p >> 32);
        ds->ctr_tag = ((uint32_t)ep << 8) | GWO_STREAM_SGNS;
        if (a >= 0) {
            const int pending = xn == a;
            for (int k = 0; k < d; ++k) {
                ctxA[k] = syn0[(int64_t)a * d + k];
                if (pending) ctxA[k] = ctxA[k] + pX[k];
            }
            ds->ctr_base = (uint32_t)(2 * i) << 4;
            ds->n_drawn = -1;
            pair_merged(sampler, ds, V, d, K, c, ctxA, rc, dC, &anyC, alpha, exp_table, syn1neg, work);
            if (pending)
                for (int k = 0; k < d; ++k) work[k] = pX[k] + work[k];
            for (int k = 0; k < d; ++k) syn0[(int64_t)a * d + k] = syn0[(int64_t)a * d + k] + work[k];
        }
        memcpy(pX, pY, sizeof(float) * (size_t)d);
        xn = yn;
        yn = -1;
        if (b >= 0) {
            const int back = a >= 0 && b == a;
            for (int k = 0; k < d; ++k) ctxB[k] = back ? ctxA[k] : syn0[(int64_t)b * d + k];
            ds->ctr_base = (uint32_t)(2 * i + 1) << 4;
            ds->n_drawn = -1;
            pair_merged(sampler, ds, V, d, K, c, ctxB, rc, dC, &anyC, alpha, exp_table, syn1neg, pY);
            yn = b;
        }
        if (anyC)
            for (int k = 0; k < d; ++k) syn1neg[(int64_t)c * d + k] = syn1neg[(int64_t)c * d + k] + dC[k];
    }
    if (xn >= 0)
        for (int k = 0; k < d; ++k) syn0[(int64_t)xn * d + k] = syn0[(int64_t)xn * d + k] + pX[k];
    if (yn >= 0)
        for (int k = 0; k < d; ++k) syn0[(int64_t)yn * d + k] = syn0[(int64_t)yn * d + k] + pY[k];
}

/*
 * tokens: position-major [L*W] indices into the V rows of syn0/syn1neg.
 * sampler 0: cum_table[cum_len] (gensim make_cum_table), job_seeds[epochs*njobs] = next_random per
 *            job (gensim: 2^24*randint(0,2^24)+randint(0,2^24) from RandomState(seed)), njobs =
 *            ceil(W/job_walks).
 * sampler 1: alias_prob/alias_idx[V], seed; pair q=2i+(j>i) of walk slot p in epoch ep uses
 *            Philox ctr {lo32(p), hi32(p), (q<<4)|c, (ep<<8)|0x53}, c = call number (4 words each;
 *            negative t=1..K uses words 2(t-1), 2(t-1)+1 of the concatenated stream).
 * nthreads > 1: jobs are handed to OpenMP threads and race on the tables (Hogwild, as gensim's
 *            worker threads do); nthreads == 1 is the deterministic sequential statement.
 * merged != 0: the CUDA kernel's merged reduction order (walk_merged) instead of gensim's.
 */
int gwo_sgns_train(int sampler, int32_t V, int32_t d, const int32_t *tokens, int64_t W, int32_t L,
                   int32_t K, int32_t epochs, double alpha0, double alpha_min, int64_t job_walks,
                   const uint32_t *cum_table, int32_t cum_len, const uint64_t *job_seeds,
                   const float *alias_prob, const int32_t *alias_idx, uint64_t seed, float *syn0,
                   float *syn1neg, int nthreads, int merged)
{
    if (V <= 0 || d <= 0 || d > 1024 || W <= 0 || L < 1 || L > 4096 || K < 0 || K > 32 ||
        epochs < 1 || epochs > 65535 || job_walks < 1)
        return -1;
    float exp_table[GWO_EXP_TABLE_SIZE];
    gwo_exp_table(exp_table);
    const int64_t njobs = (W + job_walks - 1) / job_walks;
    for (int ep = 0; ep < epochs; ++ep) {
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads > 1 ? nthreads : 1)
#endif
        for (int64_t job = 0; job < njobs; ++job) {
            float work[1024];
            int32_t walk[4096];
            draw_state ds;
            memset(&ds, 0, sizeof ds);
            ds.cum_table = cum_table;
            ds.cum_len = cum_len;
            ds.alias_prob = alias_prob;
            ds.alias_idx = alias_idx;
            ds.key[0] = (uint32_t)seed;
            ds.key[1] = (uint32_t)(seed >> 32);
            if (sampler == 0) ds.next_random = job_seeds[(int64_t)ep * njobs + job];
            const float alpha = job_alpha(alpha0, alpha_min, ep, epochs, job * job_walks, W);
            const int64_t p_end = (job + 1) * job_walks < W ? (job + 1) * job_walks : W;
            for (int64_t p = job * job_walks; p < p_end; ++p) {
                for (int i = 0; i < L; ++i) walk[i] = tokens[(int64_t)i * W + p];
                int len = 0; /* a negative token pads a ragged walk: the sentence ends there */
                while (len < L && walk[len] >= 0) ++len;
                if (merged) {
                    walk_merged(sampler, &ds, V, d, K, walk, len, p, ep, alpha, exp_table, syn0, syn1neg);
                    continue;
                }
                for (int i = 0; i < len; ++i) {
                    for (int j = i - 1; j <= i + 1; j += 2) { /* window = 1, reduced window 0 */
                        if (j < 0 || j >= len) continue;
                        const int q = 2 * i + (j > i);
                        ds.ctr_p_lo = (uint32_t)(uint64_t)p;
                        ds.ctr_p_hi = (uint32_t)((uint64_t)p >> 32);
                        ds.ctr_base = (uint32_t)q << 4;
                        ds.ctr_tag = ((uint32_t)ep << 8) | GWO_STREAM_SGNS;
                        ds.n_drawn = -1;
                        pair_update(sampler, &ds, V, d, K, walk[i], walk[j], alpha, exp_table,
                                    syn0, syn1neg, work);
                    }
                }
            }
        }
    }
    return 0;
}

int gwo_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
