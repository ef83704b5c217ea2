/*
 * genewalk_b200.h -- C ABI of libgenewalk_b200.so (hand-written sm_100a CUDA).
 *
 * This is the drop-in boundary for GeneWalk's representation-learning hot path.  The reference
 * (churchmanlab/genewalk v1.6.3) has no FFI of its own: the path is Python calling networkx,
 * gensim, numpy and statsmodels.  Each entry point below therefore names the reference call site
 * whose arithmetic it replaces (file:line under /root/reference); the Python package
 * genewalk_b200/ re-creates the reference's classes/functions on top of these calls via ctypes
 * (see INTEGRATION.md for the binding a genewalk maintainer would add).
 *
 * Conventions
 *   - plain C types only; every pointer is a CALLER-OWNED DEVICE pointer unless marked "host";
 *   - every call takes the CUDA stream to enqueue on (cudaStream_t passed as void*); calls are
 *     asynchronous unless stated; temporaries use stream-ordered cudaMallocAsync;
 *   - return 0 on success, negative gw_status otherwise; gw_last_error() gives the message of the
 *     calling thread's last failure; no global mutable state, thread-compatible;
 *   - graph = CSR over DISTINCT neighbours: col_idx[row_ptr[v]+k] is the k-th entry of
 *     list(graph[v]) (deepwalk.py:165); a self-loop lists v once in its own row; A = row_ptr[n];
 *   - walks are stored position-major: tokens[i*W + p] = i-th node of the walk in storage slot p.
 */
#ifndef GENEWALK_B200_H
#define GENEWALK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GW_VERSION 200

typedef void *gw_stream_t; /* cudaStream_t */

enum gw_status {
    GW_OK = 0,
    GW_ERR_INVALID = -1,     /* bad argument (Python side raises ValueError) */
    GW_ERR_CUDA = -2,        /* CUDA runtime error */
    GW_ERR_UNSUPPORTED = -3, /* configuration the kernels do not implement */
    GW_ERR_NO_DEVICE = -4
};

/* packed alias-table entry: slot accepted with probability prob, else redirected to alias */
typedef struct {
    float prob;
    int32_t alias;
} gw_alias_entry;

int gw_version(void);
const char *gw_last_error(void);
/* number of kernels this library has launched since it was loaded (statistic for benchmarks) */
long long gw_launch_count(void);
/* number of CUDA devices visible, or a negative gw_status */
int gw_device_count(void);

/* ---- (1) graph -------------------------------------------------------------------------- */

/* Distinct-neighbour CSR from an undirected multi-edge list (u[k], v[k]), k < m: symmetrise,
 * sort, drop duplicates; rows come out sorted by neighbour id; a self-loop is kept once.
 * Replaces the networkx adjacency that `graph[node]` exposes for a generated graph
 * (null_distributions.py:25,40; deepwalk.py:165).  col_idx must hold 2*m entries.
 * SYNCHRONOUS on `stream` for the 8-byte nnz read-back: *nnz_out_host = A. */
int gw_csr_from_edges(int32_t n, int64_t m, const int32_t *u, const int32_t *v, int32_t *row_ptr,
                      int32_t *col_idx, int64_t *nnz_out_host, gw_stream_t stream);

/* Configuration model: replaces nx.configuration_model(d_seq) as called by get_rand_graph
 * (null_distributions.py:23-25; networkx/generators/degree_seq.py:_configuration_model).
 * degrees[n] must already be sorted descending (the caller's `sorted(..., reverse=True)`), sum S
 * even.  Stub list [i]*d_i, uniform shuffle = stable sort on 64-bit Philox keys
 * (ctr {lo32(k), hi32(k), 0, 0x43}, key = seed), first half paired with second half; self-loops and
 * parallel edges are kept.  Writes the S/2 edges (edge_u[k], edge_v[k]) in pairing order. */
int gw_config_model(int32_t n, const int32_t *degrees, int64_t S, uint64_t seed, int32_t *edge_u,
                    int32_t *edge_v, gw_stream_t stream);

/* ---- (2) walks -------------------------------------------------------------------------- */

/* Uniform first-order random walks: replaces DeepWalk.get_walks / run_walks_for_node /
 * run_single_walk (deepwalk.py:50-96,145-200).  W = niter*A walks; canonical walk id
 * w = e*niter + it is the walk's position in the reference's serial corpus (node-major).
 * Storage slot p holds walk (it = p / A, e = ((p % A) * stride) % A); stride must be coprime to A
 * (1 = plain iteration-major order; 0 = slot p holds walk w = p, the reference's corpus order).
 * Step s (1..L-1) of walk w consumes word (s-1)&3 of
 * Philox4x32-10(ctr {lo32(w), hi32(w), (s-1)>>2, 0x57}, key {lo32(seed), hi32(seed)}) and moves
 * to col_idx[row_ptr[v] + mulhi32(word, deg(v))].
 * Writes slots [slot_begin, slot_end) to tokens[i*(slot_end-slot_begin) + (p - slot_begin)].
 * nnz = A = row_ptr[n] (the caller knows it; passing it keeps the call asynchronous). */
int gw_walk_uniform(int32_t n, int64_t nnz, const int32_t *row_ptr, const int32_t *col_idx,
                    int32_t niter, int32_t L, uint64_t seed, uint64_t stride, int64_t slot_begin,
                    int64_t slot_end, int32_t *tokens, gw_stream_t stream);

/* counts[v] = number of occurrences of v in the W = niter*nnz walks gw_walk_uniform(seed) draws,
 * without storing them (gensim build_vocab's raw counts; the trainer regenerates the walks). */
int gw_walk_count(int32_t n, int64_t nnz, const int32_t *row_ptr, const int32_t *col_idx,
                  int32_t niter, int32_t L, uint64_t seed, int64_t *counts, gw_stream_t stream);

/* ---- (3) skip-gram negative sampling ------------------------------------------------------
 * Replaces gensim.models.Word2Vec(sentences=walks, sg=1, vector_size=d, window=1, min_count=1,
 * negative=K, sample=0) as called at deepwalk.py:135-139. */

/* counts[v] = number of occurrences of v among `count` tokens (gensim build_vocab's raw counts);
 * negative tokens (padding of ragged walks) are ignored */
int gw_count_tokens(const int32_t *tokens, int64_t count, int32_t n, int64_t *counts,
                    gw_stream_t stream);

/* rank_of_node[v] = index of v in gensim's index_to_key (descending count; ties by node id;
 * -1 when counts[v] == 0, i.e. the node is absent from the vocabulary, min_count=1);
 * node_of_rank = the inverse; *n_vocab (device int32) = vocabulary size. */
int gw_vocab_rank(int32_t n, const int64_t *counts, int32_t *rank_of_node, int32_t *node_of_rank,
                  int32_t *n_vocab, gw_stream_t stream);

/* alias table for the noise distribution counts**power / sum (gensim ns_exponent=0.75;
 * replaces make_cum_table + bisect).  Parallel construction, see csrc/gw_sgns.cu. */
int gw_alias_build(int32_t n, const int64_t *counts, double power, gw_alias_entry *table,
                   gw_stream_t stream);

/* syn0[v] = rng_rows[rank_of_node[v]] (zero row when absent), syn1neg = 0.  rng_rows is the
 * [n, d] matrix (default_rng(seed=1).random((V, d), float32)*2-1)/d of gensim's init_weights,
 * generated by the host and uploaded once. */
int gw_sgns_init(int32_t n, int32_t d, const int32_t *rank_of_node, const float *rng_rows,
                 float *syn0, float *syn1neg, gw_stream_t stream);

/* One work unit of the trainer's queue: `count` consecutive walks starting at walk id `first`;
 * `node` = start node of walk `first` (owner of adjacency entry first / niter; -1 when the plan was
 * made without a graph).  16 bytes, 16-byte aligned. */
typedef struct {
    int64_t first;
    int32_t count;
    int32_t node;
} gw_sgns_unit;

/* How the corpus is cut into units (gw_sgns_plan):
 *   GW_PLAN_CHUNK  units of unit_walks consecutive walks (the last one shorter), like gensim's
 *                  jobs; a unit may span several start nodes;
 *   GW_PLAN_BURST  a unit never crosses a start node: the niter*len(graph[v]) walks that start at v
 *                  (deepwalk.py:67-70,197) are one unit, cut into ceil(len/unit_walks) near-equal
 *                  pieces when longer than unit_walks -- but into at most max_pieces (> 0) pieces, so
 *                  that no more than max_pieces workers ever train walks of one start node at the
 *                  same time (a hub's rows overshoot when too many stale updates meet in them). */
enum gw_plan_policy { GW_PLAN_CHUNK = 0, GW_PLAN_BURST = 1 };

/* number of gw_sgns_unit entries gw_sgns_plan may write (host-side arithmetic) */
int64_t gw_sgns_plan_capacity(int32_t n, int64_t W, int64_t unit_walks, int32_t policy);

/* Writes the units of a W = niter*A walk corpus in corpus order and *num_units (device int64).
 * row_ptr may be NULL for GW_PLAN_CHUNK (corpora that do not come from gw_walk_uniform);
 * max_pieces (0 = unlimited) applies to GW_PLAN_BURST. */
int gw_sgns_plan(int32_t n, const int32_t *row_ptr, int32_t niter, int64_t W, int64_t unit_walks,
                 int32_t policy, int64_t max_pieces, gw_sgns_unit *units, int64_t capacity,
                 int64_t *num_units, gw_stream_t stream);

/* Trains epochs [epoch_begin, epoch_end) of `epochs` over the W walks of the corpus (window 1), in
 * the corpus order the learning-rate schedule refers to (for GeneWalk: the reference's node-major
 * order, walk id w = e*niter + it).
 * The walks are either read from `tokens` (position-major [L][W], slot p = walk p; a negative token
 * ends its walk early: ragged corpora are padded with -1) or, with tokens == NULL, regenerated by
 * the workers from (row_ptr, col_idx, niter, walk_seed) exactly as gw_walk_uniform draws them, so
 * that the corpus never has to exist in memory (C2: 6 GB per replicate).
 * Parallel structure = gensim's: the corpus is cut into jobs of job_walks consecutive walks
 * (gensim batch_words=10000 -> 1000 walks of 10); every walk of job j of epoch ep is trained with
 * alpha = max(alpha_min, alpha0 - (alpha0-alpha_min) * ((ep + j*job_walks/W) / epochs))
 * (gensim _get_next_alpha).  `lanes` workers (a worker = one group of d/2 <= 32 GPU threads; the
 * analogue of gensim's `workers`) take the units of `units` (gw_sgns_plan) from a queue in order
 * and train each unit's walks one after the other, racing on the tables with red.global.add
 * (Hogwild).  lanes = 1 is the sequential statement of the algorithm (bit-exact against the
 * oracle for any plan).
 * Pair q = 2i + (j>i) of walk w draws its K negatives from Philox4x32-10(ctr {lo32(w), hi32(w),
 * (q<<4)|c, (ep<<8)|0x53}, key = seed): negative t uses words 2(t-1), 2(t-1)+1:
 * slot = mulhi32(x0, n), target = (x1>>16) < min(65535, floor(prob[slot]*2^16)) ? slot : alias[slot];
 * a draw equal to the centre is skipped.  Update rule = gensim w2v_fast_sentence_sg_neg, all
 * arithmetic unfused fp32; the dot product sums d/2 (<= 32) chunks left to right and combines the
 * chunk sums by a pairwise tree.  d in {8,16,32,64,128}; K <= 32; L <= 4096.
 * flags (gw_sgns_flags):
 *   GW_SGNS_MERGED  how a worker's updates of the rows it visits twice reach memory.  Default:
 *     one reduction per (pair, row) in gensim's order.  MERGED: the two updates of syn1neg[centre]
 *     inside a step, and the two updates of syn0[walk[j]] from steps j-1 and j+1, are summed in
 *     registers and reduced once; rows are still read from the table right before use and the
 *     worker adds its own held-back update, so its arithmetic sees every one of its updates.
 *     Both orders are bit-exact against the oracle when run with lanes == 1.  MERGED is
 *     implemented for K == 5 (GeneWalk's setting); other K return GW_ERR_UNSUPPORTED with it. */
enum gw_sgns_flags { GW_SGNS_MERGED = 4 };
int gw_sgns_train(int32_t n, int32_t d, const int32_t *tokens, const int32_t *row_ptr,
                  const int32_t *col_idx, int32_t niter, uint64_t walk_seed, int64_t W, int32_t L,
                  int32_t window, int32_t K, int32_t epochs, int32_t epoch_begin,
                  int32_t epoch_end, double alpha0, double alpha_min, int64_t job_walks,
                  const gw_sgns_unit *units, const int64_t *num_units,
                  const gw_alias_entry *alias_table, float *syn0, float *syn1neg, uint64_t seed,
                  int64_t lanes, int32_t flags, gw_stream_t stream);

/* The default (lanes, unit_walks, policy) of a training (host pointers).  A function of the
 * network alone (n, W and max_start_walks = niter * the largest number of distinct neighbours of a
 * node, i.e. the longest run of walks with one start node; pass -1 when unknown): it does not
 * depend on how many trainings share the GPU, nor on how many GPUs the replicates are spread over,
 * so a seed gives the same statistics at every GPU count. */
int gw_sgns_auto_shape(int32_t n, int64_t W, int64_t job_walks, int64_t max_start_walks,
                       int64_t *lanes_out, int64_t *unit_walks_out, int32_t *policy_out);

/* How many trainings of `lanes` workers (vector size d) are resident on the current device at
 * once: the replicate driver keeps at most that many in flight, so that every training really runs
 * with the workers it was given whatever else shares the GPU. */
int gw_sgns_resident_trainings(int32_t d, int64_t lanes);

/* ---- (4) similarity statistics ------------------------------------------------------------ */

/* sims[k] = dot(unitvec(V[a_idx[k]]), unitvec(V[b_idx[k]])) in float32: replaces
 * KeyedVectors.similarity at perform_statistics.py:102. */
int gw_cosine_pairs(int32_t d, const float *vectors, int64_t m, const int32_t *a_idx,
                    const int32_t *b_idx, float *sims, gw_stream_t stream);

/* For every CSR entry (u, v) with u != v, in CSR order: 1 - (1 - dot(V[v],V[u])/(|V[u]|*|V[v]|)),
 * float32: replaces get_null_distributions (null_distributions.py:33-48).  sims_out holds up to A
 * values; SYNCHRONOUS for the count read-back: *count_out_host = number written. */
int gw_null_similarities(int32_t n, const int32_t *row_ptr, const int32_t *col_idx, int32_t d,
                         const float *vectors, float *sims_out, int64_t *count_out_host,
                         gw_stream_t stream);

/* ascending in-place sort: replaces np.asarray(sorted(srd)) at cli.py:238 */
int gw_sort_f32(float *keys, int64_t count, gw_stream_t stream);

/* ranks[k] = np.searchsorted(null_sorted, sims[k]) (side='left'); pvals[k] =
 * max(1 - ranks[k]/n_null, 1e-16) in float64: replaces GeneWalk.psim
 * (perform_statistics.py:247-258).  ranks may be NULL. */
int gw_psim(const float *null_sorted, int64_t n_null, const float *sims, int64_t m,
            int64_t *ranks, double *pvals, gw_stream_t stream);

/* Benjamini-Hochberg adjusted p-values per segment [seg_ptr[s], seg_ptr[s+1]): replaces
 * statsmodels fdrcorrection(pvals, method='indep')[1] at perform_statistics.py:112 (one segment
 * per gene) and :273 (one segment = all rows). */
int gw_bh_segmented(const double *pvals, int64_t m, const int64_t *seg_ptr, int64_t nseg,
                    double *qvals, gw_stream_t stream);

/* out[r] = {g, g*s^(-1.96/sqrt(nreps)), g*s^(+1.96/sqrt(nreps))} with g = gmean(vals+1e-16)-1e-16,
 * s = gstd(vals+1e-16) (1e-16 when nreps == 1) over the nreps values vals[r*nreps ..]:
 * replaces GeneWalk.log_stats (perform_statistics.py:117-124). */
int gw_log_stats(const double *vals, int64_t rows, int32_t nreps, double *out,
                 gw_stream_t stream);

/* mean[r] = np.mean(sims[r, :]) in float32, sem[r] = np.std(sims[r, :]) / sqrt(nreps) (float64) over
 * the nreps per-replicate similarities sims[r*nreps ..]: replaces perform_statistics.py:180-184. */
int gw_mean_sem_f32(const float *sims, int64_t rows, int32_t nreps, float *mean, double *sem,
                    gw_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GENEWALK_B200_H */
