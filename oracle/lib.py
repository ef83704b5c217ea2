"""ctypes front-end of the C oracle (oracle/gw_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Also holds the numpy restatement of the gensim host-side steps around the SGNS loop
(vocabulary, cum_table, weight init, per-job LCG seeds); see SURVEY.md section 8(c) items 1-4.
gensim is a third-party dependency absent from /root/reference and from this image: PARITY
UNPINNED against real gensim (no golden vector exists in the reference's tests for this step).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'libgw_oracle.so')
_lib = None

c_i32p = ctypes.POINTER(ctypes.c_int32)
c_u32p = ctypes.POINTER(ctypes.c_uint32)
c_u64p = ctypes.POINTER(ctypes.c_uint64)
c_f32p = ctypes.POINTER(ctypes.c_float)


def build(force=False):
    """Compile the C oracle with the committed Makefile (gcc only)."""
    if force or not os.path.exists(_SO) or \
            os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, 'gw_oracle.c')):
        subprocess.check_call(['make', '-C', _HERE, '-s'])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        L.gwo_philox4x32_10.argtypes = [c_u32p, c_u32p, c_u32p]
        L.gwo_philox4x32_10.restype = None
        L.gwo_walk_uniform.argtypes = [ctypes.c_int32, c_i32p, c_i32p, ctypes.c_int32,
                                       ctypes.c_int32, ctypes.c_uint64, ctypes.c_uint64,
                                       ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, c_i32p]
        L.gwo_walk_uniform.restype = ctypes.c_int
        L.gwo_exp_table.argtypes = [c_f32p]
        L.gwo_exp_table.restype = None
        L.gwo_sgns_train.argtypes = [
            ctypes.c_int, ctypes.c_int32, ctypes.c_int32, c_i32p, ctypes.c_int64, ctypes.c_int32,
            ctypes.c_int32, ctypes.c_int32, ctypes.c_double, ctypes.c_double, ctypes.c_int64,
            c_u32p, ctypes.c_int32, c_u64p, c_f32p, c_i32p, ctypes.c_uint64, c_f32p, c_f32p,
            ctypes.c_int, ctypes.c_int]
        L.gwo_sgns_train.restype = ctypes.c_int
        L.gwo_max_threads.argtypes = []
        L.gwo_max_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def philox(ctr, key):
    ctr = np.asarray(ctr, dtype=np.uint32)
    key = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().gwo_philox4x32_10(_p(ctr, c_u32p), _p(key, c_u32p), _p(out, c_u32p))
    return out


def walk_uniform(row_ptr, col_idx, niter, L, seed, stride=1, slot_begin=0, slot_end=None):
    """Position-major tokens [L, nslots] of storage slots [slot_begin, slot_end)."""
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int32)
    col_idx = np.ascontiguousarray(col_idx, dtype=np.int32)
    n = len(row_ptr) - 1
    W = int(niter) * int(row_ptr[-1])
    if slot_end is None:
        slot_end = W
    ns = int(slot_end - slot_begin)
    out = np.empty((L, ns), dtype=np.int32)
    rc = lib().gwo_walk_uniform(n, _p(row_ptr, c_i32p), _p(col_idx, c_i32p), int(niter), int(L),
                                int(seed), int(stride), int(slot_begin), int(slot_end), ns,
                                _p(out, c_i32p))
    if rc != 0:
        raise RuntimeError('gwo_walk_uniform rc=%d' % rc)
    return out


def set_exp_scale(scale):
    """EXP_TABLE index scale of the oracle's sigmoid (default 1000/6/2; the alternative 83.0 is what
    C-style folding of gensim's DEF constants would give): see tools/pin_gensim.py."""
    L = lib()
    L.gwo_set_exp_scale.argtypes = [ctypes.c_double]
    L.gwo_set_exp_scale.restype = None
    L.gwo_set_exp_scale(float(scale))


def exp_table():
    t = np.empty(1000, dtype=np.float32)
    lib().gwo_exp_table(_p(t, c_f32p))
    return t


def sgns_train(tokens, V, d, syn0, syn1neg, sampler, K=5, epochs=5, alpha0=0.025,
               alpha_min=0.0001, job_walks=1000, cum_table=None, job_seeds=None, alias_prob=None,
               alias_idx=None, seed=0, nthreads=1, merged=False):
    """tokens: int32 [L, W] position-major; syn0/syn1neg float32 [V, d] updated in place.
    merged=True: the CUDA kernel's merged reduction order (gw_oracle.c:walk_merged)."""
    tokens = np.ascontiguousarray(tokens, dtype=np.int32)
    L, W = tokens.shape
    assert syn0.dtype == np.float32 and syn0.flags.c_contiguous and syn0.shape == (V, d)
    assert syn1neg.dtype == np.float32 and syn1neg.flags.c_contiguous and syn1neg.shape == (V, d)
    if sampler == 0:
        cum_table = np.ascontiguousarray(cum_table, dtype=np.uint32)
        job_seeds = np.ascontiguousarray(job_seeds, dtype=np.uint64)
        njobs = (W + job_walks - 1) // job_walks
        assert job_seeds.size >= epochs * njobs
        cum_len = len(cum_table)
    else:
        alias_prob = np.ascontiguousarray(alias_prob, dtype=np.float32)
        alias_idx = np.ascontiguousarray(alias_idx, dtype=np.int32)
        assert alias_prob.size == V and alias_idx.size == V
        cum_len = 0
    rc = lib().gwo_sgns_train(int(sampler), int(V), int(d), _p(tokens, c_i32p), int(W), int(L),
                              int(K), int(epochs), float(alpha0), float(alpha_min),
                              int(job_walks), _p(cum_table, c_u32p), int(cum_len),
                              _p(job_seeds, c_u64p), _p(alias_prob, c_f32p),
                              _p(alias_idx, c_i32p), int(seed), _p(syn0, c_f32p),
                              _p(syn1neg, c_f32p), int(nthreads), 1 if merged else 0)
    if rc != 0:
        raise RuntimeError('gwo_sgns_train rc=%d' % rc)


# ---------------------------------------------------------------------------------------------
# gensim host-side steps (restated)
# ---------------------------------------------------------------------------------------------
def gensim_vocab(tokens, n, corpus_order_first_seen=True):
    """gensim build_vocab for min_count=1, sample=0: index words by descending count, ties in
    first-seen order (dict insertion order of the corpus scan + stable sort).

    tokens: [L, W] position-major node ids; corpus scan order is walk-major (walk 0 tokens 0..L-1,
    walk 1 ...).  Returns (index_to_node [V'], node_to_index [n] with -1 for absent, counts [V']).
    """
    tokens = np.asarray(tokens)
    counts = np.zeros(n, dtype=np.int64)
    for row in tokens:      # row by row: no 8-byte copy of the whole corpus
        counts += np.bincount(row[row >= 0], minlength=n)
    present = np.nonzero(counts)[0]
    if corpus_order_first_seen:
        # first position of every node in the walk-major corpus scan (walk 0 tokens 0..L-1, walk 1,
        # ...), found chunk by chunk from the end (a later assignment overwrites an earlier one)
        L, W = tokens.shape
        first = np.full(n, np.iinfo(np.int64).max, dtype=np.int64)
        step = max(1, (1 << 24) // max(L, 1))
        for b in range(((W - 1) // step) * step, -1, -step):
            e = min(W, b + step)
            blk = np.ascontiguousarray(tokens[:, b:e].T).ravel()
            pos = np.arange(b * L, e * L, dtype=np.int64)
            ok = blk >= 0
            first[blk[ok][::-1]] = pos[ok][::-1]
        order = present[np.argsort(first[present], kind='stable')]   # nodes by first appearance
    else:
        order = present
    rank = np.argsort(-counts[order], kind='stable')
    index_to_node = order[rank].astype(np.int32)
    node_to_index = np.full(n, -1, dtype=np.int32)
    node_to_index[index_to_node] = np.arange(len(index_to_node), dtype=np.int32)
    return index_to_node, node_to_index, counts[index_to_node]


def gensim_cum_table(counts, ns_exponent=0.75, domain=2 ** 31 - 1):
    """gensim Word2Vec.make_cum_table: uint32 cumulative count**0.75 scaled to 2^31-1."""
    counts = np.asarray(counts, dtype=np.int64)
    train_words_pow = 0.0
    for c in counts:
        train_words_pow += float(c) ** float(ns_exponent)
    cum = np.zeros(len(counts), dtype=np.uint32)
    cumulative = 0.0
    for i, c in enumerate(counts):
        cumulative += float(c) ** float(ns_exponent)
        cum[i] = round(cumulative / train_words_pow * domain)
    return cum


def gensim_init_weights(V, d, seed=1):
    """gensim KeyedVectors.resize_vectors/prep_vectors + Word2Vec.init_weights."""
    rng = np.random.default_rng(seed=seed)
    syn0 = rng.random((V, d), dtype=np.float32)
    syn0 *= 2.0
    syn0 -= 1.0
    syn0 /= d
    syn1neg = np.zeros((V, d), dtype=np.float32)
    return syn0, syn1neg


def gensim_job_seeds(njobs_total, seed=1):
    """gensim init_w2v_config: next_random = 2^24*randint(0,2^24) + randint(0,2^24), drawn per job
    from the model's RandomState(seed) (single-worker order)."""
    rs = np.random.RandomState(seed)
    out = np.empty(njobs_total, dtype=np.uint64)
    for i in range(njobs_total):
        a = int(rs.randint(0, 2 ** 24))
        b = int(rs.randint(0, 2 ** 24))
        out[i] = (2 ** 24) * a + b
    return out


def word2vec_gensim_like(tokens, n, d=8, K=5, epochs=5, alpha0=0.025, alpha_min=0.0001,
                         job_walks=1000, seed=1, nthreads=1, inplace=False):
    """Whole Word2Vec(sentences=walks, sg=1, vector_size=d, window=1, min_count=1, negative=K,
    sample=0) as genewalk/deepwalk.py:135-139 runs it.  Returns (index_to_node, vectors[V', d])."""
    index_to_node, node_to_index, counts = gensim_vocab(tokens, n)
    V = len(index_to_node)
    # inplace: overwrite the caller's int32 token array with vocabulary indices (a human-scale
    # corpus is 6-8 GB; the study runs several oracle processes side by side)
    tok_idx = tokens if inplace else np.empty(np.shape(tokens), dtype=np.int32)
    for i, row in enumerate(np.asarray(tokens)):
        tok_idx[i] = node_to_index[row]
    cum = gensim_cum_table(counts)
    syn0, syn1neg = gensim_init_weights(V, d, seed)
    W = tok_idx.shape[1]
    njobs = (W + job_walks - 1) // job_walks
    seeds = gensim_job_seeds(njobs * epochs, seed)
    sgns_train(tok_idx, V, d, syn0, syn1neg, sampler=0, K=K, epochs=epochs, alpha0=alpha0,
               alpha_min=alpha_min, job_walks=job_walks, cum_table=cum, job_seeds=seeds,
               nthreads=nthreads)
    return index_to_node, syn0, syn1neg


# ---------------------------------------------------------------------------------------------
# alias table (sequential Vose; the oracle-side statement of "unigram^0.75 alias table")
# ---------------------------------------------------------------------------------------------
def alias_table(counts, power=0.75):
    """Vose alias method over weights counts**power (zero-count rows get probability 0).
    Returns (prob float32 [n], alias int32 [n])."""
    counts = np.asarray(counts, dtype=np.float64)
    n = len(counts)
    w = np.where(counts > 0, counts ** power, 0.0)
    q = w * (n / w.sum())
    prob = np.ones(n, dtype=np.float64)
    alias = np.arange(n, dtype=np.int32)
    small = [i for i in range(n) if q[i] < 1.0]
    large = [i for i in range(n) if q[i] >= 1.0]
    q = q.copy()
    while small and large:
        s = small.pop()
        g = large.pop()
        prob[s] = q[s]
        alias[s] = g
        q[g] = (q[g] + q[s]) - 1.0
        if q[g] < 1.0:
            small.append(g)
        else:
            large.append(g)
    return prob.astype(np.float32), alias


def alias_implied_distribution(prob, alias):
    """Probability each index is returned by (slot uniform, coin < prob ? slot : alias[slot])."""
    prob = np.asarray(prob, dtype=np.float64)
    n = len(prob)
    out = prob / n
    np.add.at(out, np.asarray(alias), (1.0 - prob) / n)
    return out
