"""DeepWalk on the GPU behind genewalk's own API.

Mirrors /root/reference/genewalk/deepwalk.py: ``DeepWalk(graph, walk_length=10, niter=100)`` with
``get_walks(workers)`` and ``word2vec(...)``, module functions ``get_start_nodes``,
``run_walks_for_node``, ``run_single_walk`` and ``run_walks(graph, **kwargs)``.  The arithmetic runs
in libgenewalk_b200.so (hand-written sm_100a CUDA) through the C ABI in include/genewalk_b200.h:
gw_walk_uniform replaces the Python walk loops (deepwalk.py:50-96,145-200) and
gw_count_tokens / gw_vocab_rank / gw_alias_build / gw_sgns_init / gw_sgns_train replace
``gensim.models.Word2Vec`` (deepwalk.py:135-139).  There is no CPU fallback.

Differences a caller can observe (all documented in DESIGN.md):
  * walks are drawn by a counter-based Philox sampler, not CPython's MT19937; the seed is taken
    from ``random.getrandbits(64)`` so ``random.seed(...)`` (cli.py:187-189) still makes runs
    reproducible -- and, unlike the reference, also for ``workers > 1`` (a no-op hint here);
  * ``self.walks`` is a lazy sequence over device memory (``len``, indexing and iteration behave
    like the reference's list of lists of node labels);
  * unsupported Word2Vec settings (sg=0, hs=1, window != 1, sample > 0, min_count > 1) raise
    instead of silently training something else.
"""
import logging
import random
import time
from collections.abc import Sequence

import numpy as np

from . import _lib
from .csr import GraphCSR, coprime_stride
from .keyedvectors import NodeVectors, Word2VecModel

logger = logging.getLogger('genewalk.deepwalk')

default_walk_length = 10
default_niter = 100


def _csr_of(graph):
    """GraphCSR of an nx.MultiGraph (or a GraphCSR / an object carrying one)."""
    if isinstance(graph, GraphCSR):
        return graph
    attrs = getattr(graph, 'graph', None)
    if isinstance(attrs, dict) and isinstance(attrs.get('_gw_csr'), GraphCSR):
        return attrs['_gw_csr']
    return GraphCSR.from_networkx(graph)


class WalkCorpus(Sequence):
    """Lazy list-of-walks view over the device token array.

    Index ``i`` is the walk's position in the reference's serial corpus (node-major: node v owns
    ``niter*len(graph[v])`` consecutive walks, deepwalk.py:67-70,197); the device stores walks in
    slot order (see gw_walk_uniform) and position-major (``tokens[k, p]`` = k-th node of slot p).
    """

    def __init__(self, csr, tokens, niter, stride):
        self.csr = csr
        self.tokens = tokens          # torch int32 [L, W] on the device (or numpy after pickling)
        self.niter = int(niter)
        self.stride = int(stride)
        self.A = csr.nnz
        self._inv = pow(self.stride, -1, self.A) if (self.A > 1 and self.stride > 0) else 0

    def __len__(self):
        return int(self.tokens.shape[1])

    @property
    def walk_length(self):
        return int(self.tokens.shape[0])

    def slots_of(self, idx):
        idx = np.asarray(idx, dtype=np.int64)
        if self.stride == 0:
            return idx
        e, it = idx // self.niter, idx % self.niter
        r = (e * self._inv) % self.A if self.A > 1 else e
        return it * self.A + r

    def _labels(self, cols):
        nodes = self.csr.nodes
        return [[nodes[t] for t in col if t >= 0] for col in cols]

    def _fetch(self, idx):
        slots = self.slots_of(idx)
        if isinstance(self.tokens, np.ndarray):
            return self.tokens[:, slots].T
        import torch
        sl = torch.as_tensor(slots, device=self.tokens.device)
        return self.tokens.index_select(1, sl).t().cpu().numpy()

    def __getitem__(self, i):
        if isinstance(i, slice):
            return self._labels(self._fetch(np.arange(*i.indices(len(self)))))
        n = len(self)
        if i < 0:
            i += n
        if not 0 <= i < n:
            raise IndexError(i)
        return self._labels(self._fetch([i]))[0]

    def __iter__(self, chunk=1 << 16):
        n = len(self)
        for b in range(0, n, chunk):
            for w in self._labels(self._fetch(np.arange(b, min(n, b + chunk)))):
                yield w

    def start_node_ids(self):
        """Node id of the first node of every walk, in corpus order (cheap: row of the entry)."""
        self.csr.to_host()
        return np.repeat(np.repeat(np.arange(self.csr.n), np.diff(self.csr.row_ptr)), self.niter)

    def __getstate__(self):
        st = dict(self.__dict__)
        if not isinstance(self.tokens, np.ndarray):
            st['tokens'] = self.tokens.cpu().numpy()
        self.csr.to_host()
        st['csr'] = GraphCSR(self.csr.nodes, self.csr.row_ptr, self.csr.col_idx)
        return st


class DeepWalk(object):
    """Perform DeepWalk (node2vec), i.e., unbiased random walk over nodes
    on an undirected networkx MultiGraph.

    Parameters
    ----------
    graph : networkx.MultiGraph
        A networkx multigraph to be used as the basis for DeepWalk.
    walk_length : Optional[int]
        The length of each random walk on the graph. Default: 10
    niter : Optional[int]
        The number of iterations for each node to run (this is multiplied by
        the number of neighbors of the node when determining the overall number
        of walks to start from a given node). Default: 100
    seed : Optional[int]
        64-bit seed of the Philox sampler.  Default: ``random.getrandbits(64)``, so Python's
        ``random.seed`` governs reproducibility as in the reference.

    Attributes
    ----------
    walks : sequence
        The walks (lazy sequence of lists of node labels).
    """

    def __init__(self, graph, walk_length=default_walk_length, niter=default_niter, seed=None):
        self.graph = graph
        self.walks = []
        self.wl = walk_length
        self.niter = niter
        self.model = None
        self._pending = None
        self.seed = seed
        self._csr = None
        self.timings = {}

    # ------------------------------------------------------------------------------------------
    def _get_csr(self):
        if self._csr is None:
            self._csr = _csr_of(self.graph)
        return self._csr

    def get_walks(self, workers=1, stride=None):
        """Generates walks (sentences) sampled by an (unbiased) random walk over the graph.

        ``workers`` is accepted for API compatibility (the GPU runs one thread per walk).
        ``stride`` selects the storage order (0, default: the reference's node-major corpus order,
        which is also the order the learning-rate schedule of ``word2vec`` refers to; s > 0:
        iteration-major with adjacency entries permuted by a stride coprime to A).
        """
        import torch
        _lib.require_cuda()
        logger.info('Running random walks...')
        start = time.time()
        if int(self.wl) < 1 or int(self.niter) < 1:
            raise ValueError('walk_length and niter must be positive')
        seed = self.seed if self.seed is not None else random.getrandbits(64)
        self._walk_seed = int(seed) & (2 ** 64 - 1)
        csr = self._get_csr().to_device()
        A = csr.nnz
        if A == 0:
            self.walks = []
            return
        W = int(self.niter) * A
        stride = 0 if stride is None else int(stride)
        tokens = torch.empty((int(self.wl), W), dtype=torch.int32, device=csr.d_row_ptr.device)
        _lib.call('gw_walk_uniform', csr.n, A, _lib.ptr(csr.d_row_ptr), _lib.ptr(csr.d_col_idx),
                  int(self.niter), int(self.wl), self._walk_seed, stride, 0, W, _lib.ptr(tokens),
                  _lib.stream_ptr())
        self.walks = WalkCorpus(csr, tokens, self.niter, stride)
        end = time.time()
        self.timings['get_walks_enqueue_s'] = end - start
        logger.info('Running random walks done in %.2fs' % (end - start))

    # ------------------------------------------------------------------------------------------
    def _device_corpus(self):
        """(csr-like node list, n, tokens [L, W] device tensor) of ``self.walks``."""
        import torch
        if isinstance(self.walks, WalkCorpus):
            tok = self.walks.tokens
            if isinstance(tok, np.ndarray):
                tok = torch.from_numpy(tok).cuda()
            return self.walks.csr.nodes, self.walks.csr.n, tok
        # a plain list of lists of labels (what the reference stores): encode, pad ragged with -1
        walks = list(self.walks)
        if not walks:
            raise RuntimeError('you must first build vocabulary before training the model')
        nodes, index = [], {}
        L = max(len(w) for w in walks)
        arr = np.full((L, len(walks)), -1, dtype=np.int32)
        for p, w in enumerate(walks):
            for k, t in enumerate(w):
                i = index.get(t)
                if i is None:
                    i = index[t] = len(nodes)
                    nodes.append(t)
                arr[k, p] = i
        return nodes, len(nodes), torch.from_numpy(arr).cuda()

    def word2vec(self, sg=1, vector_size=8, window=1, min_count=1, negative=5,
                 workers=1, sample=0, **kwargs):
        """Set the model based on Word2Vec (skip-gram with negative sampling, trained on the GPU).

        Same signature and defaults as the reference (deepwalk.py:98-99); the remaining gensim
        defaults can be overridden by keyword: ``epochs`` (5), ``alpha`` (0.025), ``min_alpha``
        (0.0001), ``seed`` (1, weight init), ``ns_exponent`` (0.75), ``batch_words`` (10000),
        ``hs`` (0).  GPU-only knobs: ``lanes`` (number of concurrent workers, the analogue of
        gensim's ``workers``: each trains one 1000-walk job at a time, taken from a queue in corpus
        order; 0 = library default, 1 = sequential), ``chunk_walks`` (walks a worker takes from
        the queue at a time; must divide the job size; 0 = library default), ``streams`` (corpus
        slices the queue deals chunks from in turn; 0 = library default), ``concurrent`` (number of
        trainings the caller runs at once on this GPU on different CUDA streams: they share the
        GPU's resident workers and leave the SMs' shared memory to each other; default 1),
        ``reductions`` ('auto' | 'per_pair' | 'merged': see ``gw_sgns_flags`` in
        include/genewalk_b200.h), ``wait`` (False: only enqueue the training on the current CUDA
        stream; ``finish_word2vec()`` then downloads the vectors) and ``sgns_seed`` (Philox key of
        the negative draws).
        """
        import torch
        _lib.require_cuda()
        epochs = int(kwargs.pop('epochs', 5))
        alpha = float(kwargs.pop('alpha', 0.025))
        min_alpha = float(kwargs.pop('min_alpha', 0.0001))
        init_seed = int(kwargs.pop('seed', 1))
        ns_exponent = float(kwargs.pop('ns_exponent', 0.75))
        batch_words = int(kwargs.pop('batch_words', 10000))
        hs = int(kwargs.pop('hs', 0))
        lanes = int(kwargs.pop('lanes', 0))
        chunk_walks = int(kwargs.pop('chunk_walks', 0))
        streams = int(kwargs.pop('streams', 0))
        concurrent = int(kwargs.pop('concurrent', 1))
        shared_sm = bool(kwargs.pop('shared_sm', concurrent > 1))
        reductions = kwargs.pop('reductions', 'auto')
        kwargs_async = not bool(kwargs.pop('wait', True))
        sgns_seed = kwargs.pop('sgns_seed', None)
        if kwargs:
            raise TypeError('word2vec() got unexpected keyword arguments %s' % sorted(kwargs))
        if sg != 1:
            raise ValueError('only skip-gram (sg=1) is implemented; GeneWalk always uses sg=1')
        if hs != 0 or negative <= 0:
            raise ValueError('only negative sampling (hs=0, negative>0) is implemented')
        if window != 1:
            raise ValueError('only window=1 is implemented; GeneWalk always uses window=1')
        if sample != 0:
            raise ValueError('sub-sampling (sample>0) is not implemented; GeneWalk uses sample=0')
        if min_count not in (0, 1):
            raise ValueError('min_count>1 is not implemented; GeneWalk uses min_count=1')
        logger.info('Generating node vectors...')
        start = time.time()
        self._w2v_args = dict(sg=1, vector_size=int(vector_size), window=window, min_count=min_count,
                              negative=negative, sample=sample, epochs=epochs, alpha=alpha,
                              min_alpha=min_alpha, seed=init_seed, ns_exponent=ns_exponent,
                              workers=workers)
        self._pending = self.train_device(vector_size=vector_size, negative=negative, epochs=epochs,
                                          alpha=alpha, min_alpha=min_alpha, init_seed=init_seed,
                                          ns_exponent=ns_exponent, batch_words=batch_words,
                                          lanes=lanes, chunk_walks=chunk_walks, streams=streams,
                                          shared_sm=shared_sm, reductions=reductions, concurrent=concurrent,
                                          sgns_seed=sgns_seed)
        if kwargs_async:
            return          # training is enqueued; finish_word2vec() downloads the vectors
        self.finish_word2vec()
        end = time.time()
        self.timings['word2vec_s'] = end - start
        logger.info('Generating node vectors done in %.2fs' % (end - start))

    def finish_word2vec(self):
        """Download the vectors of an enqueued training (``word2vec(..., wait=False)``) and set
        ``self.model`` (the only host synchronisation of a replicate)."""
        dv = self._pending
        self._pending = None
        nodes = dv['nodes']
        nv = int(dv['n_vocab'].item())
        order = dv['node_of_rank'][:nv].long()
        vectors = dv['syn0'].index_select(0, order).cpu().numpy()
        out_layer = dv['syn1neg'].index_select(0, order).cpu().numpy()
        cnt = dv['counts'].index_select(0, order).cpu().numpy()
        order_h = order.cpu().numpy()
        wv = NodeVectors([nodes[i] for i in order_h], vectors, counts=cnt)
        wv.node_ids = order_h.astype(np.int32)   # row r of wv.vectors is graph node node_ids[r]
        self.model = Word2VecModel(wv, syn1neg=out_layer, corpus_count=dv['W'], **self._w2v_args)
        return self.model

    def train_device(self, vector_size=8, negative=5, epochs=5, alpha=0.025, min_alpha=0.0001,
                     init_seed=1, ns_exponent=0.75, batch_words=10000, lanes=0, chunk_walks=0,
                     streams=0, shared_sm=None, reductions='auto', concurrent=1, sgns_seed=None,
                     epoch_hook=None):
        """Enqueue vocabulary, alias table, init and the SGNS epochs on the current stream and
        return the device tensors (no host synchronisation).  ``epoch_hook(ep, when)`` is called
        with when in {'before', 'after'} around every epoch's launch (bench.py records CUDA
        events there)."""
        import torch
        nodes, n, tokens = self._device_corpus()
        L, W = int(tokens.shape[0]), int(tokens.shape[1])
        d = int(vector_size)
        dev = tokens.device
        st = _lib.stream_ptr()
        if sgns_seed is None:
            sgns_seed = (getattr(self, '_walk_seed', 0) ^ 0x5DEECE66D5DEECE6) & (2 ** 64 - 1)
        counts = torch.empty(n, dtype=torch.int64, device=dev)
        rank_of_node = torch.empty(n, dtype=torch.int32, device=dev)
        node_of_rank = torch.empty(n, dtype=torch.int32, device=dev)
        n_vocab = torch.zeros(1, dtype=torch.int32, device=dev)
        alias = torch.empty((n, 2), dtype=torch.int32, device=dev)   # gw_alias_entry[n]
        syn0 = torch.empty((n, d), dtype=torch.float32, device=dev)
        syn1neg = torch.empty((n, d), dtype=torch.float32, device=dev)
        # gensim init_weights: rows of default_rng(seed).random((V, d), float32); V <= n and the
        # generator fills row-major, so the first V rows of an (n, d) draw are the (V, d) draw
        rng = np.random.default_rng(seed=init_seed)
        rows = rng.random((n, d), dtype=np.float32)
        rows *= 2.0
        rows -= 1.0
        rows /= d
        rng_rows = torch.from_numpy(rows).to(dev)
        _lib.call('gw_count_tokens', _lib.ptr(tokens), L * W, n, _lib.ptr(counts), st)
        _lib.call('gw_vocab_rank', n, _lib.ptr(counts), _lib.ptr(rank_of_node),
                  _lib.ptr(node_of_rank), _lib.ptr(n_vocab), st)
        _lib.call('gw_alias_build', n, _lib.ptr(counts), float(ns_exponent), _lib.ptr(alias), st)
        _lib.call('gw_sgns_init', n, d, _lib.ptr(rank_of_node), _lib.ptr(rng_rows), _lib.ptr(syn0),
                  _lib.ptr(syn1neg), st)
        job_walks = max(1, int(batch_words) // max(L, 1))
        if int(lanes) == 0 or int(chunk_walks) == 0 or int(streams) == 0:
            import ctypes
            a_l, a_c, a_s = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0)
            _lib.call('gw_sgns_auto_shape', n, W, job_walks, max(1, int(concurrent)),
                      ctypes.byref(a_l), ctypes.byref(a_c),
                      ctypes.byref(a_s))
            if int(lanes) == 0:
                lanes = a_l.value
            if int(chunk_walks) == 0:
                chunk_walks = job_walks if int(lanes) == 1 else a_c.value
            if int(streams) == 0:
                streams = 1 if int(lanes) == 1 else min(a_s.value, int(lanes))
        if shared_sm is None:
            shared_sm = int(concurrent) > 1
        flags = (1 if shared_sm else 0) | {'auto': 0, 'per_pair': 2, 'merged': 4}[reductions]
        for ep in range(int(epochs)):
            if epoch_hook is not None:
                epoch_hook(ep, 'before')
            _lib.call('gw_sgns_train', n, d, _lib.ptr(tokens), W, L, 1, int(negative), int(epochs),
                      ep, ep + 1, float(alpha), float(min_alpha), job_walks, int(chunk_walks),
                      int(streams), _lib.ptr(alias),
                      _lib.ptr(syn0), _lib.ptr(syn1neg), int(sgns_seed), int(lanes),
                      flags, st)
            if epoch_hook is not None:
                epoch_hook(ep, 'after')
        return {'nodes': nodes, 'n': n, 'W': W, 'L': L, 'syn0': syn0, 'syn1neg': syn1neg,
                'counts': counts, 'rank_of_node': rank_of_node, 'node_of_rank': node_of_rank,
                'n_vocab': n_vocab, 'alias': alias, 'rng_rows': rng_rows,
                'lanes': int(lanes), 'chunk_walks': int(chunk_walks), 'streams': int(streams),
                'job_walks': job_walks,
                'h2d_bytes': rows.nbytes}


# ---------------------------------------------------------------------------------------------
# module-level functions of the reference API
# ---------------------------------------------------------------------------------------------
def get_start_nodes(graph, niter):
    """Start node of every walk, corpus order: each node ``niter * len(graph[node])`` times
    (deepwalk.py:170-174)."""
    csr = _csr_of(graph)
    csr.to_host()
    deg = np.diff(csr.row_ptr)
    return [csr.nodes[v] for v in np.repeat(np.arange(csr.n), deg * int(niter))]


def run_walks_for_node(node, graph, niter, walk_length, seed=None):
    """Run all random walks starting from a given node (deepwalk.py:177-200)."""
    import torch
    _lib.require_cuda()
    csr = _csr_of(graph).to_device()
    csr.to_host()
    v = csr.index[node]
    b, e = int(csr.row_ptr[v]), int(csr.row_ptr[v + 1])
    if e == b:
        return []
    seed = random.getrandbits(64) if seed is None else int(seed)
    A, it_n = csr.nnz, int(niter)
    # plain iteration-major slots (stride 1): walk (e, it) sits in slot it*A + e
    out = []
    buf = torch.empty((int(walk_length), e - b), dtype=torch.int32, device=csr.d_row_ptr.device)
    cols = []
    for it in range(it_n):
        _lib.call('gw_walk_uniform', csr.n, A, _lib.ptr(csr.d_row_ptr), _lib.ptr(csr.d_col_idx),
                  it_n, int(walk_length), seed, 1, it * A + b, it * A + e, _lib.ptr(buf),
                  _lib.stream_ptr())
        cols.append(buf.t().cpu().numpy().copy())
    arr = np.stack(cols, axis=1).reshape(-1, int(walk_length))   # entry-major, iteration-minor
    for row in arr:
        out.append([csr.nodes[t] for t in row])
    return out


def run_single_walk(start_node, graph, length, seed=None):
    """Run a single random walk on a graph from a given start node (deepwalk.py:145-167)."""
    walks = run_walks_for_node(start_node, graph, 1, length, seed=seed)
    if not walks:
        raise IndexError('Cannot choose from an empty sequence')   # random.choice on no neighbours
    return walks[0]


def run_walks(graph, **kwargs):
    """Run random walks and get node vectors on a given graph.

    Parameters
    ----------
    graph : networkx.MultiGraph
        The graph on which random walks are going to be run and node vectors
        calculated.
    **kwargs
        Key word arguments passed as the arguments of the DeepWalk constructor,
        as well as the get_walks method and the word2vec method. See the
        DeepWalk class documentation for more information on these.

    Returns
    -------
    :py:class:`genewalk_b200.deepwalk.DeepWalk`
        A DeepWalk instance whose walks attribute contains the random walks produced on the graph
        and whose ``model.wv`` holds the node vectors.
    """
    dw_args = {'walk_length': kwargs.pop('walk_length', default_walk_length),
               'niter': kwargs.pop('niter', default_niter),
               'seed': kwargs.pop('walk_seed', None)}
    DW = DeepWalk(graph, **dw_args)
    DW.get_walks(kwargs.get('workers', 1))
    DW.word2vec(**kwargs)
    return DW
