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
from .csr import GraphCSR
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
    """Lazy list-of-walks view.

    Index ``i`` is the walk's position in the reference's serial corpus (node-major: node v owns
    ``niter*len(graph[v])`` consecutive walks, deepwalk.py:67-70,197).  Walks are a pure function
    of (graph, seed, walk id) -- gw_walk_uniform draws walk w from its own Philox counter -- so the
    corpus is not stored: slices are generated on the device when they are read, the trainer
    regenerates every walk in its own registers, and ``tokens`` (the full position-major
    ``[L, W]`` int32 device array, 6 GB at human scale) is only materialised when asked for.
    With ``stride > 0`` the storage order is iteration-major (see gw_walk_uniform).
    """

    def __init__(self, csr, niter, walk_length, seed, stride=0):
        self.csr = csr
        self.niter = int(niter)
        self.wl = int(walk_length)
        self.seed = int(seed)
        self.stride = int(stride)
        self.A = csr.nnz
        self._tokens = None
        self._inv = pow(self.stride, -1, self.A) if (self.A > 1 and self.stride > 0) else 0

    def __len__(self):
        return self.niter * self.A

    @property
    def walk_length(self):
        return self.wl

    def generate(self, slot_begin, slot_end):
        """Device int32 tensor [L, slot_end - slot_begin] of the walks in those storage slots."""
        import torch
        csr = self.csr.to_device()
        out = torch.empty((self.wl, int(slot_end - slot_begin)), dtype=torch.int32,
                          device=csr.d_row_ptr.device)
        if slot_end > slot_begin:
            _lib.call('gw_walk_uniform', csr.n, self.A, _lib.ptr(csr.d_row_ptr), _lib.ptr(csr.d_col_idx),
                      self.niter, self.wl, self.seed, self.stride, int(slot_begin), int(slot_end),
                      _lib.ptr(out), _lib.stream_ptr())
        return out

    @property
    def tokens(self):
        """The whole corpus as a device tensor [L, W] (materialised on first access)."""
        if self._tokens is None:
            self._tokens = self.generate(0, len(self))
        return self._tokens

    def release(self):
        """Drop a materialised token array (the walks themselves stay reproducible)."""
        self._tokens = None

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
        if isinstance(self._tokens, np.ndarray):
            return self._tokens[:, slots].T
        if len(slots) == 0:
            return np.zeros((0, self.wl), dtype=np.int32)
        if self._tokens is None and np.all(np.diff(slots) == 1):   # a contiguous run: generate it
            return self.generate(int(slots[0]), int(slots[-1]) + 1).t().cpu().numpy()
        if self._tokens is None and len(slots) == 1:
            return self.generate(int(slots[0]), int(slots[0]) + 1).t().cpu().numpy()
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
        if st['_tokens'] is not None and not isinstance(st['_tokens'], np.ndarray):
            st['_tokens'] = st['_tokens'].cpu().numpy()
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

    def get_walks(self, workers=1, stride=None, materialize=False):
        """Generates walks (sentences) sampled by an (unbiased) random walk over the graph.

        ``workers`` is accepted for API compatibility (the GPU runs one thread per walk).
        ``stride`` selects the storage order (0, default: the reference's node-major corpus order,
        which is also the order the learning-rate schedule of ``word2vec`` refers to; s > 0:
        iteration-major with adjacency entries permuted by a stride coprime to A).
        The walks are a lazy sequence (see ``WalkCorpus``): nothing is stored unless
        ``materialize`` is set or ``self.walks.tokens`` is read.
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
        if csr.nnz == 0:
            self.walks = []
            return
        self.walks = WalkCorpus(csr, self.niter, self.wl, self._walk_seed, 0 if stride is None else int(stride))
        if materialize:
            self.walks.tokens
        end = time.time()
        self.timings['get_walks_enqueue_s'] = end - start
        logger.info('Running random walks done in %.2fs' % (end - start))

    # ------------------------------------------------------------------------------------------
    def _device_corpus(self):
        """(node list, n, tokens [L, W] device tensor or None, WalkCorpus or None) of ``self.walks``.
        tokens is None when the trainer can regenerate the walks itself (a lazy corpus in the
        reference's order)."""
        import torch
        if isinstance(self.walks, WalkCorpus):
            wc = self.walks
            if wc.stride == 0 and wc._tokens is None:
                return wc.csr.nodes, wc.csr.n, None, wc
            tok = wc.tokens
            if isinstance(tok, np.ndarray):
                tok = torch.from_numpy(tok).cuda()
            return wc.csr.nodes, wc.csr.n, tok, (wc if wc.stride == 0 else None)
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
        return nodes, len(nodes), torch.from_numpy(arr).cuda(), None

    def word2vec(self, sg=1, vector_size=8, window=1, min_count=1, negative=5,
                 workers=1, sample=0, **kwargs):
        """Set the model based on Word2Vec (skip-gram with negative sampling, trained on the GPU).

        Same signature and defaults as the reference (deepwalk.py:98-99); the remaining gensim
        defaults can be overridden by keyword: ``epochs`` (5), ``alpha`` (0.025), ``min_alpha``
        (0.0001), ``seed`` (1, weight init), ``ns_exponent`` (0.75), ``batch_words`` (10000),
        ``hs`` (0).  GPU-only knobs: ``lanes`` (number of concurrent workers, the analogue of
        gensim's ``workers``; 0 = library default, a function of the network only; 1 =
        sequential), ``unit_walks`` and ``policy`` ('chunk' | 'burst': how the corpus is cut into
        the units the workers take from the queue, see ``gw_sgns_plan``; 0 / None = library
        default; ``max_pieces`` caps the workers per start node of a 'burst' plan), ``reductions`` ('per_pair' | 'merged': see ``gw_sgns_flags`` in
        include/genewalk_b200.h), ``wait`` (False: only enqueue the training on the current CUDA
        stream; ``finish_word2vec()`` then downloads the vectors), ``sgns_seed`` (Philox key of
        the negative draws) and ``epoch_hook`` (see ``train_device``).
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
        unit_walks = int(kwargs.pop('unit_walks', 0))
        policy = kwargs.pop('policy', None)
        max_pieces = int(kwargs.pop('max_pieces', 0))
        reductions = kwargs.pop('reductions', 'per_pair')
        kwargs_async = not bool(kwargs.pop('wait', True))
        sgns_seed = kwargs.pop('sgns_seed', None)
        epoch_hook = kwargs.pop('epoch_hook', None)
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
                                          lanes=lanes, unit_walks=unit_walks, policy=policy,
                                          max_pieces=max_pieces,
                                          reductions=reductions, sgns_seed=sgns_seed,
                                          epoch_hook=epoch_hook)
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
        self.train_shape = {k: dv[k] for k in ('lanes', 'unit_walks', 'policy', 'regenerated')}
        nodes = dv['nodes']
        # plain device-to-host copies and host-side gathers: no torch kernel is launched here, so
        # nothing can trigger a lazy kernel load (which waits for an idle device) while other
        # replicates are training on their streams
        nv = int(dv['n_vocab'].item())
        order_h = dv['node_of_rank'].cpu().numpy()[:nv].astype(np.int64)
        vectors = dv['syn0'].cpu().numpy()[order_h]
        out_layer = dv['syn1neg'].cpu().numpy()[order_h]
        cnt = dv['counts'].cpu().numpy()[order_h]
        wv = NodeVectors([nodes[i] for i in order_h], vectors, counts=cnt)
        wv.node_ids = order_h.astype(np.int32)   # row r of wv.vectors is graph node node_ids[r]
        self.model = Word2VecModel(wv, syn1neg=out_layer, corpus_count=dv['W'], **self._w2v_args)
        return self.model

    def train_device(self, vector_size=8, negative=5, epochs=5, alpha=0.025, min_alpha=0.0001,
                     init_seed=1, ns_exponent=0.75, batch_words=10000, lanes=0, unit_walks=0,
                     policy=None, max_pieces=0, reductions='per_pair', sgns_seed=None, epoch_hook=None):
        """Enqueue vocabulary, alias table, init and the SGNS epochs on the current stream and
        return the device tensors (no host synchronisation).  ``epoch_hook(stage, when)`` is called
        with when in {'before', 'after'} around the vocabulary pass (stage 'count') and around every
        epoch's launch (stage = epoch number); bench.py records CUDA events there."""
        import ctypes
        import torch
        nodes, n, tokens, wc = self._device_corpus()
        if tokens is not None:
            L, W = int(tokens.shape[0]), int(tokens.shape[1])
            dev = tokens.device
        else:
            L, W = wc.wl, len(wc)
            dev = wc.csr.to_device().d_row_ptr.device
        d = int(vector_size)
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
        rng_rows = _init_rows(n, d, init_seed, dev)
        graph = wc.csr.to_device() if wc is not None else None
        if epoch_hook is not None:
            epoch_hook('count', 'before')
        if tokens is not None:
            _lib.call('gw_count_tokens', _lib.ptr(tokens), L * W, n, _lib.ptr(counts), st)
        else:
            _lib.call('gw_walk_count', n, graph.nnz, _lib.ptr(graph.d_row_ptr), _lib.ptr(graph.d_col_idx),
                      wc.niter, L, wc.seed, _lib.ptr(counts), st)
        if epoch_hook is not None:
            epoch_hook('count', 'after')
        _lib.call('gw_vocab_rank', n, _lib.ptr(counts), _lib.ptr(rank_of_node),
                  _lib.ptr(node_of_rank), _lib.ptr(n_vocab), st)
        _lib.call('gw_alias_build', n, _lib.ptr(counts), float(ns_exponent), _lib.ptr(alias), st)
        _lib.call('gw_sgns_init', n, d, _lib.ptr(rank_of_node), _lib.ptr(rng_rows), _lib.ptr(syn0),
                  _lib.ptr(syn1neg), st)
        job_walks = max(1, int(batch_words) // max(L, 1))
        a_l, a_u, a_p = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int32(0)
        md = getattr(graph, 'max_degree', None) if graph is not None else None
        _lib.call('gw_sgns_auto_shape', n, W, job_walks, -1 if md is None else int(md) * wc.niter,
                  ctypes.byref(a_l), ctypes.byref(a_u), ctypes.byref(a_p))
        if int(lanes) == 0:
            lanes = a_l.value
        if int(unit_walks) == 0:
            unit_walks = a_u.value
        policy_id = a_p.value if policy is None else {'chunk': 0, 'burst': 1}[policy]
        if graph is None:
            policy_id = 0          # no graph behind the corpus: plain chunks
        lib = _lib.load()
        cap = int(lib.gw_sgns_plan_capacity(n, W, int(unit_walks), policy_id))
        units = torch.empty((cap, 2), dtype=torch.int64, device=dev)    # gw_sgns_unit[cap]
        num_units = torch.zeros(1, dtype=torch.int64, device=dev)
        _lib.call('gw_sgns_plan', n, _lib.ptr(graph.d_row_ptr) if graph is not None else None,
                  wc.niter if wc is not None else 1, W, int(unit_walks), policy_id, int(max_pieces),
                  _lib.ptr(units), cap,
                  _lib.ptr(num_units), st)
        flags = {'per_pair': 0, 'merged': 4}[reductions]
        regen = tokens is None
        for ep in range(int(epochs)):
            if epoch_hook is not None:
                epoch_hook(ep, 'before')
            _lib.call('gw_sgns_train', n, d, None if regen else _lib.ptr(tokens),
                      _lib.ptr(graph.d_row_ptr) if regen else None,
                      _lib.ptr(graph.d_col_idx) if regen else None,
                      wc.niter if regen else 0, wc.seed if regen else 0, W, L, 1, int(negative),
                      int(epochs), ep, ep + 1, float(alpha), float(min_alpha), job_walks,
                      _lib.ptr(units), _lib.ptr(num_units), _lib.ptr(alias),
                      _lib.ptr(syn0), _lib.ptr(syn1neg), int(sgns_seed), int(lanes), flags, st)
            if epoch_hook is not None:
                epoch_hook(ep, 'after')
        return {'nodes': nodes, 'n': n, 'W': W, 'L': L, 'syn0': syn0, 'syn1neg': syn1neg,
                'counts': counts, 'rank_of_node': rank_of_node, 'node_of_rank': node_of_rank,
                'n_vocab': n_vocab, 'alias': alias, 'rng_rows': rng_rows, 'units': units,
                'num_units': num_units, 'lanes': int(lanes), 'unit_walks': int(unit_walks),
                'policy': ('chunk', 'burst')[policy_id], 'job_walks': job_walks,
                'regenerated': regen, 'h2d_bytes': n * d * 4}


_INIT_ROWS = {}


def _init_rows(n, d, seed, dev):
    """Device copy of gensim's init_weights rows, (default_rng(seed).random((n, d), f32)*2-1)/d;
    every replicate of a run uses the same matrix (gensim's seed is always 1 in GeneWalk,
    deepwalk.py:135-139), so it is generated and uploaded once per (n, d, seed, device)."""
    import torch
    key = (int(n), int(d), int(seed), str(dev))
    t = _INIT_ROWS.get(key)
    if t is None:
        rng = np.random.default_rng(seed=seed)
        rows = rng.random((n, d), dtype=np.float32)
        rows *= 2.0
        rows -= 1.0
        rows /= d
        if len(_INIT_ROWS) >= 8:
            _INIT_ROWS.clear()
        t = _INIT_ROWS[key] = torch.from_numpy(rows).to(dev)
        torch.cuda.current_stream().synchronize()   # other streams will read it
    return t


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
