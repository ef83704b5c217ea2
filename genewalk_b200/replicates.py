"""Replicate driver: the loops of /root/reference/genewalk/cli.py:200-214 (nreps_graph real-network
replicates) and :219-238 (nreps_null randomized-network replicates, pooled and sorted null
similarities), sharded over the GPUs and, on every GPU, run several at a time on CUDA streams.

Replicates are independent units: unit u runs on rank u % world_size with no
data-path collective; the only exchange is one all-gather of (a) the fixed-length gene-GO similarity
vector of every real replicate and (b) the variable-length null-similarity array of every null
replicate, after which every rank holds everything the statistics step needs.  With
``torch.distributed`` initialised on NCCL the gathers run over NVLink; the host logic is backend-
agnostic (the CPU tests drive it over gloo).
"""
import logging
import random

import numpy as np

from . import _lib
from .csr import GraphCSR, multigraph_degrees
from .deepwalk import DeepWalk, _csr_of
from .null_distributions import get_rand_graph
from .perform_statistics import GeneWalk

logger = logging.getLogger('genewalk.replicates')

KIND_REAL, KIND_NULL = 0, 1


def mix_seed(base_seed, stream, rep):
    """splitmix64 finaliser over (base_seed, stream, rep) -> independent 64-bit seeds."""
    m = 2 ** 64
    x = (base_seed * 0x9E3779B97F4A7C15 + stream * 0xBF58476D1CE4E5B9 +
         rep * 0x94D049BB133111EB + 1) % m
    x ^= x >> 30
    x = (x * 0xBF58476D1CE4E5B9) % m
    x ^= x >> 27
    x = (x * 0x94D049BB133111EB) % m
    x ^= x >> 31
    return x


def plan_units(nreps_graph, nreps_null, world_size=1, rank=0):
    """Units of this rank: [(kind, rep)], round-robin over real replicates then null replicates."""
    units = [(KIND_REAL, r) for r in range(nreps_graph)] + [(KIND_NULL, r) for r in range(nreps_null)]
    mine = [u for i, u in enumerate(units) if i % world_size == rank]
    # longest first: a randomized network has ~1.3x the adjacency entries (hence walks) of the real
    # one, so its replicates are started first and the shorter ones fill the tail
    return sorted(mine, key=lambda u: (-u[0], u[1]))


def owner_of(kind, rep, nreps_graph, world_size):
    idx = rep if kind == KIND_REAL else nreps_graph + rep
    return idx % world_size


# ---------------------------------------------------------------------------------------------
# collectives (work on CPU tensors over gloo and CUDA tensors over NCCL)
# ---------------------------------------------------------------------------------------------
def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def world():
    d = _dist()
    return (d.get_world_size(), d.get_rank()) if d is not None else (1, 0)


def gather_varlen(t, group=None):
    """All ranks get the list [t_rank0, t_rank1, ...] of 1-D tensors of different lengths."""
    import torch
    d = _dist()
    if d is None:
        return [t]
    ws = d.get_world_size(group)
    n = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(ws)]
    d.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    pad = torch.zeros(cap, dtype=t.dtype, device=t.device)
    pad[:t.numel()] = t
    out = [torch.empty_like(pad) for _ in range(ws)]
    d.all_gather(out, pad, group=group)
    return [o[:s] for o, s in zip(out, sizes)]


def gather_fixed(t, group=None):
    """All ranks get the list of same-shaped tensors [t_rank0, ...]."""
    import torch
    d = _dist()
    if d is None:
        return [t]
    out = [torch.empty_like(t) for _ in range(d.get_world_size(group))]
    d.all_gather(out, t.contiguous(), group=group)
    return out


def exchange_results(local_sims, local_null, nreps_graph, nreps_null, m, device, group=None):
    """local_sims: {rep: tensor[m]}, local_null: {rep: tensor[len]} of this rank's units ->
    (sims [nreps_graph, m], [null_rep0, null_rep1, ...]) on every rank."""
    import torch
    ws, rank = world()
    per_rank_real = (nreps_graph + ws - 1) // ws
    # fixed-length: slot k of a rank holds its k-th real replicate (zeros when it has none)
    mine = torch.zeros((max(per_rank_real, 1), m), dtype=torch.float32, device=device)
    my_real = sorted(local_sims)
    for k, rep in enumerate(my_real):
        mine[k] = local_sims[rep]
    gathered = gather_fixed(mine, group)
    sims = torch.zeros((nreps_graph, m), dtype=torch.float32, device=device)
    for rep in range(nreps_graph):
        own = owner_of(KIND_REAL, rep, nreps_graph, ws)
        k = [r for r in range(nreps_graph) if owner_of(KIND_REAL, r, nreps_graph, ws) == own].index(rep)
        sims[rep] = gathered[own][k]
    # variable-length: every rank concatenates its null arrays, lengths travel alongside
    my_null = sorted(local_null)
    lens = torch.tensor([local_null[r].numel() for r in my_null] or [0], dtype=torch.int64,
                        device=device)
    flat = torch.cat([local_null[r].reshape(-1) for r in my_null]) if my_null else \
        torch.zeros(0, dtype=torch.float32, device=device)
    all_flat = gather_varlen(flat, group)
    all_lens = gather_varlen(lens, group)
    nulls = [None] * nreps_null
    for rk in range(ws):
        reps = [r for r in range(nreps_null) if owner_of(KIND_NULL, r, nreps_graph, ws) == rk]
        off = 0
        for k, rep in enumerate(reps):
            ln = int(all_lens[rk][k].item())
            nulls[rep] = all_flat[rk][off:off + ln]
            off += ln
    return sims, nulls


# ---------------------------------------------------------------------------------------------
# the replicate bodies
# ---------------------------------------------------------------------------------------------
DEFAULT_CONCURRENT = 24   # upper bound of the replicates in flight per GPU


def fit_concurrent(requested, n, vector_size=8, walks=None, lanes=0, max_start_walks=-1):
    """Replicates to keep in flight at once: at most ``requested``, at most as many trainings as
    are RESIDENT on the GPU together with the workers each of them gets (``lanes``, default: the
    library's shape for an ``n``-node network) -- so that every training really runs with its full
    set of workers, whatever else shares the GPU, and a replicate's result does not depend on how
    many replicates run next to it -- and capped by the free HBM.  The walks are never stored (the
    trainer regenerates them), so a replicate in flight holds only its tables, vocabulary, noise
    table, unit plan and (null replicates) its randomized graph: a few MB at human scale."""
    import ctypes
    import torch
    lib = _lib.load()
    if not lanes:
        a_l, a_u, a_p = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int32(0)
        _lib.call('gw_sgns_auto_shape', int(n), int(walks) if walks else 100 * int(n), 1000,
                  int(max_start_walks), ctypes.byref(a_l), ctypes.byref(a_u), ctypes.byref(a_p))
        lanes = a_l.value
    resident = int(lib.gw_sgns_resident_trainings(int(vector_size), int(lanes)))
    per_replicate = int(n) * (2 * 4 * int(vector_size) + 64) + (64 << 20)
    free, _ = torch.cuda.mem_get_info()
    return max(1, min(int(requested), resident, int(0.8 * free) // per_replicate))


def real_replicate(csr, base_seed, rep, dim_rep=8, walk_length=10, niter=100, wait=True, **w2v):
    """cli.py:202-203: run_walks(MG.graph, workers, vector_size) -> DeepWalk.  wait=False only
    enqueues the work on the current CUDA stream; ``dw.finish_word2vec()`` completes it."""
    dw = DeepWalk(csr, walk_length, niter, seed=mix_seed(base_seed, 0, rep))
    dw.get_walks()
    dw.word2vec(vector_size=dim_rep, sgns_seed=mix_seed(base_seed, 1, rep), wait=wait, **w2v)
    return dw


class _NullJob(object):
    """A null replicate in flight: randomized graph, walks and training enqueued on a stream."""

    def __init__(self, mg, base_seed, rep, dim_rep, walk_length, niter, w2v, degrees=None):
        import torch
        self.dim_rep = dim_rep
        self.rg = get_rand_graph(mg, seed=mix_seed(base_seed, 2, rep), materialize=False,
                                 degrees=degrees)
        self.dw = DeepWalk(self.rg, walk_length, niter, seed=mix_seed(base_seed, 3, rep))
        self.dw.get_walks()
        if len(self.dw.walks) > 0:
            self.dw.word2vec(vector_size=dim_rep, sgns_seed=mix_seed(base_seed, 4, rep), wait=False,
                             **w2v)
        self.stream = torch.cuda.current_stream()

    def finish(self, keep_model=False):
        """Null similarities (device float32 tensor), cli.py:235 / null_distributions.py:33-48;
        keep_model=True also downloads the vectors into ``self.dw.model``."""
        import ctypes
        import torch
        dv = getattr(self.dw, '_pending', None)
        if dv is None:
            return torch.zeros(0, dtype=torch.float32, device='cuda')
        with torch.cuda.stream(self.stream):
            csr = self.rg.csr
            sims = torch.empty(csr.nnz, dtype=torch.float32, device=dv['syn0'].device)
            cnt = ctypes.c_int64(0)
            _lib.call('gw_null_similarities', csr.n, _lib.ptr(csr.d_row_ptr), _lib.ptr(csr.d_col_idx),
                      int(self.dim_rep), _lib.ptr(dv['syn0']), _lib.ptr(sims), ctypes.byref(cnt),
                      _lib.stream_ptr())
            out = sims[:cnt.value].clone()
            if keep_model:
                self.dw.finish_word2vec()
        self.dw._pending = None
        return out


def null_replicate(mg, base_seed, rep, dim_rep=8, walk_length=10, niter=100, **w2v):
    """cli.py:221-235: get_rand_graph -> run_walks -> get_null_distributions, all on the device.
    Returns (device float32 tensor of null similarities, DeepWalk)."""
    job = _NullJob(mg, base_seed, rep, dim_rep, walk_length, niter, w2v)
    return job.finish(), job.dw


_PRELOADED = set()
_SIDE_STREAMS = {}


def _side_streams(count):
    """The same stream objects on every call: torch's allocator caches blocks per stream, so a
    later batch reuses the corpora buffers of an earlier one instead of allocating anew."""
    import torch
    pool = _SIDE_STREAMS.setdefault(torch.cuda.current_device(), [])
    while len(pool) < count:
        pool.append(torch.cuda.Stream())
    return pool[:count]


def preload_kernels(dim_rep=8, **w2v):
    """Launch every kernel of a real and a null replicate once, on a 12-node network, while the
    device is idle.  CUDA loads a kernel at its first launch and that load waits until ALL streams
    have drained: with replicates in flight on several streams, the first randomized network and
    the first null-similarity call otherwise stall the host for seconds (measured: 6.6 s and 11.8 s
    bubbles in a 20-replicate run)."""
    import networkx as nx
    import torch
    key = (torch.cuda.current_device(), int(dim_rep), tuple(sorted(w2v.items())))
    if key in _PRELOADED:
        return
    g = nx.MultiGraph()
    g.add_edges_from([(i, (i + 1) % 12) for i in range(12)] + [(0, 6), (0, 6), (3, 3)])
    csr = GraphCSR.from_networkx(g).to_device()
    dw = real_replicate(csr, 0, 0, dim_rep, niter=2, wait=False, **w2v)
    pg = torch.zeros(1, dtype=torch.int32, device='cuda')
    one = torch.empty(1, dtype=torch.float32, device='cuda')
    _lib.call('gw_cosine_pairs', int(dim_rep), _lib.ptr(dw._pending['syn0']), 1, _lib.ptr(pg),
              _lib.ptr(pg), _lib.ptr(one), _lib.stream_ptr())
    dw.finish_word2vec()
    null = _NullJob(g, 0, 0, dim_rep, 10, 2, dict(w2v)).finish()
    if null.numel() > 0:
        _lib.call('gw_sort_f32', _lib.ptr(null), null.numel(), _lib.stream_ptr())
    torch.cuda.synchronize()
    _PRELOADED.add(key)


def _new_event(timing=False):
    import torch
    return torch.cuda.Event(enable_timing=timing)


def _run_rolling(n_units, nstreams, enqueue, finish, timeline=None):
    """Keeps up to ``nstreams`` units in flight, one per CUDA stream: whenever the unit of a stream
    has completed on the device (polled with CUDA events, so the host never blocks on one stream
    while another has gone idle), it is finished and the next unit -- units are handed out in
    order 0, 1, 2, ... -- is enqueued on that stream.  There is no barrier between rounds and at
    most ``nstreams`` replicates are alive.  ``enqueue(i)`` and ``finish(i, job)`` run with the
    unit's stream current.  ``timeline`` (a list) receives ``(unit, stream, begin_ms, end_ms)`` of
    the device work enqueued by ``enqueue``, measured with CUDA events relative to the first
    unit's begin."""
    import time
    import torch
    cur = torch.cuda.current_stream()
    nstreams = max(1, min(int(nstreams), n_units))
    streams = _side_streams(nstreams) if nstreams > 1 else [cur]
    for st in streams:
        if st is not cur:
            st.wait_stream(cur)
    inflight = {}
    events = []
    next_unit = 0

    def start(k):
        nonlocal next_unit
        i = next_unit
        next_unit += 1
        with torch.cuda.stream(streams[k]):
            e0 = _new_event(True) if timeline is not None else None
            if e0 is not None:
                e0.record()
            job = enqueue(i)
            e1 = _new_event(timeline is not None)
            e1.record()
            if timeline is not None:
                events.append((i, k, e0, e1))
        inflight[k] = (i, job, e1)

    for k in range(nstreams):
        start(k)
    while inflight:
        ready = [k for k in sorted(inflight, key=lambda k: inflight[k][0]) if inflight[k][2].query()]
        if not ready:
            time.sleep(0.0005)
            continue
        for k in ready:
            i, job, _ = inflight.pop(k)
            with torch.cuda.stream(streams[k]):
                finish(i, job)
            if next_unit < n_units:
                start(k)
    for st in streams:
        if st is not cur:
            cur.wait_stream(st)
    if timeline is not None and events:
        torch.cuda.synchronize()
        origin = events[0][2]
        timeline.extend((i, k, origin.elapsed_time(e0), origin.elapsed_time(e1))
                        for i, k, e0, e1 in events)


def run_walks_many(graph, n_replicates, base_seed=None, concurrent=DEFAULT_CONCURRENT,
                   walk_length=10, niter=100, vector_size=8, **w2v):
    """``[run_walks(graph, ...) for _ in range(n_replicates)]`` (cli.py:200-203) with up to
    ``concurrent`` replicates in flight on separate CUDA streams.  Returns the DeepWalk
    objects (``.model.wv`` set)."""
    _lib.require_cuda()
    if base_seed is None:
        base_seed = random.getrandbits(64)
    csr = _csr_of(graph).to_device()
    out = [None] * n_replicates          # replicates complete in any order, results keep theirs

    def enqueue(rep):
        return real_replicate(csr, base_seed, rep, vector_size, walk_length, niter, wait=False, **w2v)

    def finish(rep, dw):
        dw.finish_word2vec()
        out[rep] = dw
    nstreams = fit_concurrent(concurrent, csr.n, vector_size, niter * csr.nnz, w2v.get('lanes', 0))
    if min(nstreams, n_replicates) > 1:
        preload_kernels(vector_size, **w2v)
    _run_rolling(n_replicates, nstreams, enqueue, finish)
    return out


def null_replicates_many(mg, n_replicates, base_seed=None, concurrent=DEFAULT_CONCURRENT,
                         walk_length=10, niter=100, vector_size=8, keep_model=True, **w2v):
    """The body of the null_distribution loop (cli.py:219-236) for ``n_replicates`` randomized
    networks, up to ``concurrent`` at once: ``[(DeepWalk, null similarities as numpy float32)]``."""
    _lib.require_cuda()
    if base_seed is None:
        base_seed = random.getrandbits(64)
    out = [None] * n_replicates

    def enqueue(rep):
        return _NullJob(mg, base_seed, rep, vector_size, walk_length, niter, dict(w2v), degrees=degrees)

    def finish(rep, job):
        out[rep] = (job.dw, job.finish(keep_model=keep_model).cpu().numpy())
    degrees = multigraph_degrees(mg)
    nstreams = fit_concurrent(concurrent, len(degrees), vector_size, niter * int(degrees.sum()),
                              w2v.get('lanes', 0))
    if min(nstreams, n_replicates) > 1:
        preload_kernels(vector_size, **w2v)
    _run_rolling(n_replicates, nstreams, enqueue, finish)
    return out


def run_genewalk(mg, genes, nreps_graph=3, nreps_null=3, alpha_fdr=1, gene_id_type='hgnc_symbol',
                 dim_rep=8, base_seed=None, group=None, return_parts=False,
                 concurrent=DEFAULT_CONCURRENT, shard=True, walk_length=10, niter=100, **w2v):
    """node_vectors + null_distribution + statistics stages of cli.run_main (cli.py:194-252) for an
    assembled ``nx.MultiGraph``; replicates are sharded over the ranks of the initialised process
    group (one process per GPU) and, on every GPU, up to ``concurrent`` of them are trained at once
    on separate CUDA streams.  ``shard=False`` runs every replicate on the calling rank and uses no
    collective (independent runs on the GPUs of a node).  Returns the results DataFrame (identical
    on every rank)."""
    import time
    import torch
    _lib.require_cuda()
    ws, rank = world() if shard else (1, 0)
    t_start = time.perf_counter()
    if base_seed is None:
        base_seed = random.getrandbits(64)
        if ws > 1:   # all ranks must agree on the seed: rank 0's draw wins
            t = torch.tensor([base_seed & (2 ** 63 - 1)], dtype=torch.int64, device='cuda')
            _dist().broadcast(t, 0, group=group)
            base_seed = int(t.item())
    csr = _csr_of(mg).to_device()
    gw = GeneWalk(mg, genes, [], np.zeros(1, np.float32), gene_id_type=gene_id_type)
    pairs = gw.collect_pairs()
    m = len(pairs['pair_go'])
    local_sims, local_null, nvs = {}, {}, {}
    walks_trained, shapes = [], []
    units = plan_units(nreps_graph, nreps_null, ws, rank)
    degrees = multigraph_degrees(mg) if any(k == KIND_NULL for k, _ in units) else None
    t_setup = time.perf_counter()

    d_pg = torch.from_numpy(np.ascontiguousarray(pairs['pair_gene'], dtype=np.int32)).cuda()
    d_po = torch.from_numpy(np.ascontiguousarray(pairs['pair_go'], dtype=np.int32)).cuda()

    def enqueue(i):
        kind, rep = units[i]
        if kind == KIND_REAL:
            logger.info('real replicate %d/%d on rank %d' % (rep + 1, nreps_graph, rank))
            dw = real_replicate(csr, base_seed, rep, dim_rep, walk_length, niter, wait=False, **w2v)
            # gene-GO similarities straight from the device table (rows are graph node ids, the
            # index space of `pairs`; every node of a pair has an edge, hence walks, hence a row):
            # perform_statistics.py:102 without the round trip through host vectors
            sims = torch.empty(m, dtype=torch.float32, device='cuda')
            if m > 0:
                _lib.call('gw_cosine_pairs', int(dim_rep), _lib.ptr(dw._pending['syn0']), m,
                          _lib.ptr(d_pg), _lib.ptr(d_po), _lib.ptr(sims), _lib.stream_ptr())
            local_sims[rep] = sims
            walks_trained.append(dw._pending['W'])
            shapes.append((dw._pending['lanes'], dw._pending['unit_walks'], dw._pending['policy']))
            return dw
        logger.info('null replicate %d/%d on rank %d' % (rep + 1, nreps_null, rank))
        job = _NullJob(mg, base_seed, rep, dim_rep, walk_length, niter, dict(w2v), degrees=degrees)
        walks_trained.append(len(job.dw.walks))
        if job.dw._pending is not None:
            shapes.append((job.dw._pending['lanes'], job.dw._pending['unit_walks'], job.dw._pending['policy']))
        return job

    def finish(i, job):
        kind, rep = units[i]
        if kind == KIND_REAL:
            if return_parts:          # host copies of the vectors only when they are asked for
                job.finish_word2vec()
                nvs[rep] = job.model.wv
            job._pending = None
        else:
            local_null[rep] = job.finish()
    timeline = [] if return_parts else None
    nstreams = fit_concurrent(concurrent, csr.n, dim_rep, niter * csr.nnz, w2v.get('lanes', 0))
    if min(nstreams, len(units)) > 1:
        preload_kernels(dim_rep, **w2v)
    _run_rolling(len(units), nstreams, enqueue, finish, timeline)
    torch.cuda.synchronize()
    t_units = time.perf_counter()
    dev = torch.device('cuda')
    if ws > 1:
        sims, nulls = exchange_results(local_sims, local_null, nreps_graph, nreps_null, m, dev, group)
    else:
        sims = torch.stack([local_sims[r] for r in range(nreps_graph)]) if nreps_graph else \
            torch.zeros((0, m), dtype=torch.float32, device=dev)
        nulls = [local_null[r] for r in range(nreps_null)]
    pooled = torch.cat(nulls) if nulls else torch.zeros(0, dtype=torch.float32, device=dev)
    if pooled.numel() > 0:   # cli.py:238: srd = np.asarray(sorted(srd))
        pooled = pooled.contiguous()
        _lib.call('gw_sort_f32', _lib.ptr(pooled), pooled.numel(), _lib.stream_ptr())
    gw.srd = pooled.cpu().numpy()
    gw._d_srd = pooled
    gw.nvs = [nvs.get(r) for r in range(nreps_graph)]
    t_exchange = time.perf_counter()
    df = gw.generate_output(alpha_fdr=alpha_fdr, sims=sims, pairs=pairs)
    t_end = time.perf_counter()
    timings = {'setup_s': t_setup - t_start, 'replicates_s': t_units - t_setup,
               'exchange_s': t_exchange - t_units, 'statistics_s': t_end - t_exchange}
    logger.info('run_genewalk timings: %s' % timings)
    if return_parts:
        return df, {'sims': sims, 'null': gw.srd, 'nvs': nvs, 'base_seed': base_seed,
                    'timings': timings, 'walks_trained_local': int(sum(walks_trained)), 'train_shapes_local': shapes,
                    'timeline': [(units[i], k, b, e) for i, k, b, e in timeline]}
    return df
