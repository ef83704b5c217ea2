"""Randomized networks and null similarity distributions on the GPU.

Mirrors /root/reference/genewalk/null_distributions.py: ``get_rand_graph(mg)`` and
``get_null_distributions(rg, nv)``.  gw_config_model + gw_csr_from_edges replace
``nx.configuration_model`` and the adjacency it yields (null_distributions.py:23-29);
gw_null_similarities replaces the per-node ``nv.distances`` loop (null_distributions.py:39-47).
"""
import logging
import random

import networkx as nx
import numpy as np

from . import _lib
from .csr import GraphCSR, multigraph_degrees

logger = logging.getLogger('genewalk.get_null_distributions')


class DeviceGraph(object):
    """A generated graph that lives on the device: what ``get_rand_graph(mg, materialize=False)``
    returns.  ``run_walks`` and ``get_null_distributions`` accept it in place of the MultiGraph;
    ``to_networkx()`` builds the ``nx.MultiGraph`` the reference would have returned."""

    def __init__(self, csr, edge_u, edge_v):
        self.csr = csr
        self.edge_u = edge_u      # torch int32 [S/2] (pairing order, parallel edges/loops kept)
        self.edge_v = edge_v
        self.graph = {'_gw_csr': csr}

    def number_of_nodes(self):
        return self.csr.n

    def number_of_edges(self):
        return int(self.edge_u.numel())

    def nodes(self):
        return list(self.csr.nodes)

    def to_networkx(self):
        rg = nx.MultiGraph()
        rg.add_nodes_from(self.csr.nodes)
        names = self.csr.nodes
        u = self.edge_u.cpu().numpy()
        v = self.edge_v.cpu().numpy()
        rg.add_edges_from((names[a], names[b]) for a, b in zip(u.tolist(), v.tolist()))
        rg.graph['_gw_csr'] = self.csr
        return rg


def get_rand_graph(mg, seed=None, materialize=True, degrees=None):
    """Return a random graph with the same degree distribution as the input.

    Parameters
    ----------
    mg : networkx.MultiGraph
        An input graph based on which a random graph is generated.
    seed : Optional[int]
        64-bit seed of the Philox-keyed stub shuffle; default ``random.getrandbits(64)`` so that
        Python's ``random.seed`` governs it like the reference's ``configuration_model``.
    materialize : bool
        True (default, API parity): return an ``nx.MultiGraph`` with nodes ``'n0'..'n{N-1}'``,
        self-loops and parallel edges kept; the device CSR rides along in ``rg.graph['_gw_csr']``
        so ``run_walks``/``get_null_distributions`` do not re-extract it.  False: return the
        ``DeviceGraph`` only (no Python edge objects are created).
    degrees : Optional[numpy.ndarray]
        ``[mg.degree(n) for n in mg.nodes()]`` when the caller already has it (the replicate
        drivers randomize the same network many times).

    Returns
    -------
    networkx.MultiGraph
        A random graph whose degree distribution matches that of the input.
    """
    import torch
    _lib.require_cuda()
    # this is not randomized: order is same (null_distributions.py:22-23)
    d_seq = np.sort(multigraph_degrees(mg) if degrees is None else np.asarray(degrees))[::-1].astype(np.int32)
    n = len(d_seq)
    S = int(d_seq.astype(np.int64).sum())
    if S % 2 != 0:
        raise nx.NetworkXError('Invalid degree sequence: sum of degrees must be even, not odd')
    seed = random.getrandbits(64) if seed is None else int(seed) & (2 ** 64 - 1)
    dev = torch.device('cuda')
    st = _lib.stream_ptr()
    d_deg = torch.from_numpy(np.ascontiguousarray(d_seq)).to(dev)
    half = S // 2
    edge_u = torch.empty(half, dtype=torch.int32, device=dev)
    edge_v = torch.empty(half, dtype=torch.int32, device=dev)
    row_ptr = torch.zeros(n + 1, dtype=torch.int32, device=dev)
    col_idx = torch.empty(max(S, 1), dtype=torch.int32, device=dev)
    if n > 0 and S > 0:
        _lib.call('gw_config_model', n, _lib.ptr(d_deg), S, seed, _lib.ptr(edge_u),
                  _lib.ptr(edge_v), st)
    import ctypes
    nnz = ctypes.c_int64(0)
    if n > 0:
        _lib.call('gw_csr_from_edges', n, half, _lib.ptr(edge_u), _lib.ptr(edge_v),
                  _lib.ptr(row_ptr), _lib.ptr(col_idx), ctypes.byref(nnz), st)
    # the node labels are numbers which gives problems in word2vec so adjust 0 to n0 (:26-29)
    nodes = ['n%s' % i for i in range(n)]
    # multigraph degree of the first (largest) node: an upper bound of its distinct neighbours
    csr = GraphCSR.from_device(nodes, row_ptr, col_idx, nnz.value,
                               max_degree=int(d_seq[0]) if n > 0 else 0)
    dg = DeviceGraph(csr, edge_u, edge_v)
    return dg.to_networkx() if materialize else dg


def get_null_distributions(rg, nv):
    """Return a distribution with similarity values between (random) node
    vectors originating from the input randomized graph.

    Returns a float32 numpy array (the reference returns a list of numpy float32; both support
    ``srd += ...`` accumulation through ``list(...)`` -- see ``NullDistribution`` in replicates.py
    for the pooled, device-sorted form).
    """
    import ctypes
    import torch
    from .deepwalk import _csr_of
    _lib.require_cuda()
    csr = _csr_of(rg).to_device()
    if csr.nnz == 0:
        return []
    dev = csr.d_row_ptr.device
    # vectors in node-id order; nodes missing from the vocabulary have no neighbours
    d = nv.vector_size
    full = np.zeros((csr.n, d), dtype=np.float32)
    ids = getattr(nv, 'node_ids', None)
    if ids is None or len(ids) != len(nv.index_to_key):
        idx = csr.index
        ids = np.fromiter((idx[k] for k in nv.index_to_key), dtype=np.int64,
                          count=len(nv.index_to_key))
    ids = np.asarray(ids, dtype=np.int64)
    full[ids] = nv.vectors
    # a node that has neighbours but no vector: the reference's nv.distances raises KeyError
    # (null_distributions.py:46); zero rows here would put NaN similarities into the null
    have = torch.zeros(csr.n, dtype=torch.bool, device=dev)
    have[torch.from_numpy(ids).to(dev)] = True
    lost = torch.nonzero(~have & (csr.d_row_ptr[1:] > csr.d_row_ptr[:-1])).reshape(-1)
    if lost.numel() > 0:
        raise KeyError("Key '%s' not present" % (csr.nodes[int(lost[0])],))
    vec = torch.from_numpy(full).to(dev)
    sims = torch.empty(csr.nnz, dtype=torch.float32, device=dev)
    cnt = ctypes.c_int64(0)
    _lib.call('gw_null_similarities', csr.n, _lib.ptr(csr.d_row_ptr), _lib.ptr(csr.d_col_idx), d,
              _lib.ptr(vec), _lib.ptr(sims), ctypes.byref(cnt), _lib.stream_ptr())
    return list(sims[:cnt.value].cpu().numpy())
