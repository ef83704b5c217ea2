"""CSR graph builder fed from the networkx graph (north_star subsystem 1).

``GraphCSR`` holds the distinct-neighbour adjacency of an ``nx.MultiGraph`` in the exact order
networkx exposes it -- ``col_idx[row_ptr[v] + k] == list(graph[v])[k]`` -- which is the object
``random.choice(list(graph[node]))`` samples from in the reference
(/root/reference/genewalk/deepwalk.py:165) and whose length defines the number of walks per node
(deepwalk.py:173,197).  Parallel edges collapse; a self-loop lists the node once.
"""
import numpy as np


class GraphCSR(object):
    """Host + device CSR of a graph.

    Attributes
    ----------
    nodes : list            node labels, ``list(nx.nodes(graph))`` order (index = node id)
    index : dict            label -> node id
    row_ptr, col_idx : np.ndarray (int32)   host copies
    d_row_ptr, d_col_idx : torch.Tensor     device copies (created lazily by ``to_device``)
    """

    def __init__(self, nodes, row_ptr, col_idx, index=None):
        self.nodes = nodes
        self._index = index
        self.row_ptr = None if row_ptr is None else np.ascontiguousarray(row_ptr, dtype=np.int32)
        self.col_idx = None if col_idx is None else np.ascontiguousarray(col_idx, dtype=np.int32)
        self.d_row_ptr = None
        self.d_col_idx = None
        self._nnz = None if self.row_ptr is None else int(self.row_ptr[-1])
        # largest number of distinct neighbours (or an upper bound of it), None = unknown: the
        # trainer's default shape looks at the longest run of walks with one start node
        self.max_degree = None if self.row_ptr is None or len(self.row_ptr) < 2 else \
            int(np.diff(self.row_ptr).max())

    # -- construction ---------------------------------------------------------------------------
    @classmethod
    def from_networkx(cls, graph):
        nodes = list(graph.nodes())
        index = {v: i for i, v in enumerate(nodes)}
        n = len(nodes)
        adj = getattr(graph, '_adj', None) or graph.adj   # raw dict-of-dicts (insertion order)
        deg = np.fromiter((len(adj[v]) for v in nodes), dtype=np.int64, count=n)
        row_ptr = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(deg, out=row_ptr[1:])
        if row_ptr[-1] >= 2 ** 31:
            raise ValueError('graph has too many adjacency entries for int32 CSR')
        col_idx = np.fromiter((index[w] for v in nodes for w in adj[v]), dtype=np.int32,
                              count=int(row_ptr[-1]))
        return cls(nodes, row_ptr.astype(np.int32), col_idx, index)

    @classmethod
    def from_device(cls, nodes, d_row_ptr, d_col_idx, nnz, max_degree=None):
        g = cls(nodes, None, None)
        g.d_row_ptr = d_row_ptr
        g.d_col_idx = d_col_idx
        g._nnz = int(nnz)
        g.max_degree = None if max_degree is None else int(max_degree)
        return g

    # -- accessors --------------------------------------------------------------------------------
    @property
    def n(self):
        return len(self.nodes)

    @property
    def nnz(self):
        return self._nnz

    @property
    def index(self):
        if self._index is None:
            self._index = {v: i for i, v in enumerate(self.nodes)}
        return self._index

    def to_device(self, device=None):
        import torch
        if self.d_row_ptr is None:
            dev = torch.device('cuda' if device is None else device)
            self.d_row_ptr = torch.from_numpy(self.row_ptr).to(dev, non_blocking=False)
            self.d_col_idx = torch.from_numpy(self.col_idx).to(dev, non_blocking=False)
        return self

    def to_host(self):
        if self.row_ptr is None:
            self.row_ptr = self.d_row_ptr.cpu().numpy()
            self.col_idx = self.d_col_idx[:self._nnz].cpu().numpy()
        return self


def multigraph_degrees(graph):
    """``[mg.degree(n) for n in mg.nodes()]``: parallel edges counted, a self-loop counts 2
    (/root/reference/genewalk/null_distributions.py:23)."""
    if hasattr(graph, 'degrees') and hasattr(graph, 'csr'):    # edgelist.EdgeListNetwork
        return np.asarray(graph.degrees, dtype=np.int64)
    adj = getattr(graph, '_adj', None) or graph.adj   # raw dict-of-dicts: no per-node view objects
    if graph.is_multigraph():   # same numbers as graph.degree()
        it = (sum(map(len, nbrs.values())) + (len(nbrs[v]) if v in nbrs else 0)
              for v, nbrs in adj.items())
    else:
        it = (len(nbrs) + (1 if v in nbrs else 0) for v, nbrs in adj.items())
    return np.fromiter(it, dtype=np.int64, count=graph.number_of_nodes())


def coprime_stride(A, frac=0.6180339887498949):
    """A stride close to frac*A that is coprime to A (slot -> adjacency-entry permutation)."""
    from math import gcd
    if A <= 2:
        return 1
    s = max(1, int(A * frac))
    while gcd(s, A) != 1:
        s += 1
        if s >= A:
            s = 1
    return s
