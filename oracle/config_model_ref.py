"""numpy restatement of get_rand_graph (oracle; TEST INFRASTRUCTURE ONLY).

Follows /root/reference/genewalk/null_distributions.py:9-30 and the networkx generator it calls
(networkx/generators/degree_seq.py: _to_stublist / _configuration_model): degree sequence sorted
descending, stub list [i]*d_i, uniform shuffle, first half paired with second half, self-loops and
parallel edges kept, nodes relabelled 'n<i>'.  The reference shuffles with CPython's MT19937; the
build replaces it with a Philox-keyed sort (north_star item 3), and this file is the CPU statement of
exactly that shuffle, so the CUDA randomizer must reproduce its edge list bit for bit.

Stub k gets the 64-bit key (x0 << 32 | x1) of Philox(ctr={lo32(k), hi32(k), 0, 0x43}, key=seed);
the permutation is the stable ascending argsort of the keys.
"""
import numpy as np

from .philox import philox4x32_10


def multigraph_degrees_desc(degrees):
    """d_seq = sorted([mg.degree(n) ...], reverse=True)  (null_distributions.py:23)."""
    return np.sort(np.asarray(degrees, dtype=np.int64))[::-1].astype(np.int32)


def shuffle_keys(S, seed):
    k = np.arange(S, dtype=np.uint64)
    x0, x1, _, _ = philox4x32_10(k & np.uint64(0xFFFFFFFF), k >> np.uint64(32), 0, 0x43,
                                 seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return (x0.astype(np.uint64) << np.uint64(32)) | x1.astype(np.uint64)


def config_model_edges(d_seq, seed):
    """Returns (u, v) int32 arrays of the S/2 edges in pairing order."""
    d_seq = np.asarray(d_seq, dtype=np.int64)
    S = int(d_seq.sum())
    if S % 2 != 0:
        raise ValueError('Invalid degree sequence: sum of degrees must be even, not odd')
    owner = np.repeat(np.arange(len(d_seq), dtype=np.int32), d_seq)
    perm = np.argsort(shuffle_keys(S, seed), kind='stable')
    shuffled = owner[perm]
    half = S // 2
    return shuffled[:half].copy(), shuffled[half:].copy()


def csr_distinct(n, u, v):
    """Distinct-neighbour CSR (rows sorted by neighbour id); a self-loop lists the node once."""
    a = np.concatenate([u, v]).astype(np.int64)
    b = np.concatenate([v, u]).astype(np.int64)
    key = np.unique(a * n + b)
    src = (key // n).astype(np.int32)
    dst = (key % n).astype(np.int32)
    row_ptr = np.zeros(n + 1, dtype=np.int32)
    np.add.at(row_ptr, src + 1, 1)
    row_ptr = np.cumsum(row_ptr).astype(np.int32)
    return row_ptr, dst
