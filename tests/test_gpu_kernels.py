"""GPU parity tests of the C ABI against the CPU oracle (bit-exact for integer/index work; float
tolerances are written at each assert).  Everything here calls libgenewalk_b200.so through ctypes."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')

from genewalk_b200 import _lib  # noqa: E402
from oracle import lib as olib  # noqa: E402
from oracle import config_model_ref, stats_ref  # noqa: E402
from genewalk_b200 import workloads as graphs  # noqa: E402


_KEEP = []   # device tensors whose raw pointers were handed to the library stay alive per test


def dev(a, dtype=None):
    a = np.ascontiguousarray(a, dtype=dtype)
    t = torch.from_numpy(a).cuda()
    _KEEP.append(t)
    return t


@pytest.fixture(autouse=True)
def _release_device_tensors():
    yield
    torch.cuda.synchronize()
    del _KEEP[:]


def st():
    return _lib.stream_ptr()


# ------------------------------------------------------------------------------------------------
# walks
# ------------------------------------------------------------------------------------------------
def gpu_walks(row_ptr, col_idx, niter, L, seed, stride=1, begin=0, end=None):
    n = len(row_ptr) - 1
    A = int(row_ptr[-1])
    W = niter * A
    end = W if end is None else end
    tok = torch.empty((L, end - begin), dtype=torch.int32, device='cuda')
    _lib.call('gw_walk_uniform', n, A, _lib.ptr(dev(row_ptr, np.int32)),
              _lib.ptr(dev(col_idx, np.int32)), niter, L, seed, stride, begin, end, _lib.ptr(tok),
              st())
    torch.cuda.synchronize()
    return tok.cpu().numpy()


@pytest.mark.parametrize('n,m,niter,L,stride', [(50, 200, 7, 10, 1), (300, 2000, 5, 10, None),
                                                (1000, 5000, 3, 13, None), (10, 12, 100, 2, 1),
                                                (200, 900, 6, 10, 0),
                                                (64, 300, 4, 1, 1)])
def test_walks_bit_exact(n, m, niter, L, stride):
    from genewalk_b200.csr import coprime_stride
    row_ptr, col_idx = graphs.csr_arrays(n, m, seed=n + m, self_loops=3)
    A = int(row_ptr[-1])
    stride = coprime_stride(A) if stride is None else stride
    seed = 0x1234567890ABCDEF ^ n
    got = gpu_walks(row_ptr, col_idx, niter, L, seed, stride)
    want = olib.walk_uniform(row_ptr, col_idx, niter, L, seed, stride)
    assert got.shape == want.shape == (L, niter * A)
    assert np.array_equal(got, want)


def test_walks_slot_window_and_large():
    # a sub-range of slots of a larger problem equals the oracle on the same range
    row_ptr, col_idx = graphs.csr_arrays(20000, 300000, seed=5)
    A = int(row_ptr[-1])
    from genewalk_b200.csr import coprime_stride
    stride = coprime_stride(A)
    W = 100 * A
    b, e = W // 3, W // 3 + 50000
    got = gpu_walks(row_ptr, col_idx, 100, 10, 99, stride, b, e)
    want = olib.walk_uniform(row_ptr, col_idx, 100, 10, 99, stride, b, e)
    assert np.array_equal(got, want)


def test_walks_structure_and_chi_square():
    """Start-node multiplicity niter*deg (deepwalk.py:173,197); every step lands on a neighbour;
    transition frequencies out of a hub are uniform over its distinct neighbours (chi-square)."""
    row_ptr, col_idx = graphs.csr_arrays(400, 3000, seed=11, self_loops=5)
    n, A, niter, L = 400, int(row_ptr[-1]), 50, 10
    tok = gpu_walks(row_ptr, col_idx, niter, L, 2024, 1)
    deg = np.diff(row_ptr)
    starts = np.bincount(tok[0], minlength=n)
    assert np.array_equal(starts, niter * deg)
    nbr = set(zip(np.repeat(np.arange(n), deg).tolist(), col_idx.tolist()))
    a, b = tok[:-1].ravel(), tok[1:].ravel()
    assert all((x, y) in nbr for x, y in zip(a[:200000].tolist(), b[:200000].tolist()))
    hub = int(np.argmax(deg))
    nxt = b[a == hub]
    obs = np.bincount(nxt, minlength=n)[col_idx[row_ptr[hub]:row_ptr[hub + 1]]]
    exp = len(nxt) / deg[hub]
    chi2 = ((obs - exp) ** 2 / exp).sum()
    from scipy.stats import chi2 as chi2_dist
    assert chi2_dist.sf(chi2, deg[hub] - 1) > 1e-4


def test_walks_bad_arguments():
    row_ptr, col_idx = graphs.csr_arrays(50, 200, seed=1)
    A = int(row_ptr[-1])
    with pytest.raises(ValueError):
        gpu_walks(row_ptr, col_idx, 0, 10, 1)
    with pytest.raises(ValueError):
        gpu_walks(row_ptr, col_idx, 2, 10, 1, stride=A)   # stride out of range


# ------------------------------------------------------------------------------------------------
# sort / scan based pieces
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('count', [0, 1, 2, 255, 4096, 4097, 100003, 3000000])
def test_sort_f32(count):
    rng = np.random.default_rng(count)
    x = (rng.standard_normal(count) * 3).astype(np.float32)
    if count > 10:
        x[:5] = [0.0, -0.0, 1.0, 1.0, -1.0]
    d = dev(x)
    _lib.call('gw_sort_f32', _lib.ptr(d), count, st())
    got = d.cpu().numpy()
    want = np.sort(x, kind='stable')
    assert np.array_equal(got, want)


@pytest.mark.parametrize('n,S2', [(5, 4), (100, 500), (3000, 40000), (40000, 1000000)])
def test_config_model_bit_exact(n, S2):
    rng = np.random.default_rng(n)
    deg = rng.zipf(1.7, size=n).astype(np.int64)
    deg = np.minimum(deg, n)
    deg[0] += (S2 * 2 - deg.sum()) if deg.sum() < S2 * 2 else 0
    if deg.sum() % 2:
        deg[0] += 1
    d_seq = config_model_ref.multigraph_degrees_desc(deg)
    S = int(d_seq.astype(np.int64).sum())
    seed = 0xC0FFEE + n
    eu = torch.empty(S // 2, dtype=torch.int32, device='cuda')
    ev = torch.empty(S // 2, dtype=torch.int32, device='cuda')
    _lib.call('gw_config_model', n, _lib.ptr(dev(d_seq, np.int32)), S, seed, _lib.ptr(eu),
              _lib.ptr(ev), st())
    u, v = config_model_ref.config_model_edges(d_seq, seed)
    assert np.array_equal(eu.cpu().numpy(), u)
    assert np.array_equal(ev.cpu().numpy(), v)
    # degree sequence is preserved exactly (loops count twice), as nx.configuration_model does
    got_deg = np.bincount(np.concatenate([u, v]), minlength=n)
    assert np.array_equal(got_deg, d_seq)
    # CSR of distinct neighbours
    row_ptr = torch.zeros(n + 1, dtype=torch.int32, device='cuda')
    col = torch.empty(S, dtype=torch.int32, device='cuda')
    nnz = ctypes.c_int64(0)
    _lib.call('gw_csr_from_edges', n, S // 2, _lib.ptr(eu), _lib.ptr(ev), _lib.ptr(row_ptr),
              _lib.ptr(col), ctypes.byref(nnz), st())
    rp, ci = config_model_ref.csr_distinct(n, u, v)
    assert nnz.value == rp[-1]
    assert np.array_equal(row_ptr.cpu().numpy(), rp)
    assert np.array_equal(col[:nnz.value].cpu().numpy(), ci)


def test_config_model_odd_sum_raises():
    d = dev(np.array([2, 1], dtype=np.int32))
    e = torch.empty(1, dtype=torch.int32, device='cuda')
    with pytest.raises(ValueError):
        _lib.call('gw_config_model', 2, _lib.ptr(d), 3, 1, _lib.ptr(e), _lib.ptr(e), st())


def test_csr_from_edges_empty_rows_and_loops():
    n = 9
    u = np.array([1, 1, 4, 4, 7, 2], dtype=np.int32)
    v = np.array([2, 2, 4, 6, 1, 1], dtype=np.int32)   # parallel edge 1-2 x3, self-loop at 4
    row_ptr = torch.zeros(n + 1, dtype=torch.int32, device='cuda')
    col = torch.empty(2 * len(u), dtype=torch.int32, device='cuda')
    nnz = ctypes.c_int64(0)
    _lib.call('gw_csr_from_edges', n, len(u), _lib.ptr(dev(u)), _lib.ptr(dev(v)),
              _lib.ptr(row_ptr), _lib.ptr(col), ctypes.byref(nnz), st())
    rp, ci = config_model_ref.csr_distinct(n, u, v)
    assert np.array_equal(row_ptr.cpu().numpy(), rp)
    assert np.array_equal(col[:nnz.value].cpu().numpy(), ci)
    assert nnz.value == 7     # 1:{2,7} 2:{1} 4:{4,6} 6:{4} 7:{1}


# ------------------------------------------------------------------------------------------------
# vocabulary / alias
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('n,count', [(10, 1000), (5000, 2000003), (60000, 3000000)])
def test_count_tokens_and_vocab_rank(n, count):
    rng = np.random.default_rng(n)
    w = (np.arange(1, n + 1, dtype=np.float64)) ** -1.1
    tok = rng.choice(n, size=count, p=w / w.sum()).astype(np.int32)
    tok[::1000] = -1                       # padding of ragged walks is ignored
    d_tok = dev(tok)
    counts = torch.empty(n, dtype=torch.int64, device='cuda')
    _lib.call('gw_count_tokens', _lib.ptr(d_tok), count, n, _lib.ptr(counts), st())
    want = np.bincount(tok[tok >= 0], minlength=n)
    assert np.array_equal(counts.cpu().numpy(), want)
    rank = torch.empty(n, dtype=torch.int32, device='cuda')
    node = torch.empty(n, dtype=torch.int32, device='cuda')
    nv = torch.zeros(1, dtype=torch.int32, device='cuda')
    _lib.call('gw_vocab_rank', n, _lib.ptr(counts), _lib.ptr(rank), _lib.ptr(node), _lib.ptr(nv), st())
    order = np.argsort(-want, kind='stable')       # descending count, ties by node id
    V = int((want > 0).sum())
    assert int(nv.item()) == V
    assert np.array_equal(node.cpu().numpy()[:V], order[:V])
    r = rank.cpu().numpy()
    assert np.array_equal(r[order[:V]], np.arange(V))
    assert (r[want == 0] == -1).all()


@pytest.mark.parametrize('n', [1, 2, 7, 1000, 38000, 300000])
def test_alias_table_reproduces_distribution(n):
    rng = np.random.default_rng(n)
    counts = rng.zipf(1.5, size=n).astype(np.int64) * 10
    if n > 5:
        counts[rng.integers(0, n, size=max(1, n // 50))] = 0     # nodes absent from the vocab
        counts[:3] = [10 ** 9, 1, 1]
    d_counts = dev(counts)
    table = torch.empty((n, 2), dtype=torch.int32, device='cuda')
    _lib.call('gw_alias_build', n, _lib.ptr(d_counts), 0.75, _lib.ptr(table), st())
    t = table.cpu().numpy()
    prob = t[:, 0].copy().view(np.float32).astype(np.float64)
    alias = t[:, 1]
    assert ((prob >= 0) & (prob <= 1)).all() and ((alias >= 0) & (alias < n)).all()
    implied = olib.alias_implied_distribution(prob, alias)
    w = np.where(counts > 0, counts.astype(np.float64) ** 0.75, 0.0)
    target = w / w.sum()
    # float32 prob storage: each slot's split is exact to 2^-24 of a slot (1/n); an item collects
    # one such rounding from its own slot and from every slot aliased to it
    refs = 1 + np.bincount(alias, minlength=n)
    assert (np.abs(implied - target) <= refs * (2.0 ** -24 / n) + 1e-15).all()
    assert abs(implied.sum() - 1.0) < 1e-9
    assert implied[counts == 0].max(initial=0.0) == 0.0      # absent nodes are never drawn


# ------------------------------------------------------------------------------------------------
# SGNS: 1-thread CUDA run is bit-exact against the oracle's device-sampler mode
# ------------------------------------------------------------------------------------------------
def sgns_case(n, m, niter, L, d, K, epochs, seed):
    row_ptr, col_idx = graphs.csr_arrays(n, m, seed=seed, self_loops=2)
    tok = olib.walk_uniform(row_ptr, col_idx, niter, L, seed)
    counts = np.bincount(tok.ravel(), minlength=n)
    prob, alias = olib.alias_table(counts)
    syn0, syn1 = olib.gensim_init_weights(n, d, seed=1)
    return tok, prob, alias, syn0, syn1


def gpu_plan(n, W, unit_walks, policy, row_ptr=None, niter=1, max_pieces=0):
    """gw_sgns_plan -> (device units tensor, device num_units, host copy of the units)."""
    lib = _lib.load()
    cap = int(lib.gw_sgns_plan_capacity(n, W, unit_walks, policy))
    units = torch.empty((cap, 2), dtype=torch.int64, device='cuda')
    num = torch.zeros(1, dtype=torch.int64, device='cuda')
    d_rp = dev(row_ptr, np.int32) if row_ptr is not None else None
    _lib.call('gw_sgns_plan', n, _lib.ptr(d_rp), niter, W, unit_walks, policy, max_pieces, _lib.ptr(units), cap,
              _lib.ptr(num), st())
    torch.cuda.synchronize()
    k = int(num.item())
    raw = units[:k].cpu().numpy()
    host = np.zeros(k, dtype=[('first', np.int64), ('count', np.int32), ('node', np.int32)])
    host['first'] = raw[:, 0]
    host['count'] = (raw[:, 1] & 0xffffffff).astype(np.int64).astype(np.int32)
    host['node'] = (raw[:, 1] >> 32).astype(np.int32)
    return units, num, host


def gpu_sgns(tok, n, d, K, epochs, prob, alias, syn0, syn1, seed, lanes, job_walks=1000,
             alpha0=0.025, alpha_min=0.0001, unit_walks=0, policy=0, flags=0, graph=None):
    """graph = (row_ptr, col_idx, niter, walk_seed): the trainer regenerates the walks (tok is then
    only used for its shape); otherwise it reads the token array."""
    L, W = tok.shape
    table = np.empty((n, 2), dtype=np.int32)
    table[:, 0] = prob.view(np.int32)
    table[:, 1] = alias
    d0, d1 = dev(syn0), dev(syn1)
    if unit_walks == 0:
        unit_walks = job_walks if lanes == 1 else 8
    row_ptr = graph[0] if graph is not None else None
    units, num, _ = gpu_plan(n, W, unit_walks, policy, row_ptr, graph[2] if graph is not None else 1)
    if graph is None:
        args = (_lib.ptr(dev(tok, np.int32)), None, None, 0, 0)
    else:
        args = (None, _lib.ptr(dev(graph[0], np.int32)), _lib.ptr(dev(graph[1], np.int32)), graph[2], graph[3])
    _lib.call('gw_sgns_train', n, d, *args, W, L, 1, K, epochs, 0, epochs, alpha0, alpha_min, job_walks,
              _lib.ptr(units), _lib.ptr(num), _lib.ptr(dev(table)), _lib.ptr(d0), _lib.ptr(d1),
              seed, lanes, flags, st())
    torch.cuda.synchronize()
    return d0.cpu().numpy(), d1.cpu().numpy()


@pytest.mark.parametrize('n,m,niter,L,d,K,epochs', [(40, 120, 3, 10, 8, 5, 5),
                                                    (40, 120, 2, 10, 16, 5, 2),
                                                    (30, 90, 2, 7, 8, 3, 2),
                                                    (30, 90, 2, 5, 32, 2, 1),
                                                    (25, 60, 2, 10, 16, 7, 1),
                                                    (25, 60, 1, 6, 64, 5, 1),
                                                    (25, 60, 1, 6, 128, 5, 1),
                                                    (25, 60, 1, 6, 64, 3, 1)])
def test_sgns_sequential_bit_exact(n, m, niter, L, d, K, epochs):
    tok, prob, alias, syn0, syn1 = sgns_case(n, m, niter, L, d, K, epochs, seed=7 + d)
    seed = 0xABCDEF01 + K
    want0, want1 = syn0.copy(), syn1.copy()
    olib.sgns_train(tok, n, d, want0, want1, sampler=1, K=K, epochs=epochs, job_walks=100,
                    alias_prob=prob, alias_idx=alias, seed=seed)
    # sequential = one worker: the plan (how the corpus is cut into units) cannot matter
    for unit_walks, policy in ((100, 0), (7, 0), (1000, 0)):
        got0, got1 = gpu_sgns(tok, n, d, K, epochs, prob, alias, syn0, syn1, seed, lanes=1,
                              job_walks=100, unit_walks=unit_walks, policy=policy)
        assert np.array_equal(got0.view(np.uint32), want0.view(np.uint32))
        assert np.array_equal(got1.view(np.uint32), want1.view(np.uint32))


@pytest.mark.parametrize('n,m,niter,L,d,K,epochs', [(40, 120, 3, 10, 8, 5, 3),
                                                    (40, 120, 2, 10, 16, 5, 2),
                                                    (30, 90, 5, 7, 8, 3, 2),
                                                    (25, 60, 3, 6, 64, 5, 1),
                                                    (25, 60, 3, 13, 128, 5, 1),
                                                    (12, 14, 6, 10, 8, 5, 3)])
def test_sgns_regenerated_walks_bit_exact(n, m, niter, L, d, K, epochs):
    """tokens == NULL: the workers regenerate every walk from its Philox counter (no corpus in
    memory).  Sequential run, both plan policies, per-pair and merged order: the same bits as the
    oracle trained on the oracle's walks."""
    wseed = 1000 + d
    row_ptr, col_idx = graphs.csr_arrays(n, m, seed=7 + d, self_loops=2)
    tok = olib.walk_uniform(row_ptr, col_idx, niter, L, wseed, 0)
    counts = np.bincount(tok.ravel(), minlength=n)
    prob, alias = olib.alias_table(counts)
    syn0, syn1 = olib.gensim_init_weights(n, d, seed=1)
    for merged in ((False, True) if K == 5 else (False,)):
        want0, want1 = syn0.copy(), syn1.copy()
        olib.sgns_train(tok, n, d, want0, want1, sampler=1, K=K, epochs=epochs, job_walks=50,
                        alias_prob=prob, alias_idx=alias, seed=5, merged=merged)
        for unit_walks, policy in ((50, 0), (7, 0), (1000, 1), (13, 1)):
            got0, got1 = gpu_sgns(tok, n, d, K, epochs, prob, alias, syn0, syn1, 5, lanes=1, job_walks=50,
                                  unit_walks=unit_walks, policy=policy, flags=4 if merged else 0,
                                  graph=(row_ptr, col_idx, niter, wseed))
            assert np.array_equal(got0.view(np.uint32), want0.view(np.uint32)), (merged, unit_walks, policy)
            assert np.array_equal(got1.view(np.uint32), want1.view(np.uint32))


def test_walk_count_equals_token_histogram():
    """gw_walk_count (vocabulary without a stored corpus) == bincount of the oracle's walks, on the
    shared-memory path (n <= 49152) and the global-atomics path."""
    for n, m, niter in ((300, 2000, 7), (60000, 90000, 2)):
        row_ptr, col_idx = graphs.csr_arrays(n, m, seed=3, self_loops=3)
        tok = olib.walk_uniform(row_ptr, col_idx, niter, 10, 99, 0)
        counts = torch.empty(n, dtype=torch.int64, device='cuda')
        _lib.call('gw_walk_count', n, len(col_idx), _lib.ptr(dev(row_ptr, np.int32)),
                  _lib.ptr(dev(col_idx, np.int32)), niter, 10, 99, _lib.ptr(counts), st())
        torch.cuda.synchronize()
        assert np.array_equal(counts.cpu().numpy(), np.bincount(tok.ravel(), minlength=n))


def test_sgns_plan_units():
    """gw_sgns_plan: the units tile the corpus exactly, in order; chunk units have the asked size;
    burst units never cross a start node, carry it, and are near-equal pieces of long bursts."""
    n, niter = 500, 10
    row_ptr, col_idx = graphs.csr_arrays(n, 4000, seed=5, self_loops=2)
    W = niter * len(col_idx)
    start_of_walk = np.repeat(np.repeat(np.arange(n), np.diff(row_ptr)), niter)
    for policy, unit_walks in ((0, 64), (0, 1000), (1, 100), (1, 25), (1, 100000)):
        _, _, u = gpu_plan(n, W, unit_walks, policy, row_ptr, niter)
        assert u['first'][0] == 0 and (u['count'] > 0).all() and (u['count'] <= unit_walks).all()
        assert np.array_equal(u['first'][1:], (u['first'] + u['count'])[:-1])
        assert u['first'][-1] + u['count'][-1] == W
        assert np.array_equal(u['node'], start_of_walk[u['first']])
        if policy == 0:
            assert (u['count'][:-1] == unit_walks).all()
        else:
            assert np.array_equal(start_of_walk[u['first']], start_of_walk[u['first'] + u['count'] - 1])
            burst = np.diff(row_ptr) * niter
            pieces = np.bincount(u['node'], minlength=n)
            assert np.array_equal(pieces, -(-burst // unit_walks))
            big = np.argmax(burst)
            c = u['count'][u['node'] == big]
            assert c.max() - c.min() <= 1
    # hub cap: no start node is cut into more than max_pieces units
    _, _, u = gpu_plan(n, W, 25, 1, row_ptr, niter, max_pieces=7)
    assert u['first'][-1] + u['count'][-1] == W and np.array_equal(u['first'][1:], (u['first'] + u['count'])[:-1])
    burst = np.diff(row_ptr) * niter
    assert np.array_equal(np.bincount(u['node'], minlength=n), np.minimum(-(-burst // 25), 7))
    assert u['count'].max() == -(-burst.max() // 7)
    # a plan without a graph (corpora given as lists of walks)
    _, _, u = gpu_plan(10, 1234, 100, 0, None, 1)
    assert len(u) == 13 and (u['node'] == -1).all() and u['count'][-1] == 34


@pytest.mark.parametrize('n,m,niter,L,d,K,epochs', [(40, 120, 3, 10, 8, 5, 5),
                                                    (40, 120, 2, 10, 16, 5, 2),
                                                    (25, 60, 1, 6, 64, 5, 1),
                                                    (25, 60, 1, 6, 128, 5, 1),
                                                    (30, 90, 2, 7, 32, 5, 2),
                                                    (12, 14, 6, 10, 8, 5, 3)])
def test_sgns_merged_reductions_bit_exact(n, m, niter, L, d, K, epochs):
    """GW_SGNS_MERGED run sequentially: a worker sums its two updates of a row
    in registers and reduces once, computing on its own up-to-date view meanwhile.  Bit-exact against
    the oracle's restatement of that order (gw_oracle.c:walk_merged), which differs from gensim's
    per-pair order by float summation order only (amplified by the 1000-bin sigmoid table: the two
    orders drift apart by ~1e-3 per epoch).  The 12-node case makes walks step back and revisit
    nodes all the time; every 7th walk is cut short."""
    tok, prob, alias, syn0, syn1 = sgns_case(n, m, niter, L, d, K, epochs, seed=11 + d)
    tok = tok.copy()
    tok[L // 2:, ::7] = -1
    want0, want1 = syn0.copy(), syn1.copy()
    olib.sgns_train(tok, n, d, want0, want1, sampler=1, K=K, epochs=epochs, job_walks=100,
                    alias_prob=prob, alias_idx=alias, seed=77, merged=True)
    pp0, pp1 = syn0.copy(), syn1.copy()
    olib.sgns_train(tok, n, d, pp0, pp1, sampler=1, K=K, epochs=epochs, job_walks=100,
                    alias_prob=prob, alias_idx=alias, seed=77)
    assert not np.array_equal(pp0, want0) and np.abs(pp0 - want0).max() < 0.05
    got0, got1 = gpu_sgns(tok, n, d, K, epochs, prob, alias, syn0, syn1, 77, lanes=1,
                          job_walks=100, flags=4)
    assert np.array_equal(got0.view(np.uint32), want0.view(np.uint32))
    assert np.array_equal(got1.view(np.uint32), want1.view(np.uint32))
    # Hogwild mode in both orders (finite; not bit-comparable)
    for flags in (0, 4):
        hw0, _ = gpu_sgns(tok, n, d, K, epochs, prob, alias, syn0, syn1, 77, lanes=8, job_walks=100,
                          unit_walks=4, flags=flags)
        assert np.isfinite(hw0).all()


def test_sgns_large_vocabulary_bit_exact():
    """A vocabulary larger than any shared-memory table (n = 70000): same bits."""
    n, d, K = 70000, 8, 5
    row_ptr, col_idx = graphs.csr_arrays(n, 120000, seed=9, power_law=False)
    tok = olib.walk_uniform(row_ptr, col_idx, 1, 10, 4, 0, 0, 3000)
    counts = np.bincount(tok.ravel(), minlength=n)
    prob, alias = olib.alias_table(counts)
    syn0, syn1 = olib.gensim_init_weights(n, d, seed=1)
    want0, want1 = syn0.copy(), syn1.copy()
    olib.sgns_train(tok, n, d, want0, want1, sampler=1, K=K, epochs=2, alias_prob=prob,
                    alias_idx=alias, seed=8)
    got0, got1 = gpu_sgns(tok, n, d, K, 2, prob, alias, syn0, syn1, 8, lanes=1)
    assert np.array_equal(got0, want0) and np.array_equal(got1, want1)
    # and the concurrent run on the same input stays finite and close in distribution
    got0c, _ = gpu_sgns(tok, n, d, K, 2, prob, alias, syn0, syn1, 8, lanes=64, unit_walks=8)
    assert np.isfinite(got0c).all()
    touched = counts > 0
    assert np.abs(got0c[touched] - want0[touched]).mean() < 0.02


def test_sgns_ragged_walks_bit_exact():
    n, d, K = 30, 8, 5
    tok, prob, alias, syn0, syn1 = sgns_case(n, 90, 2, 10, d, K, 1, seed=3)
    tok = tok.copy()
    rng = np.random.default_rng(0)
    for p in range(tok.shape[1]):
        cut = rng.integers(1, 11)
        tok[cut:, p] = -1
    want0, want1 = syn0.copy(), syn1.copy()
    olib.sgns_train(tok, n, d, want0, want1, sampler=1, K=K, epochs=1, alias_prob=prob,
                    alias_idx=alias, seed=5)
    got0, got1 = gpu_sgns(tok, n, d, K, 1, prob, alias, syn0, syn1, 5, lanes=1)
    assert np.array_equal(got0, want0) and np.array_equal(got1, want1)


def test_sgns_unsupported_raises():
    tok, prob, alias, syn0, syn1 = sgns_case(20, 50, 1, 5, 8, 5, 1, seed=1)
    with pytest.raises(NotImplementedError):      # merged order exists for negative=5 only
        gpu_sgns(tok, 20, 8, 3, 1, prob, alias, syn0, syn1, 1, lanes=1, flags=4)
    with pytest.raises(NotImplementedError):
        gpu_sgns(tok, 20, 12, 5, 1, prob, alias, np.zeros((20, 12), np.float32),
                 np.zeros((20, 12), np.float32), 1, lanes=1)


def test_sgns_hogwild_close_to_sequential():
    """Concurrent workers (16 workers, 25-walk units on a 200-node graph) vs the sequential oracle on
    the same corpus and negatives: the runs differ only by stale reads and the order of float adds;
    the neighbour-similarity distributions agree."""
    from scipy.stats import ks_2samp
    n, d, K, epochs = 200, 8, 5, 5
    tok, prob, alias, syn0, syn1 = sgns_case(n, 1500, 20, 10, d, K, epochs, seed=21)
    want0, want1 = syn0.copy(), syn1.copy()
    olib.sgns_train(tok, n, d, want0, want1, sampler=1, K=K, epochs=epochs, alias_prob=prob,
                    alias_idx=alias, seed=77)
    got0, _ = gpu_sgns(tok, n, d, K, epochs, prob, alias, syn0, syn1, 77, lanes=16, unit_walks=25)
    assert np.isfinite(got0).all()
    row_ptr, col_idx = graphs.csr_arrays(n, 1500, seed=21, self_loops=2)
    s_want = stats_ref.null_similarities(row_ptr, col_idx, want0)
    s_got = stats_ref.null_similarities(row_ptr, col_idx, got0)
    assert ks_2samp(s_want, s_got).statistic < 0.08
    assert abs(s_want.mean() - s_got.mean()) < 0.03


def test_sgns_auto_shape():
    """The default shape is a function of the network alone (never of GPU fill or GPU count)."""
    lanes, unit, policy = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int32(0)

    def shape(n, W, hub):
        _lib.call('gw_sgns_auto_shape', n, W, 1000, hub, ctypes.byref(lanes), ctypes.byref(unit),
                  ctypes.byref(policy))
        return lanes.value, unit.value, policy.value
    assert shape(2300, 3465000, 36100) == (95, 120, 0)            # C1: no hubs, short units
    assert shape(8500, 45109200, 192800) == (354, 320, 0)         # hubs: units of >= 320 walks
    assert shape(38000, 152681600, 729700) == (1536, 980, 0)      # C2
    assert shape(38000, 152681600, -1) == (1536, 980, 0)          # hub size unknown
    assert shape(1000000, 1000000000, 497900) == (7812, 1000, 0)  # C5: tables of few trainings fit L2
    assert shape(50000, 977650800, 1836900) == (1536, 1000, 0)    # C4
    assert shape(3, 400, 200)[0] == 1
    for n, W in ((2300, 3465000), (38000, 152681600), (1000000, 1000000000), (3, 400)):
        l, u, p = shape(n, W, -1)
        assert 1 <= l <= max(1, n // 2) and u >= 1 and p in (0, 1) and l * u <= max(W, u)
    lib = _lib.load()
    assert lib.gw_sgns_resident_trainings(8, 1536) == 148 * 128 // 1536
    assert lib.gw_sgns_resident_trainings(8, 1) >= 1024 and lib.gw_sgns_resident_trainings(128, 10 ** 6) == 1


# ------------------------------------------------------------------------------------------------
# statistics kernels
# ------------------------------------------------------------------------------------------------
def test_cosine_pairs_and_null_similarities():
    rng = np.random.default_rng(0)
    n, d = 500, 8
    vec = rng.standard_normal((n, d)).astype(np.float32)
    a = rng.integers(0, n, size=3000).astype(np.int32)
    b = rng.integers(0, n, size=3000).astype(np.int32)
    out = torch.empty(3000, dtype=torch.float32, device='cuda')
    _lib.call('gw_cosine_pairs', d, _lib.ptr(dev(vec)), 3000, _lib.ptr(dev(a)), _lib.ptr(dev(b)),
              _lib.ptr(out), st())
    want = stats_ref.cosine_similarity(vec, a, b)
    assert np.abs(out.cpu().numpy() - want).max() <= 1e-5      # north_star: within 1e-5 fp32
    row_ptr, col_idx = graphs.csr_arrays(n, 4000, seed=2, self_loops=9)
    sims = torch.empty(int(row_ptr[-1]), dtype=torch.float32, device='cuda')
    cnt = ctypes.c_int64(0)
    _lib.call('gw_null_similarities', n, _lib.ptr(dev(row_ptr)), _lib.ptr(dev(col_idx)), d,
              _lib.ptr(dev(vec)), _lib.ptr(sims), ctypes.byref(cnt), st())
    want = stats_ref.null_similarities(row_ptr, col_idx, vec)
    assert cnt.value == len(want)
    assert np.abs(sims[:cnt.value].cpu().numpy() - want).max() <= 1e-5


def test_psim_matches_reference_formula():
    rng = np.random.default_rng(1)
    srd = np.sort(rng.random(100000).astype(np.float32))
    sims = np.concatenate([rng.random(5000).astype(np.float32), srd[::997], [0.0, 1.0, -1.0, 2.0]]
                          ).astype(np.float32)
    ranks = torch.empty(len(sims), dtype=torch.int64, device='cuda')
    p = torch.empty(len(sims), dtype=torch.float64, device='cuda')
    _lib.call('gw_psim', _lib.ptr(dev(srd)), len(srd), _lib.ptr(dev(sims)), len(sims),
              _lib.ptr(ranks), _lib.ptr(p), st())
    assert np.array_equal(ranks.cpu().numpy(), stats_ref.psim_rank(srd, sims))
    assert np.array_equal(p.cpu().numpy(), stats_ref.psim(srd, sims))     # bit-exact float64


@pytest.mark.parametrize('m,nseg', [(1, 1), (10, 1), (5000, 1), (5000, 700), (200000, 1),
                                    (200000, 20000)])
def test_bh_segmented(m, nseg):
    rng = np.random.default_rng(m + nseg)
    p = np.maximum(1.0 - rng.integers(0, 1000, size=m) / 1000.0, 1e-16)    # many ties, like psim
    cuts = np.sort(rng.choice(np.arange(1, m), size=nseg - 1, replace=False)) if nseg > 1 else []
    seg = np.concatenate([[0], cuts, [m]]).astype(np.int64)
    q = torch.empty(m, dtype=torch.float64, device='cuda')
    _lib.call('gw_bh_segmented', _lib.ptr(dev(p)), m, _lib.ptr(dev(seg)), nseg, _lib.ptr(q), st())
    want = stats_ref.fdrcorrection_segmented(p, seg)
    assert np.allclose(q.cpu().numpy(), want, rtol=1e-14, atol=0)


def test_log_stats_and_mean_sem():
    rng = np.random.default_rng(2)
    for nreps in (1, 2, 3, 10):
        vals = np.maximum(rng.random((300, nreps)) ** 4, 1e-16)
        vals[0] = 1e-16
        vals[1] = 0.5
        out = torch.empty((300, 3), dtype=torch.float64, device='cuda')
        _lib.call('gw_log_stats', _lib.ptr(dev(vals)), 300, nreps, _lib.ptr(out), st())
        want = np.array([stats_ref.log_stats(v) for v in vals])
        assert np.allclose(out.cpu().numpy(), want, rtol=1e-10, atol=1e-300)
        sims = rng.random((300, nreps)).astype(np.float32)
        mean = torch.empty(300, dtype=torch.float32, device='cuda')
        sem = torch.empty(300, dtype=torch.float64, device='cuda')
        _lib.call('gw_mean_sem_f32', _lib.ptr(dev(sims)), 300, nreps, _lib.ptr(mean), _lib.ptr(sem), st())
        assert np.allclose(mean.cpu().numpy(), sims.mean(axis=1), rtol=1e-6)
        assert np.allclose(sem.cpu().numpy(), sims.std(axis=1) / np.sqrt(nreps), rtol=1e-4, atol=1e-7)
