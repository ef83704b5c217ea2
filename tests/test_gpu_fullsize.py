"""Full-size (BASELINE.json configs[1], human-scale network) checks through size-independent
properties: the oracle cannot be run at this size in seconds, so the CUDA path is checked against
invariants of the reference's algorithm."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')

from genewalk_b200 import _lib                      # noqa: E402
from genewalk_b200.csr import GraphCSR              # noqa: E402
from genewalk_b200.deepwalk import DeepWalk         # noqa: E402
from genewalk_b200 import workloads as graphs                            # noqa: E402


@pytest.fixture(scope='module')
def c2():
    n, u, v, is_go = graphs.human_scale_edges(seed=2)
    rp, ci = graphs.csr_from_edges_numpy(n, u, v)
    csr = GraphCSR(['v%d' % i for i in range(n)], rp, ci).to_device()
    dw = DeepWalk(csr, 10, 100, seed=77)
    dw.get_walks()
    return csr, dw


def test_walks_full_size_invariants(c2):
    csr, dw = c2
    tok = dw.walks.tokens                          # [10, W] int32 on the device
    L, W = tok.shape
    A, n = csr.nnz, csr.n
    assert W == 100 * A and L == 10                # niter * sum(len(graph[v])) (deepwalk.py:173,197)
    deg = torch.from_numpy(np.diff(csr.row_ptr)).cuda()
    # start nodes: node v starts exactly niter*deg(v) walks, in node-major corpus order
    starts = torch.bincount(tok[0].long(), minlength=n)
    assert torch.equal(starts, 100 * deg.long())
    assert bool((tok[0, 1:] >= tok[0, :-1]).all())
    # every step follows an edge: (u, v) keys of a 4M-step sample are all present in the CSR
    src = torch.repeat_interleave(torch.arange(n, device='cuda'), deg.long())
    keys = src * n + csr.d_col_idx.long()          # CSR from csr_from_edges_numpy is sorted by key
    g = torch.Generator(device='cuda').manual_seed(1)
    p = torch.randint(0, W, (4_000_000,), device='cuda', generator=g)
    i = torch.randint(0, L - 1, (4_000_000,), device='cuda', generator=g)
    q = tok[i, p].long() * n + tok[i + 1, p].long()
    pos = torch.searchsorted(keys, q).clamp(max=keys.numel() - 1)
    assert bool((keys[pos] == q).all())
    # token counts: the count kernel equals torch's histogram; stationarity: count ~ niter*L*deg
    counts = torch.empty(n, dtype=torch.int64, device='cuda')
    _lib.call('gw_count_tokens', _lib.ptr(tok), L * W, n, _lib.ptr(counts), _lib.stream_ptr())
    assert torch.equal(counts, torch.bincount(tok.reshape(-1).long(), minlength=n))
    rel = (counts.double() - 1000.0 * deg.double()).abs() / (1000.0 * deg.double()).clamp(min=1)
    assert float(rel[deg > 50].max()) < 0.05


def test_sgns_full_size_epoch_is_sane(c2):
    csr, dw = c2
    dv = dw.train_device(epochs=1)
    torch.cuda.synchronize()
    s0, s1 = dv['syn0'], dv['syn1neg']
    assert bool(torch.isfinite(s0).all()) and bool(torch.isfinite(s1).all())
    present = dv['counts'] > 0
    assert float(s0[present].norm(dim=1).max()) < 50 and float(s1.norm(dim=1).max()) < 50
    assert float(s1[present].norm(dim=1).min()) > 0          # every vocabulary row was trained
    assert bool((s0[~present] == 0).all())                    # nodes outside the vocabulary untouched
    # neighbours are more similar than random pairs after one epoch
    full = torch.nn.functional.normalize(s0, dim=1)
    deg = torch.from_numpy(np.diff(csr.row_ptr)).cuda()
    src = torch.repeat_interleave(torch.arange(csr.n, device='cuda'), deg.long())[::50]
    dst = csr.d_col_idx.long()[::50]
    nbr = (full[src] * full[dst]).sum(1).mean()
    g = torch.Generator(device='cuda').manual_seed(2)
    a = torch.randint(0, csr.n, (100000,), device='cuda', generator=g)
    b = torch.randint(0, csr.n, (100000,), device='cuda', generator=g)
    rnd = (full[a] * full[b]).sum(1).mean()
    assert float(nbr) > float(rnd) + 0.1


def test_statistics_full_size_invariants():
    g = torch.Generator(device='cuda').manual_seed(3)
    null = torch.rand(20_000_000, device='cuda', generator=g) * 2 - 1      # nreps_null * A values
    chk = null.double().sum()
    _lib.call('gw_sort_f32', _lib.ptr(null), null.numel(), _lib.stream_ptr())
    assert bool((null[1:] >= null[:-1]).all())
    assert abs(float(null.double().sum() - chk)) < 1e-6 * null.numel()      # same multiset (checksum)
    sims = torch.rand(1_000_000, device='cuda', generator=g) * 2 - 1
    p = torch.empty(sims.numel(), dtype=torch.float64, device='cuda')
    rk = torch.empty(sims.numel(), dtype=torch.int64, device='cuda')
    _lib.call('gw_psim', _lib.ptr(null), null.numel(), _lib.ptr(sims), sims.numel(), _lib.ptr(rk),
              _lib.ptr(p), _lib.stream_ptr())
    assert torch.equal(rk, torch.searchsorted(null, sims))                    # side='left'
    order = torch.argsort(sims)
    assert bool((p[order][1:] <= p[order][:-1]).all()) and float(p.min()) >= 1e-16
    # global Benjamini-Hochberg over 1M p-values: q >= p, q <= 1, monotone in p, max q == max p
    seg = torch.tensor([0, p.numel()], dtype=torch.int64, device='cuda')
    q = torch.empty_like(p)
    _lib.call('gw_bh_segmented', _lib.ptr(p), p.numel(), _lib.ptr(seg), 1, _lib.ptr(q), _lib.stream_ptr())
    assert bool((q >= p).all()) and float(q.max()) <= 1.0
    po = torch.argsort(p)
    assert bool((q[po][1:] >= q[po][:-1]).all())
    assert float(q[po][-1]) == float(p[po][-1])
    # idempotence of the sort, and the segmented form equals the global form on one segment per row
    before = null.clone()
    _lib.call('gw_sort_f32', _lib.ptr(null), null.numel(), _lib.stream_ptr())
    assert torch.equal(before, null)


def test_randomizer_full_size_preserves_degrees():
    n, u, v, is_go = graphs.human_scale_edges(seed=2)
    deg = np.bincount(np.concatenate([u, v]), minlength=n).astype(np.int64)   # multigraph degree
    d_seq = np.sort(deg)[::-1].astype(np.int32)
    S = int(d_seq.astype(np.int64).sum())
    d = torch.from_numpy(np.ascontiguousarray(d_seq)).cuda()
    eu = torch.empty(S // 2, dtype=torch.int32, device='cuda')
    ev = torch.empty(S // 2, dtype=torch.int32, device='cuda')
    _lib.call('gw_config_model', n, _lib.ptr(d), S, 99, _lib.ptr(eu), _lib.ptr(ev), _lib.stream_ptr())
    got = torch.bincount(torch.cat([eu, ev]).long(), minlength=n)
    assert torch.equal(got.cpu(), torch.from_numpy(d_seq.astype(np.int64)))   # loops count twice
    row_ptr = torch.zeros(n + 1, dtype=torch.int32, device='cuda')
    col = torch.empty(S, dtype=torch.int32, device='cuda')
    nnz = ctypes.c_int64(0)
    _lib.call('gw_csr_from_edges', n, S // 2, _lib.ptr(eu), _lib.ptr(ev), _lib.ptr(row_ptr),
              _lib.ptr(col), ctypes.byref(nnz), _lib.stream_ptr())
    rp = row_ptr.long()
    assert int(rp[-1]) == nnz.value <= S and bool((rp[1:] >= rp[:-1]).all())
    # symmetric, rows sorted and duplicate-free
    src = torch.repeat_interleave(torch.arange(n, device='cuda'), (rp[1:] - rp[:-1]))
    c = col[:nnz.value].long()
    key = src * n + c
    assert bool((key[1:] > key[:-1]).all())
    rev = torch.sort(c * n + src).values
    assert torch.equal(rev, key)
    # a different seed gives a different pairing with the same degrees
    eu2 = torch.empty_like(eu)
    ev2 = torch.empty_like(ev)
    _lib.call('gw_config_model', n, _lib.ptr(d), S, 100, _lib.ptr(eu2), _lib.ptr(ev2), _lib.stream_ptr())
    assert not torch.equal(eu, eu2)
    assert torch.equal(torch.bincount(torch.cat([eu2, ev2]).long(), minlength=n), got)
