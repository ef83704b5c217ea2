"""CPU tests: pin the oracle against the reference's own behaviour (golden fixtures produced by
tests/golden/make_golden.py from the unmodified reference), against known-answer vectors and
against independent library implementations (scipy, networkx)."""
import json
import os

import numpy as np
import networkx as nx
import pytest

from oracle import lib as olib
from oracle import config_model_ref, stats_ref
from oracle.philox import philox4x32_10
from genewalk_b200.csr import GraphCSR, multigraph_degrees, coprime_stride

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_graph(spec):
    mg = nx.MultiGraph()
    for n, attrs in spec['nodes']:
        mg.add_node(n, **attrs)
    mg.add_edges_from(spec['edges'])
    return mg


# ---- Philox ---------------------------------------------------------------------------------
KAT = [  # Random123 kat_vectors, philox4x32 10 rounds
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


@pytest.mark.parametrize('ctr,key,want', KAT)
def test_philox_known_answers(ctr, key, want):
    assert tuple(int(x) for x in olib.philox(ctr, key)) == want
    assert tuple(int(x) for x in philox4x32_10(*ctr, *key)) == want


def test_philox_numpy_matches_c():
    rng = np.random.default_rng(0)
    c = rng.integers(0, 2 ** 32, size=(50, 4), dtype=np.uint64)
    k = rng.integers(0, 2 ** 32, size=2, dtype=np.uint64)
    out = np.stack(philox4x32_10(c[:, 0], c[:, 1], c[:, 2], c[:, 3], int(k[0]), int(k[1])), axis=1)
    for i in range(50):
        assert np.array_equal(out[i], olib.philox(c[i], k))


# ---- walks vs the reference -------------------------------------------------------------------
with open(os.path.join(GOLD, 'walk_structure.json')) as f:
    WALK_CASES = json.load(f)


@pytest.mark.parametrize('case', WALK_CASES, ids=[c['name'] for c in WALK_CASES])
@pytest.mark.parametrize('stride', [0, 1, None])
def test_oracle_walks_match_reference_structure(case, stride):
    """Same number of walks, same walks per start node, same length as the reference's
    DeepWalk.get_walks (golden), for every slot ordering; corpus order (stride 0) is node-major."""
    mg = load_graph(case['graph'])
    csr = GraphCSR.from_networkx(mg)
    s = coprime_stride(csr.nnz) if stride is None else stride
    tok = olib.walk_uniform(csr.row_ptr, csr.col_idx, case['niter'], case['walk_length'], 42, s)
    assert tok.shape == (case['walk_length'], case['n_walks'])
    assert case['lengths'] == [case['walk_length']]
    starts = np.bincount(tok[0], minlength=csr.n)
    want = np.array([case['starts'].get(v, 0) for v in csr.nodes])
    assert np.array_equal(starts, want)
    if stride == 0:   # reference serial order: walks grouped by start node in graph order
        assert np.array_equal(tok[0], np.repeat(np.arange(csr.n), want))
    # every step goes to a distinct neighbour of the current node
    nbr = {(u, v) for u in range(csr.n)
           for v in csr.col_idx[csr.row_ptr[u]:csr.row_ptr[u + 1]]}
    assert all((a, b) in nbr for a, b in zip(tok[:-1].ravel().tolist(), tok[1:].ravel().tolist()))


@pytest.mark.parametrize('case', WALK_CASES, ids=[c['name'] for c in WALK_CASES])
def test_oracle_transitions_match_reference_distribution(case):
    """Chi-square homogeneity: 20k reference (MT19937 random.choice) steps out of the hub vs the
    Philox sampler's steps out of the same node; and chi-square of the latter against 1/deg."""
    from scipy.stats import chi2_contingency, chisquare
    mg = load_graph(case['graph'])
    csr = GraphCSR.from_networkx(mg)
    hub = csr.index[case['hub']]
    tok = olib.walk_uniform(csr.row_ptr, csr.col_idx, 2000, 2, 7, 0)
    nxt = tok[1][tok[0] == hub]
    nbrs = csr.col_idx[csr.row_ptr[hub]:csr.row_ptr[hub + 1]]
    ours = np.array([(nxt == v).sum() for v in nbrs])
    ref = np.array([case['hub_transitions'].get(csr.nodes[v], 0) for v in nbrs])
    assert ours.sum() == len(nxt) and ref.sum() == 20000
    assert chisquare(ours).pvalue > 1e-4
    assert chi2_contingency(np.stack([ours, ref]))[1] > 1e-4


def test_csr_matches_networkx_adjacency_order():
    mg = load_graph(WALK_CASES[1]['graph'])
    csr = GraphCSR.from_networkx(mg)
    for v in mg.nodes():
        i = csr.index[v]
        got = [csr.nodes[j] for j in csr.col_idx[csr.row_ptr[i]:csr.row_ptr[i + 1]]]
        assert got == list(mg[v])
    # a: 3 parallel edges to b + a self-loop + d + e -> distinct neighbours b, a, d, e
    assert len(mg['a']) == 4 and mg.degree('a') == 7
    assert np.array_equal(multigraph_degrees(mg), [mg.degree(v) for v in mg.nodes()])


# ---- configuration model vs the reference -------------------------------------------------------
with open(os.path.join(GOLD, 'rand_graph.json')) as f:
    RAND_CASES = json.load(f)


@pytest.mark.parametrize('case', RAND_CASES, ids=[c['name'] for c in RAND_CASES])
def test_config_model_matches_reference_semantics(case):
    mg = load_graph(case['graph'])
    d_seq = config_model_ref.multigraph_degrees_desc(multigraph_degrees(mg))
    # reference: node i of the randomized graph ('n<i>') has the i-th largest multigraph degree
    assert case['nodes'] == ['n%d' % i for i in range(len(d_seq))]
    assert case['degrees'] == d_seq.tolist()
    loops, par = [], []
    for s in range(20):
        u, v = config_model_ref.config_model_edges(d_seq, 1000 + s)
        assert len(u) == case['n_edges']
        deg = np.bincount(np.concatenate([u, v]), minlength=len(d_seq))
        assert np.array_equal(deg, d_seq)          # degrees preserved, loops count twice
        loops.append(int((u == v).sum()))
        key = np.minimum(u, v).astype(np.int64) * len(d_seq) + np.maximum(u, v)
        par.append(len(key) - len(np.unique(key)))
    # self-loop / parallel-edge rates agree with networkx's MT19937 shuffle (20 draws each)
    assert abs(np.mean(loops) - case['mean_selfloops']) <= 3 * np.sqrt(case['mean_selfloops'] / 10 + 0.1)
    assert abs(np.mean(par) - case['mean_parallel']) <= 3 * np.sqrt(case['mean_parallel'] / 10 + 0.1)


def test_config_model_uniform_shuffle():
    """Every stub position is equally likely to hold every node's stubs (uniform permutation)."""
    d_seq = np.array([3, 2, 2, 1], dtype=np.int32)
    counts = np.zeros((4, 4))
    for s in range(4000):
        u, v = config_model_ref.config_model_edges(d_seq, s)
        counts[0, u[0]] += 1; counts[1, v[0]] += 1; counts[2, u[3]] += 1; counts[3, v[3]] += 1
    exp = 4000 * d_seq / 8.0
    assert (np.abs(counts - exp) < 5 * np.sqrt(exp)).all()
    with pytest.raises(ValueError):
        config_model_ref.config_model_edges(np.array([2, 1]), 0)


def test_csr_distinct():
    rp, ci = config_model_ref.csr_distinct(5, np.array([0, 0, 2, 3]), np.array([1, 1, 2, 0]))
    assert rp.tolist() == [0, 2, 3, 4, 5, 5] and ci.tolist() == [1, 3, 0, 2, 0]


# ---- null similarities vs the reference ---------------------------------------------------------
def test_null_similarities_match_reference():
    g = np.load(os.path.join(GOLD, 'null_dist.npz'))
    nodes = [str(x) for x in g['nodes']]
    index = {v: i for i, v in enumerate(nodes)}
    csr = GraphCSR(nodes, g['adj_ptr'], [index[str(w)] for w in g['adj']], index)
    vec = np.zeros((csr.n, 8), dtype=np.float32)
    for k, v in zip(g['keys'], g['vectors']):
        vec[csr.index[str(k)]] = v
    got = stats_ref.null_similarities(csr.row_ptr, csr.col_idx, vec)
    assert len(got) == len(g['srd'])
    # same order (node-major, adjacency order), float32: within 1e-6 of the reference's BLAS path
    assert np.abs(got - g['srd']).max() <= 1e-6


# ---- statistics vs the reference / scipy -----------------------------------------------------------
with open(os.path.join(GOLD, 'statistics.json')) as f:
    STAT = json.load(f)


def test_psim_and_log_stats_match_reference():
    null = np.load(os.path.join(GOLD, 'statistics.npz'))['null']
    sims = np.array(STAT['psim']['sims'], dtype=np.float32)
    assert np.array_equal(stats_ref.psim(null, sims), np.array(STAT['psim']['pvals']))
    for vals, want in zip(STAT['log_stats']['vals'], STAT['log_stats']['out']):
        got = stats_ref.log_stats(vals)
        assert np.allclose(got, want, rtol=1e-12, atol=1e-300)


def test_bh_matches_scipy():
    from scipy.stats import false_discovery_control
    rng = np.random.default_rng(0)
    for n in (1, 2, 5, 100, 1000):
        p = np.maximum(1 - rng.integers(0, 50, size=n) / 50.0, 1e-16)
        assert np.allclose(stats_ref.fdrcorrection(p), false_discovery_control(p, method='bh'),
                           rtol=1e-12)
    assert len(stats_ref.fdrcorrection([])) == 0


def test_gene_padj_matches_reference_output():
    """oracle cosine -> psim -> per-gene BH -> log_stats reproduces the reference DataFrame."""
    z = np.load(os.path.join(GOLD, 'statistics.npz'))
    mg = load_graph(STAT['graph'])
    index = {v: i for i, v in enumerate(mg.nodes())}
    out = STAT['outputs']['1']
    cols = out['columns']
    recs = [dict(zip(cols, r)) for r in out['records'] if r[cols.index('go_id')]]
    vecs = []
    for r in range(3):
        m = np.zeros((len(index), 8), dtype=np.float32)
        for k, v in zip(STAT['keys'][r], z['vec%d' % r]):
            m[index[k]] = v
        vecs.append(m)
    genes = sorted({r['custom'] for r in recs}, key=lambda g: index[g])
    pg, po, seg = [], [], [0]
    by = {}
    for r in recs:
        by.setdefault(r['custom'], []).append(r)
    for g in genes:
        for r in by[g]:
            pg.append(index[g]); po.append(index[r['go_id']])
        seg.append(len(po))
    sims = np.stack([stats_ref.cosine_similarity(v, pg, po) for v in vecs])
    padj = stats_ref.gene_padj(sims, z['null'], np.array(seg))
    want = np.array([r['gene_padj'] for g in genes for r in by[g]])
    assert np.allclose(padj, want, rtol=1e-9, atol=1e-15)
    want_sim = np.array([r['sim'] for g in genes for r in by[g]])
    assert np.abs(sims.mean(axis=0) - want_sim).max() <= 1e-6


# ---- SGNS oracle ---------------------------------------------------------------------------------
def test_exp_table_and_cum_table():
    t = olib.exp_table()
    assert t.dtype == np.float32 and len(t) == 1000
    x = (np.arange(1000) / 1000.0 * 2 - 1) * 6
    assert np.abs(t - 1 / (1 + np.exp(-x))).max() < 1e-6
    counts = np.array([1000, 100, 100, 10, 1])
    cum = olib.gensim_cum_table(counts)
    assert cum.dtype == np.uint32 and cum[-1] == 2 ** 31 - 1 and (np.diff(cum.astype(np.int64)) > 0).all()
    p = np.diff(np.concatenate([[0], cum.astype(np.float64)])) / (2 ** 31 - 1)
    assert np.allclose(p, counts ** 0.75 / (counts ** 0.75).sum(), atol=1e-8)


def test_gensim_vocab_order():
    tok = np.array([[3, 1, 1], [1, 3, 2], [0, 0, 3]], dtype=np.int32)   # [L=3, W=3]
    itn, nti, cnt = olib.gensim_vocab(tok, 5)
    # counts: 0:2 1:3 2:1 3:3 -> descending, ties (1 vs 3) by first appearance in the corpus scan
    # corpus scan (walk-major): walk0 = 3,1,0 -> 3 first
    assert itn.tolist() == [3, 1, 0, 2] and cnt.tolist() == [3, 3, 2, 1]
    assert nti.tolist() == [2, 1, 3, 0, -1]


def test_alias_table_oracle():
    rng = np.random.default_rng(3)
    counts = rng.zipf(1.6, size=500) * 5
    counts[::37] = 0
    prob, alias = olib.alias_table(counts)
    w = np.where(counts > 0, counts.astype(np.float64) ** 0.75, 0)
    refs = 1 + np.bincount(alias, minlength=500)
    assert (np.abs(olib.alias_implied_distribution(prob, alias) - w / w.sum())
            <= refs * 2.0 ** -24 / 500 + 1e-15).all()


def _ring_of_cliques():
    g = nx.MultiGraph(nx.ring_of_cliques(6, 5))
    return GraphCSR.from_networkx(g)


def test_sgns_oracle_learns_structure_both_samplers():
    """Both negative samplers (gensim cum_table+LCG; Philox+alias), on the same corpus in the
    reference's order, train embeddings in which neighbours are more similar than non-neighbours,
    and their neighbour-similarity distributions agree (KS)."""
    from scipy.stats import ks_2samp
    csr = _ring_of_cliques()
    n = csr.n
    sims = {}
    for mode in ('gensim', 'device'):
        acc = []
        for seed in range(6):
            if mode == 'gensim':
                tok = olib.walk_uniform(csr.row_ptr, csr.col_idx, 100, 10, seed, 0)
                itn, syn0, _ = olib.word2vec_gensim_like(tok, n)
                vec = np.zeros((n, 8), np.float32)
                vec[itn] = syn0
            else:
                tok = olib.walk_uniform(csr.row_ptr, csr.col_idx, 100, 10, seed, 0)
                counts = np.bincount(tok.ravel(), minlength=n)
                prob, alias = olib.alias_table(counts)
                vec, syn1 = olib.gensim_init_weights(n, 8)
                olib.sgns_train(tok, n, 8, vec, syn1, sampler=1, alias_prob=prob, alias_idx=alias,
                                seed=seed)
            acc.append(stats_ref.null_similarities(csr.row_ptr, csr.col_idx, vec))
            a, b = np.triu_indices(n, 1)
            allp = stats_ref.cosine_similarity(vec, a, b)
            assert acc[-1].mean() > allp.mean() + 0.2
        sims[mode] = np.concatenate(acc)
    assert ks_2samp(sims['gensim'], sims['device']).statistic < 0.1


def test_sgns_oracle_threads_smoke():
    csr = _ring_of_cliques()
    tok = olib.walk_uniform(csr.row_ptr, csr.col_idx, 40, 10, 1, 0)
    itn, syn0, _ = olib.word2vec_gensim_like(tok, csr.n, nthreads=2)
    assert np.isfinite(syn0).all() and len(itn) == csr.n


@pytest.mark.parametrize('merged', [False, True])
@pytest.mark.parametrize('d', [8, 16])
def test_sgns_c_oracle_matches_python_statement(merged, d):
    """oracle/gw_oracle.c (per-pair = gensim's order, and the kernel's merged order) against the
    pure-Python statement of the same definitions, bit for bit; ragged walks, walks that step back,
    repeated negatives and negatives equal to the centre all occur on this 9-node graph."""
    from oracle import sgns_py
    from genewalk_b200 import workloads as graphs
    n = 9
    row_ptr, col_idx = graphs.csr_arrays(n, 12, seed=4, self_loops=1)
    tok = olib.walk_uniform(row_ptr, col_idx, 2, 10, 3).copy()
    tok[6:, ::3] = -1
    counts = np.bincount(tok[tok >= 0].ravel(), minlength=n)
    prob, alias = olib.alias_table(counts)
    syn0, syn1 = olib.gensim_init_weights(n, d, seed=1)
    want0, want1 = syn0.copy(), syn1.copy()
    sgns_py.train(tok, n, want0, want1, prob, alias, 99, olib.exp_table(), epochs=2, job_walks=10,
                  merged=merged)
    got0, got1 = syn0.copy(), syn1.copy()
    olib.sgns_train(tok, n, d, got0, got1, sampler=1, K=5, epochs=2, job_walks=10, alias_prob=prob,
                    alias_idx=alias, seed=99, merged=merged)
    assert not np.array_equal(got0, syn0)
    assert np.array_equal(got0.view(np.uint32), want0.view(np.uint32))
    assert np.array_equal(got1.view(np.uint32), want1.view(np.uint32))


def test_sgns_merged_order_stays_close_to_per_pair():
    """The two orders differ by float summation order only (amplified by the 1000-bin sigmoid
    table): same structure learned, small drift."""
    csr = _ring_of_cliques()
    n = csr.n
    tok = olib.walk_uniform(csr.row_ptr, csr.col_idx, 100, 10, 5, 0)
    counts = np.bincount(tok.ravel(), minlength=n)
    prob, alias = olib.alias_table(counts)
    out = {}
    for merged in (False, True):
        vec, syn1 = olib.gensim_init_weights(n, 8)
        olib.sgns_train(tok, n, 8, vec, syn1, sampler=1, alias_prob=prob, alias_idx=alias, seed=3,
                        merged=merged)
        out[merged] = stats_ref.null_similarities(csr.row_ptr, csr.col_idx, vec)
    assert not np.array_equal(out[False], out[True])
    assert np.abs(out[False] - out[True]).max() < 0.05
