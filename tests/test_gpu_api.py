"""GPU tests of the drop-in Python API (genewalk_b200.deepwalk / null_distributions /
perform_statistics) against the reference's own test (genewalk/tests/test_deepwalk.py), the golden
fixtures produced from the unmodified reference, and the oracle."""
import copy
import json
import os
import pickle
import random

import numpy as np
import networkx as nx
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')

from genewalk_b200.deepwalk import (DeepWalk, run_walks, get_start_nodes, run_walks_for_node,  # noqa: E402
                                    run_single_walk)
from genewalk_b200.null_distributions import get_rand_graph, get_null_distributions  # noqa: E402
from genewalk_b200.perform_statistics import GeneWalk  # noqa: E402
from genewalk_b200.keyedvectors import NodeVectors  # noqa: E402
from genewalk_b200.csr import GraphCSR, multigraph_degrees  # noqa: E402
from oracle import lib as olib  # noqa: E402
from oracle import stats_ref  # noqa: E402
from genewalk_b200 import workloads as graphs  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_graph(spec):
    mg = nx.MultiGraph()
    for n, attrs in spec['nodes']:
        mg.add_node(n, **attrs)
    mg.add_edges_from(spec['edges'])
    return mg


# ---- the reference's own test, verbatim through the drop-in (genewalk/tests/test_deepwalk.py) ----
def test_run_walks():
    mg = nx.MultiGraph()
    mg.add_edge('a', 'b')
    mg.add_edge('b', 'c')
    dw = DeepWalk(mg, 10, 100)
    dw.get_walks(workers=1)
    # Number of neighbors for all nodes together is 4, times niter: 400
    assert len(dw.walks) == 400, len(dw.walks)
    assert len([w for w in dw.walks if w[0] == 'a']) == 100
    assert len([w for w in dw.walks if w[0] == 'b']) == 200

    dw.get_walks(workers=2)
    # Number of neighbors for all nodes together is 4, times niter: 400
    assert len(dw.walks) == 400, len(dw.walks)
    assert len([w for w in dw.walks if w[0] == 'a']) == 100
    assert len([w for w in dw.walks if w[0] == 'b']) == 200


with open(os.path.join(GOLD, 'walk_structure.json')) as f:
    WALK_CASES = json.load(f)


@pytest.mark.parametrize('case', WALK_CASES, ids=[c['name'] for c in WALK_CASES])
def test_walks_match_reference_structure(case):
    mg = load_graph(case['graph'])
    dw = DeepWalk(mg, case['walk_length'], case['niter'], seed=3)
    dw.get_walks()
    assert len(dw.walks) == case['n_walks']
    starts = {}
    for w in dw.walks:
        assert len(w) == case['walk_length']
        starts[w[0]] = starts.get(w[0], 0) + 1
        assert all(b in mg[a] for a, b in zip(w[:-1], w[1:]))
    assert starts == case['starts']
    # corpus order of the lazy view = the reference's node-major order
    assert [w[0] for w in dw.walks] == get_start_nodes(mg, case['niter'])
    assert dw.walks[0][0] == get_start_nodes(mg, case['niter'])[0]
    assert dw.walks[-1] == dw.walks[len(dw.walks) - 1]


def test_walks_reproducible_by_python_seed():
    mg, _ = graphs.small(seed=2)
    random.seed(5)
    a = DeepWalk(mg, 10, 3); a.get_walks()
    random.seed(5)
    b = DeepWalk(mg, 10, 3); b.get_walks(workers=4)
    c = DeepWalk(mg, 10, 3); c.get_walks()
    assert list(a.walks) == list(b.walks) and list(a.walks) != list(c.walks)


def test_walk_helpers():
    mg = load_graph(WALK_CASES[1]['graph'])
    ws = run_walks_for_node('a', mg, 5, 6, seed=1)
    assert len(ws) == 5 * len(mg['a']) and all(w[0] == 'a' and len(w) == 6 for w in ws)
    assert run_walks_for_node('iso', mg, 5, 6) == []
    w = run_single_walk('b', mg, 4, seed=2)
    assert w[0] == 'b' and len(w) == 4 and all(y in mg[x] for x, y in zip(w[:-1], w[1:]))
    with pytest.raises(IndexError):
        run_single_walk('iso', mg, 4)


def test_empty_graph_and_isolated_nodes():
    mg = nx.MultiGraph()
    mg.add_nodes_from(['x', 'y'])
    dw = DeepWalk(mg)
    dw.get_walks()
    assert len(dw.walks) == 0
    mg.add_edge('x', 'z')
    dw = run_walks(mg, niter=10)
    assert len(dw.walks) == 20
    assert sorted(dw.model.wv.index_to_key) == ['x', 'z']
    with pytest.raises(KeyError):
        dw.model.wv['y']            # isolated nodes never enter the vocabulary (min_count=1)
    with pytest.raises(KeyError):
        dw.model.wv.similarity('x', 'y')


# ---- word2vec -----------------------------------------------------------------------------------
def test_word2vec_rejects_unsupported_settings():
    mg, _ = graphs.small(seed=2)
    dw = DeepWalk(mg, 10, 2)
    dw.get_walks()
    for kw in ({'sg': 0}, {'window': 2}, {'sample': 1e-5}, {'min_count': 5}, {'hs': 1}):
        with pytest.raises(ValueError):
            dw.word2vec(**kw)
    with pytest.raises(NotImplementedError):
        dw.word2vec(vector_size=12)
    with pytest.raises(TypeError):
        dw.word2vec(bogus=1)


def test_word2vec_model_surface_and_pickle():
    mg, _ = graphs.small(seed=4)
    dw = run_walks(mg, niter=20, walk_seed=11, vector_size=8, workers=3)
    wv = dw.model.wv
    n_vocab = sum(1 for v in mg if len(mg[v]) > 0)
    assert len(wv) == n_vocab and wv.vectors.shape == (n_vocab, 8) and wv.vectors.dtype == np.float32
    # index_to_key is ordered by descending count, like gensim
    counts = [wv.get_vecattr(k, 'count') for k in wv.index_to_key]
    assert counts == sorted(counts, reverse=True)
    # expected token count of v is niter*L*len(graph[v]) (stationarity)
    tot = sum(counts)
    assert tot == len(dw.walks) * 10
    k0 = wv.index_to_key[0]
    assert abs(counts[0] / tot - len(mg[k0]) / sum(len(mg[v]) for v in mg)) < 0.01
    a, b = wv.index_to_key[0], wv.index_to_key[1]
    s = wv.similarity(a, b)
    want = stats_ref.cosine_similarity(wv.vectors, [0], [1])[0]
    assert abs(s - want) <= 1e-5
    dist = wv.distances(a, other_words=[b, a])
    assert abs(dist[0] - (1 - want)) <= 1e-5 and abs(dist[1]) <= 1e-5
    assert len(wv.distances(a)) == len(wv)
    nv2 = pickle.loads(pickle.dumps(copy.deepcopy(wv)))        # cli.py:208-210
    assert nv2.index_to_key == wv.index_to_key and np.array_equal(nv2.vectors, wv.vectors)
    assert abs(nv2.similarity(a, b) - s) <= 1e-6
    dw2 = pickle.loads(pickle.dumps(dw))                       # --save_dw, cli.py:206-207
    assert list(dw2.walks[:5]) == list(dw.walks[:5]) and len(dw2.walks) == len(dw.walks)


def test_word2vec_on_plain_list_corpus_ragged():
    walks = [['a', 'b', 'c'], ['b', 'a'], ['c', 'b', 'a', 'b'], ['d']] * 50
    dw = DeepWalk(nx.MultiGraph())
    dw.walks = walks
    dw.word2vec(lanes=1, sgns_seed=3)
    wv = dw.model.wv
    assert wv.index_to_key[0] == 'b' and sorted(wv.index_to_key) == ['a', 'b', 'c', 'd']
    assert [wv.get_vecattr(k, 'count') for k in ('a', 'b', 'c', 'd')] == [150, 200, 100, 50]
    # 'd' is never in a pair: its vector keeps its initial value
    rows, _ = olib.gensim_init_weights(4, 8)
    assert np.array_equal(wv['d'], rows[wv.key_to_index['d']])


def test_embedding_recovers_planted_structure():
    """Default-shape Hogwild training: neighbours end up more similar than random pairs."""
    mg, genes = graphs.modular_gene_go_network(300, 60, 8, 1500, 900, 100, seed=5)
    dw = run_walks(mg, walk_seed=1)
    wv = dw.model.wv
    csr = GraphCSR.from_networkx(mg)
    full = np.zeros((csr.n, 8), np.float32)
    full[wv.node_ids] = wv.vectors
    nbr = stats_ref.null_similarities(csr.row_ptr, csr.col_idx, full)
    rng = np.random.default_rng(0)
    a, b = rng.integers(0, len(wv), 5000), rng.integers(0, len(wv), 5000)
    rnd = stats_ref.cosine_similarity(wv.vectors, a, b)
    assert np.isfinite(wv.vectors).all()
    # the sequential gensim-faithful oracle gives 0.62 vs 0.34 on this network
    assert nbr.mean() > rnd.mean() + 0.15 and 0.5 < nbr.mean() < 0.75


# ---- null model ------------------------------------------------------------------------------------
with open(os.path.join(GOLD, 'rand_graph.json')) as f:
    RAND_CASES = json.load(f)


@pytest.mark.parametrize('case', RAND_CASES, ids=[c['name'] for c in RAND_CASES])
def test_get_rand_graph_matches_reference_semantics(case):
    mg = load_graph(case['graph'])
    random.seed(3)
    rg = get_rand_graph(mg)
    assert isinstance(rg, nx.MultiGraph)
    assert list(rg.nodes()) == case['nodes']
    assert [rg.degree(n) for n in rg.nodes()] == case['degrees']      # incl. loops counted twice
    assert rg.number_of_edges() == case['n_edges'] == mg.number_of_edges()
    random.seed(3)
    rg2 = get_rand_graph(mg)
    assert sorted(rg.edges()) == sorted(rg2.edges())                   # governed by random.seed
    rg3 = get_rand_graph(mg)
    assert sorted(rg.edges()) != sorted(rg3.edges())
    # the CSR that rides along equals the materialised graph's distinct-neighbour sets
    csr = rg.graph['_gw_csr'].to_host()
    for i, v in enumerate(csr.nodes):
        got = sorted(csr.nodes[j] for j in csr.col_idx[csr.row_ptr[i]:csr.row_ptr[i + 1]])
        assert got == sorted(rg[v])
    loops = [nx.number_of_selfloops(get_rand_graph(mg)) for _ in range(20)]
    assert abs(np.mean(loops) - case['mean_selfloops']) <= 3 * np.sqrt(case['mean_selfloops'] / 10 + 0.1)
    # walks + vectors on the randomized graph work through both forms
    dw = run_walks(rg, niter=5)
    assert set(dw.model.wv.index_to_key) == {v for v in rg if len(rg[v]) > 0}


def test_get_null_distributions_matches_reference():
    g = np.load(os.path.join(GOLD, 'null_dist.npz'))
    nodes = [str(x) for x in g['nodes']]
    index = {v: i for i, v in enumerate(nodes)}
    csr = GraphCSR(nodes, g['adj_ptr'], [index[str(w)] for w in g['adj']], index)
    nv = NodeVectors([str(k) for k in g['keys']], g['vectors'])
    srd = get_null_distributions(csr, nv)
    assert len(srd) == len(g['srd'])
    assert np.abs(np.asarray(srd, dtype=np.float32) - g['srd']).max() <= 1e-5   # same order too
    acc = []
    acc += srd                                        # cli.py:235 accumulates with +=
    assert np.array_equal(np.asarray(sorted(acc)), np.sort(np.asarray(srd)))   # cli.py:238


# ---- statistics --------------------------------------------------------------------------------------
with open(os.path.join(GOLD, 'statistics.json')) as f:
    STAT = json.load(f)


def _golden_genewalk():
    z = np.load(os.path.join(GOLD, 'statistics.npz'))
    mg = load_graph(STAT['graph'])
    nvs = [NodeVectors(STAT['keys'][r], z['vec%d' % r]) for r in range(3)]
    return GeneWalk(mg, STAT['genes'], nvs, z['null'], gene_id_type='custom'), z


@pytest.mark.parametrize('alpha', [1, 0.1])
def test_generate_output_matches_reference(alpha):
    gw, z = _golden_genewalk()
    df = gw.generate_output(alpha_fdr=alpha)
    want = STAT['outputs'][str(alpha)]
    assert list(df.columns) == want['columns']
    assert [str(t) for t in df.dtypes] == want['dtypes']
    cols = want['columns']
    got = json.loads(df.to_json(orient='split', double_precision=15))['data']
    assert len(got) == len(want['records'])
    key = lambda r: (r[cols.index('custom')], r[cols.index('go_id')] or '')   # noqa: E731
    got_by = {key(r): r for r in got}
    assert len(got_by) == len(got)
    for w in want['records']:
        g = got_by[key(w)]
        for c, a, b in zip(cols, g, w):
            if isinstance(b, float) and b is not None:
                tol = 1e-5 if c in ('sim', 'sem_sim') else 1e-9
                assert a == pytest.approx(b, rel=tol, abs=tol * 1e-3), (c, key(w))
            else:
                assert a == b, (c, key(w))
    # the sorted row order agrees with the reference wherever the sort keys are not tied
    pos = {key(r): i for i, r in enumerate(got)}
    order = [pos[key(w)] for w in want['records']]
    assert sum(1 for a, b in zip(order[:-1], order[1:]) if a > b) <= len(order) // 50


@pytest.mark.parametrize('id_type', ['hgnc_symbol', 'hgnc_id', 'mgi_id', 'rgd_id', 'ensembl_id',
                                     'entrez_human'])
def test_generate_output_other_id_types(id_type):
    """Column layout and values for the non-custom gene id types (perform_statistics.py:127-145,
    188-214, 226-243)."""
    z = np.load(os.path.join(GOLD, 'statistics.npz'))
    mg = load_graph(STAT['graph'])
    nvs = [NodeVectors(STAT['keys'][r], z['vec%d' % r]) for r in range(3)]
    want = STAT['id_types'][id_type]
    gw = GeneWalk(mg, want['genes'], nvs, z['null'], gene_id_type=id_type)
    df = gw.generate_output(alpha_fdr=1)
    assert list(df.columns) == want['columns']
    assert [str(t) for t in df.dtypes] == want['dtypes']
    cols = want['columns']
    got = json.loads(df.to_json(orient='split', double_precision=15))['data']
    assert len(got) == len(want['records'])
    key = lambda r: (r[cols.index('hgnc_symbol')], r[cols.index('go_id')] or '')   # noqa: E731
    got_by = {key(r): r for r in got}
    for w in want['records']:
        g = got_by[key(w)]
        for c, a, b in zip(cols, g, w):
            if isinstance(b, float):
                tol = 1e-5 if c in ('sim', 'sem_sim') else 1e-9
                assert a == pytest.approx(b, rel=tol, abs=tol * 1e-3), (c, key(w))
            else:
                assert a == b, (c, key(w))


def test_psim_and_log_stats_match_reference():
    gw, z = _golden_genewalk()
    for s, p in zip(STAT['psim']['sims'], STAT['psim']['pvals']):
        assert gw.psim(np.float32(s)) == p
    for vals, want in zip(STAT['log_stats']['vals'], STAT['log_stats']['out']):
        assert np.allclose(gw.log_stats(vals), want, rtol=1e-10, atol=1e-300)


# ---- whole pipeline ------------------------------------------------------------------------------------
def test_run_genewalk_pipeline_consistent_with_oracle_statistics():
    """node_vectors + null_distribution + statistics on a small modular network (cli.py:194-252);
    the table is re-derived from the returned similarities with the numpy oracle."""
    from genewalk_b200.replicates import run_genewalk
    mg, genes = graphs.modular_gene_go_network(300, 60, 8, 1500, 900, 100, seed=5)
    genes = genes + [{'ID': 'NOT_THERE'}]
    df, parts = run_genewalk(mg, genes, nreps_graph=2, nreps_null=2, alpha_fdr=1,
                             gene_id_type='custom', base_seed=7, return_parts=True)
    assert list(df.columns)[:4] == ['custom', 'go_name', 'go_id', 'go_domain']
    assert 'global_padj' in df.columns and 'pval_rep0' not in df.columns
    rows = df[df['go_id'] != '']
    gw = GeneWalk(mg, genes, [], parts['null'], gene_id_type='custom')
    pairs = gw.collect_pairs()
    assert len(rows) == len(pairs['pair_go'])
    null = parts['null']
    assert np.all(np.diff(null) >= 0) and len(null) > 1000
    sims = parts['sims'].cpu().numpy()
    padj = stats_ref.gene_padj(sims, null, pairs['seg_ptr'])
    nodes = list(mg.nodes())
    want = {(nodes[g], nodes[t]): p for g, t, p in zip(pairs['pair_gene'], pairs['pair_go'], padj)}
    for _, r in rows.iterrows():
        assert r['gene_padj'] == pytest.approx(want[(r['custom'], r['go_id'])], rel=1e-9, abs=1e-18)
    # planted structure: a good share of the module-specific annotations is significant
    assert (rows['gene_padj'] < 0.1).mean() > 0.2
    # same seed -> same similarities at lanes=1 would be bit-exact; at full width they are close
    df2 = run_genewalk(mg, genes, nreps_graph=2, nreps_null=2, alpha_fdr=0.1,
                       gene_id_type='custom', base_seed=7)
    assert len(df2) <= len(df) and (df2['gene_padj'] < 0.1).all()


def test_null_distribution_parity_with_gensim_faithful_oracle():
    """north_star: KS agreement of the null similarity distribution with the reference's gensim
    path on the same input (same randomized graph, same walks; the trainers differ)."""
    from scipy.stats import ks_2samp
    from genewalk_b200.replicates import null_replicate, mix_seed
    from oracle import config_model_ref
    mg, _ = graphs.modular_gene_go_network(300, 60, 8, 1500, 900, 100, seed=105)
    got, _ = null_replicate(mg, 11, 0)
    got = got.cpu().numpy()
    d_seq = config_model_ref.multigraph_degrees_desc(multigraph_degrees(mg))
    u, v = config_model_ref.config_model_edges(d_seq, mix_seed(11, 2, 0))
    rp, ci = config_model_ref.csr_distinct(len(d_seq), u, v)
    tok = olib.walk_uniform(rp, ci, 100, 10, mix_seed(11, 3, 0), 0)
    itn, syn0, _ = olib.word2vec_gensim_like(tok, len(d_seq), nthreads=1)
    vec = np.zeros((len(d_seq), 8), np.float32)
    vec[itn] = syn0
    want = stats_ref.null_similarities(rp, ci, vec)
    assert len(got) == len(want)
    # two independent reference runs differ by KS ~0.02-0.09 (profiles/parity_study_r01.md)
    assert ks_2samp(got, want).statistic < 0.1
    assert abs(got.mean() - want.mean()) < 0.05


# ---- concurrent replicates (several trained at once on separate CUDA streams) ---------------------------
def test_word2vec_enqueue_then_finish():
    """word2vec(wait=False) only enqueues; finish_word2vec() yields the model of a blocking call."""
    mg, _ = graphs.modular_gene_go_network(120, 30, 4, 500, 300, 40, seed=3)
    a = DeepWalk(mg, 10, 20, seed=5)
    a.get_walks()
    a.word2vec(lanes=1, sgns_seed=9)
    b = DeepWalk(mg, 10, 20, seed=5)
    b.get_walks()
    assert b.word2vec(lanes=1, sgns_seed=9, wait=False) is None
    assert b.model is None
    b.finish_word2vec()
    assert b.model.wv.index_to_key == a.model.wv.index_to_key
    assert np.array_equal(b.model.wv.vectors, a.model.wv.vectors)   # sequential mode is bit-exact


def test_run_walks_many_equals_serial_replicates():
    """Replicates trained at once on different streams are the replicates trained one by one
    (sequential SGNS mode: bit-exact), and distinct replicates differ."""
    from genewalk_b200.replicates import run_walks_many
    mg, _ = graphs.modular_gene_go_network(150, 40, 5, 700, 400, 50, seed=8)
    conc = run_walks_many(mg, 4, base_seed=21, concurrent=3, niter=20, lanes=1)
    ser = run_walks_many(mg, 4, base_seed=21, concurrent=1, niter=20, lanes=1)
    assert len(conc) == len(ser) == 4
    for c, s in zip(conc, ser):
        assert np.array_equal(np.asarray(c.walks.tokens.cpu()), np.asarray(s.walks.tokens.cpu()))
        assert c.model.wv.index_to_key == s.model.wv.index_to_key
        assert np.array_equal(c.model.wv.vectors, s.model.wv.vectors)
    assert not np.array_equal(conc[0].model.wv.vectors, conc[1].model.wv.vectors)


def test_run_genewalk_concurrency_does_not_change_inputs_or_statistics():
    """run_genewalk(concurrent=3) vs concurrent=1 at lanes=1: identical similarities, null
    distribution and table; and the number of replicates in flight never changes the shape of a
    training (same workers, same plan: results do not depend on GPU fill or GPU count)."""
    from genewalk_b200.replicates import run_genewalk
    mg, genes = graphs.modular_gene_go_network(120, 30, 4, 500, 300, 40, seed=13)
    kw = dict(nreps_graph=2, nreps_null=3, alpha_fdr=1, gene_id_type='custom', base_seed=3,
              return_parts=True, lanes=1)
    d1, p1 = run_genewalk(mg, genes, concurrent=1, **kw)
    d3, p3 = run_genewalk(mg, genes, concurrent=3, **kw)
    assert np.array_equal(p1['sims'].cpu().numpy(), p3['sims'].cpu().numpy())
    assert np.array_equal(p1['null'], p3['null'])
    assert d1.equals(d3)
    from genewalk_b200.replicates import run_walks_many
    shapes = []
    for conc in (1, 4):
        dws = run_walks_many(mg, 4, base_seed=2, concurrent=conc, niter=20)
        shapes.append({(dw.train_shape['lanes'], dw.train_shape['unit_walks'], dw.train_shape['policy'])
                       for dw in dws})
    assert shapes[0] == shapes[1] and len(shapes[0]) == 1
