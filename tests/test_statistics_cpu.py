"""CPU test of the host half of GeneWalk.generate_output: the column-wise table assembly
(genewalk_b200/perform_statistics.py:assemble_table) fed with the numeric columns computed by the
numpy oracle reproduces the UNMODIFIED reference's output tables (tests/golden/statistics.*,
perform_statistics.py:146-245) -- columns, dtypes (including pandas' inference quirks), values and
row order -- for alpha_fdr 1 and 0.1 and all gene id types.  The device half is checked by
tests/test_gpu_api.py::test_generate_output_matches_reference."""
import json
import os

import networkx as nx
import numpy as np
import pytest

from genewalk_b200.keyedvectors import NodeVectors
from genewalk_b200.perform_statistics import GeneWalk
from oracle import stats_ref

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
with open(os.path.join(GOLD, 'statistics.json')) as f:
    STAT = json.load(f)


def load_graph(spec):
    mg = nx.MultiGraph()
    for n, attrs in spec['nodes']:
        mg.add_node(n, **attrs)
    mg.add_edges_from(spec['edges'])
    return mg

def host_core(gw, pairs, nvs, null, alpha):
    m = len(pairs['pair_go'])
    nreps = len(nvs)
    nodes = list(gw.graph.nodes())
    sims = np.empty((nreps, m), np.float32)
    for r, nv in enumerate(nvs):
        row = {k: i for i, k in enumerate(nv.index_to_key)}
        a = [row[nodes[i]] for i in pairs['pair_gene']]
        b = [row[nodes[i]] for i in pairs['pair_go']]
        sims[r] = stats_ref.cosine_similarity(nv.vectors, a, b)
    pv = np.stack([stats_ref.psim(null, sims[r]) for r in range(nreps)])
    q = np.stack([stats_ref.fdrcorrection_segmented(pv[r], pairs['seg_ptr']) for r in range(nreps)])
    qs = np.array([stats_ref.log_stats(q[:, k]) for k in range(m)])
    ps = np.array([stats_ref.log_stats(pv[:, k]) for k in range(m)])
    kept = np.arange(m) if alpha == 1 else np.nonzero(qs[:, 0] < alpha)[0]
    gq = np.stack([stats_ref.fdrcorrection(pv[r, kept]) for r in range(nreps)])
    gs = np.array([stats_ref.log_stats(gq[:, k]) for k in range(len(kept))]).reshape(-1, 3)
    return {'kept': kept, 'gene_padj': qs[kept], 'pval': ps[kept],
            'sim': sims.mean(axis=0)[kept], 'sem_sim': (sims.std(axis=0) / np.sqrt(nreps))[kept].astype(np.float64),
            'global_padj': gs}

def check(df, want, keycol):
    assert list(df.columns) == want['columns']
    assert [str(t) for t in df.dtypes] == want['dtypes'], ([str(t) for t in df.dtypes], want['dtypes'])
    cols = want['columns']
    got = json.loads(df.to_json(orient='split', double_precision=15))['data']
    assert len(got) == len(want['records'])
    key = lambda r: (r[cols.index(keycol)], r[cols.index('go_id')] or '')
    got_by = {key(r): r for r in got}
    assert len(got_by) == len(got)
    for w in want['records']:
        g = got_by[key(w)]
        for c, a, b in zip(cols, g, w):
            if isinstance(b, float) and b is not None:
                tol = 1e-5 if c in ('sim', 'sem_sim') else 1e-9
                assert a == pytest.approx(b, rel=tol, abs=tol * 1e-3), (c, key(w))
            else:
                assert a == b, (c, key(w), a, b)
    pos = {key(r): i for i, r in enumerate(got)}
    order = [pos[key(w)] for w in want['records']]
    assert sum(1 for a, b in zip(order[:-1], order[1:]) if a > b) <= len(order) // 50


def _inputs():
    z = np.load(os.path.join(GOLD, 'statistics.npz'))
    mg = load_graph(STAT['graph'])
    nvs = [NodeVectors(STAT['keys'][r], z['vec%d' % r]) for r in range(3)]
    return mg, nvs, z['null']


@pytest.mark.parametrize('alpha', [1, 0.1])
def test_assemble_table_matches_reference(alpha):
    mg, nvs, null = _inputs()
    gw = GeneWalk(mg, STAT['genes'], nvs, null, gene_id_type='custom')
    pairs = gw.collect_pairs()
    df = gw.assemble_table(pairs, host_core(gw, pairs, nvs, null, alpha), alpha, 3)
    check(df, STAT['outputs'][str(alpha)], 'custom')


@pytest.mark.parametrize('id_type', sorted(STAT['id_types']))
def test_assemble_table_other_id_types(id_type):
    mg, nvs, null = _inputs()
    want = STAT['id_types'][id_type]
    gw = GeneWalk(mg, want['genes'], nvs, null, gene_id_type=id_type)
    pairs = gw.collect_pairs()
    df = gw.assemble_table(pairs, host_core(gw, pairs, nvs, null, 1), 1, 3)
    check(df, want, 'hgnc_symbol')


def test_assemble_table_without_pairs():
    """No gene has a GO connection: alpha_fdr = 1 lists every gene with an empty row, a lower
    alpha_fdr gives an empty table with the reference's columns."""
    mg = nx.MultiGraph()
    mg.add_edge('a', 'b')
    gw = GeneWalk(mg, [{'ID': 'a'}, {'ID': 'zz'}], [], np.zeros(1, np.float32), gene_id_type='custom')
    pairs = gw.collect_pairs()
    df = gw.assemble_table(pairs, None, 1, 2)
    assert list(df['custom']) == ['a', 'zz'] and df['ncon_gene'].tolist()[0] == '1.0'
    assert df['ncon_gene'].isna().tolist() == [False, True]      # pandas 3: astype('str') keeps NaN
    assert df['go_id'].tolist() == ['', ''] and df['gene_padj'].isna().all()
    df = gw.assemble_table(pairs, None, 0.1, 2)
    assert len(df) == 0 and 'global_padj' in df.columns and 'hgnc_id' not in df.columns
