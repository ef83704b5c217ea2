"""Network ingestion without networkx (genewalk_b200/edgelist.py) against
(a) the golden output of the UNMODIFIED reference's UserNxMgAssembler on the reference's own test
networks (tests/golden/sif_networks.json, nx_mg_assembler.py:342-422) and (b) the reference's
construction steps restated with the same library calls (nx.from_pandas_edgelist + node removal)
on a synthetic network with parallel edges, self-loops and unlisted genes.  CPU only."""
import json
import os

import networkx as nx
import numpy as np
import pandas as pd
import pytest

from genewalk_b200 import workloads
from genewalk_b200.csr import GraphCSR, multigraph_degrees
from genewalk_b200.edgelist import network_from_edges, read_network, read_obo_terms
from genewalk_b200.perform_statistics import GeneWalk

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
with open(os.path.join(GOLD, 'sif_networks.json')) as f:
    SIF = json.load(f)


@pytest.mark.parametrize('case', SIF['cases'], ids=[c['name'] for c in SIF['cases']])
def test_read_network_matches_reference_assembler(case, tmp_path):
    sif = tmp_path / 'net.sif'
    sif.write_text('\n'.join(','.join(r) for r in case['rows']))
    obo = tmp_path / 'go.obo'
    obo.write_text(SIF['obo'])
    net = read_network(str(sif), 'sif_full', case['genes'], obo=str(obo))
    assert list(net.nodes()) == case['nodes']
    assert {v: net[v] for v in net.nodes()} == case['adjacency']
    assert net.degrees.tolist() == case['degrees']
    assert net.number_of_edges() == case['n_edges']
    assert {v: net.nodes[v] for v in net.nodes() if net.nodes[v]} == case['attrs']
    assert net.go_node_labels() == set(case['attrs'])
    # and the equal nx.MultiGraph
    mg = net.to_networkx()
    assert list(mg.nodes()) == case['nodes'] and mg.number_of_edges() == case['n_edges']
    assert {v: list(mg[v]) for v in mg.nodes()} == case['adjacency']


def test_obo_terms():
    import tempfile
    with tempfile.NamedTemporaryFile('w', suffix='.obo', delete=False) as f:
        f.write(SIF['obo'] + '\n[Typedef]\nid: part_of\nname: part of\n')
    terms = read_obo_terms(f.name)
    os.unlink(f.name)
    assert terms['GO:0000186'] == ('GO:0000186', 'activation of MAPKK activity', 'biological_process')
    assert 'GO:9999999' not in terms and 'part_of' not in terms


def reference_steps(df, genes):
    """nx_mg_assembler.py:383-401, restated with the same calls."""
    g = nx.from_pandas_edgelist(df, 'source', 'target', edge_attr=True, create_using=nx.MultiGraph)
    listed = {(x['ID'] if 'ID' in x else x['HGNC_SYMBOL']) for x in genes}
    g.remove_nodes_from([n for n in g if not n.startswith('GO:') and n not in listed])
    return g


def test_network_from_edges_equals_networkx_construction():
    mg0, genes = workloads.modular_gene_go_network(300, 60, 8, 1500, 900, 100, seed=5)
    rows = [(u, 'rel', v) for u, v in mg0.edges()]
    rows += [('G00001', 'x', 'G00001'), ('XTRA1', 'x', 'G00002'), ('XTRA1', 'x', 'XTRA2'),
             ('G00005', 'x', 'GO:9000003'), ('G00005', 'y', 'GO:9000003')]
    rng = np.random.default_rng(0)
    rows = [rows[i] for i in rng.permutation(len(rows))]
    df = pd.DataFrame(rows, columns=['source', 'rel_type', 'target'])
    g = reference_steps(df, genes)
    net = network_from_edges(df['source'].to_numpy(), df['target'].to_numpy(), genes)
    ref = GraphCSR.from_networkx(g)
    assert net.csr.nodes == ref.nodes and 'XTRA1' not in net
    assert np.array_equal(net.csr.row_ptr, ref.row_ptr) and np.array_equal(net.csr.col_idx, ref.col_idx)
    assert np.array_equal(multigraph_degrees(net), multigraph_degrees(g))
    assert net.number_of_edges() == g.number_of_edges()
    # no gene list: nothing is removed
    full = network_from_edges(df['source'].to_numpy(), df['target'].to_numpy(), None)
    assert 'XTRA2' in full and len(full) == len(g) + 2
    # the statistics bookkeeping finds the same (gene, GO) pairs on either representation
    for v in g.nodes:
        if v.startswith('GO:'):
            g.nodes[v].update(GO=v, name='t', domain='biological_process')
    net.annotate_go({k: (k, 't', 'biological_process') for k in net.nodes() if k.startswith('GO:')})
    genes2 = genes + [{'ID': 'NOT_THERE'}]
    a = GeneWalk(g, genes2, [], np.zeros(1, np.float32), 'custom').collect_pairs()
    b = GeneWalk(net, genes2, [], np.zeros(1, np.float32), 'custom').collect_pairs()
    for k in ('pair_gene', 'pair_go', 'seg_ptr'):
        assert np.array_equal(a[k], b[k]), k
    assert a['gene_of_seg'] == b['gene_of_seg']
    for x, y in zip(a['gene_attribs'], b['gene_attribs']):
        assert x['custom_id'] == y['custom_id']
        assert x['ncon_gene'] == y['ncon_gene'] or (np.isnan(x['ncon_gene']) and np.isnan(y['ncon_gene']))


def test_bad_format_raises(tmp_path):
    p = tmp_path / 'x.sif'
    p.write_text('a,b,c\n')
    with pytest.raises(ValueError):
        read_network(str(p), 'csv')


def test_cli_loads_sif_full_natively(tmp_path):
    """genewalk_b200.cli reads a complete custom-id SIF without the reference package
    (cli.py:195-197 -> nx_mg_assembler.load_network) and would pickle an nx.MultiGraph."""
    import pickle
    from genewalk_b200 import cli
    case = [c for c in SIF['cases'] if c['name'] == 'sif_custom_filter'][0]
    sif = tmp_path / 'net.sif'
    sif.write_text('\n'.join(','.join(r) for r in case['rows']))
    obo = tmp_path / 'go.obo'
    obo.write_text(SIF['obo'])
    genes = tmp_path / 'genes.txt'
    genes.write_text('\n'.join(g['ID'] for g in case['genes']) + '\n')
    args = cli.build_parser().parse_args(['--project', 'p', '--genes', str(genes), '--id_type', 'custom',
                                          '--network_source', 'sif_full', '--network_file', str(sif),
                                          '--go_obo', str(obo), '--base_folder', str(tmp_path)])
    gene_list, net = cli._load_inputs(args, str(tmp_path))
    assert gene_list == case['genes'] and list(net.nodes()) == case['nodes']
    mg = pickle.loads(pickle.dumps(net.to_networkx()))
    assert {v: list(mg[v]) for v in mg.nodes()} == case['adjacency']
    assert {v: dict(mg.nodes[v]) for v in mg.nodes() if mg.nodes[v]} == case['attrs']
