"""Stage driver (genewalk_b200.cli): the reference's argument checks (genewalk/cli.py:172-185,
tests/test_cli.py:149-156) on CPU; the staged run with the reference's file set on GPU."""
import os
import pickle

import numpy as np
import pandas as pd
import pytest

from genewalk_b200 import cli
from genewalk_b200 import workloads as graphs


def _args(tmp, **kw):
    base = ['--project', 'p', '--genes', os.path.join(tmp, 'genes.txt'), '--base_folder', tmp]
    for k, v in kw.items():
        base += ['--' + k, str(v)]
    return cli.build_parser().parse_args(base)


def test_parser_defaults_match_reference():
    a = cli.build_parser().parse_args(['--project', 'x', '--genes', 'g.txt'])
    assert (a.nreps_graph, a.nreps_null, a.alpha_fdr, a.dim_rep, a.nproc) == (3, 3, 1, 8, 1)
    assert a.stage == 'all' and a.id_type == 'hgnc_symbol' and a.network_source == 'pc'
    assert a.random_seed is None and a.save_dw is False


def test_missing_network_file(tmp_path):
    with pytest.raises(ValueError):          # reference tests/test_cli.py:149-156
        cli.run_main(_args(str(tmp_path), network_source='sif', id_type='hgnc_symbol'))


def test_custom_id_type_needs_full_network(tmp_path):
    with pytest.raises(ValueError):
        cli.run_main(_args(str(tmp_path), network_source='pc', id_type='custom'))


@pytest.mark.gpu
def test_staged_run_writes_reference_file_set(tmp_path):
    tmp = str(tmp_path)
    mg, genes = graphs.modular_gene_go_network(200, 40, 6, 900, 500, 60, seed=3)
    with open(os.path.join(tmp, 'mg.pkl'), 'wb') as fh:
        pickle.dump(mg, fh)
    with open(os.path.join(tmp, 'genes.txt'), 'w') as fh:
        fh.write('\n'.join(g['ID'] for g in genes[:150]) + '\nNOT_IN_NETWORK\n')
    common = dict(network_source='pickle', network_file=os.path.join(tmp, 'mg.pkl'),
                  id_type='custom', nreps_graph=2, nreps_null=2, random_seed=5)
    for stage in ('node_vectors', 'null_distribution', 'statistics'):
        cli.run_main(_args(tmp, stage=stage, **common))
    proj = os.path.join(tmp, 'p')
    for f in ('genes.pkl', 'multi_graph.pkl', 'deepwalk_node_vectors_1.pkl',
              'deepwalk_node_vectors_2.pkl', 'deepwalk_node_vectors_rand_1.pkl',
              'deepwalk_node_vectors_rand_2.pkl', 'genewalk_rand_simdists.pkl',
              'genewalk_results.csv', 'genewalk_statistics.log'):
        assert os.path.exists(os.path.join(proj, f)), f
    srd = cli.load_pickle(proj, 'genewalk_rand_simdists')
    assert isinstance(srd, np.ndarray) and srd.dtype == np.float32 and np.all(np.diff(srd) >= 0)
    nv = cli.load_pickle(proj, 'deepwalk_node_vectors_1')
    assert nv.vectors.shape[1] == 8 and len(nv.index_to_key) == nv.vectors.shape[0]
    df = pd.read_csv(os.path.join(proj, 'genewalk_results.csv'))
    assert list(df.columns) == ['custom', 'go_name', 'go_id', 'go_domain', 'ncon_gene', 'ncon_go',
                                'global_padj', 'gene_padj', 'pval', 'sim', 'sem_sim',
                                'cilow_global_padj', 'ciupp_global_padj', 'cilow_gene_padj',
                                'ciupp_gene_padj', 'cilow_pval', 'ciupp_pval']
    assert 'NOT_IN_NETWORK' in set(df['custom']) and df['go_id'].notna().sum() > 100
    # --stage all in one go with a fresh project gives the same file set
    a = _args(tmp, stage='all', **common)
    a.project = 'q'
    cli.run_main(a)
    assert os.path.exists(os.path.join(tmp, 'q', 'genewalk_results.csv'))


@pytest.mark.gpu
def test_all_stages_from_a_sif_file(tmp_path):
    """--network_source sif_full --id_type custom: the network file goes straight to CSR arrays
    (edgelist.py, no networkx, no reference package), the three stages run, and multi_graph.pkl
    still holds the nx.MultiGraph the reference would have written."""
    import networkx as nx
    tmp = str(tmp_path)
    mg, genes = graphs.modular_gene_go_network(150, 30, 5, 600, 350, 40, seed=4)
    with open(os.path.join(tmp, 'net.sif'), 'w') as fh:
        fh.write('\n'.join('%s,%s,%s' % (u, d.get('label', 'x'), v) for u, v, d in mg.edges(data=True)))
        fh.write('\nUNLISTED1,x,UNLISTED2\n')
    with open(os.path.join(tmp, 'go.obo'), 'w') as fh:
        for v, a in mg.nodes(data=True):
            if 'GO' in a:
                fh.write('[Term]\nid: %s\nname: %s\nnamespace: %s\n\n' % (v, a['name'], a['domain']))
    with open(os.path.join(tmp, 'genes.txt'), 'w') as fh:
        fh.write('\n'.join(g['ID'] for g in genes) + '\n')
    cli.run_main(_args(tmp, stage='all', network_source='sif_full', network_file=os.path.join(tmp, 'net.sif'),
                       go_obo=os.path.join(tmp, 'go.obo'), id_type='custom', nreps_graph=2, nreps_null=2,
                       random_seed=7))
    proj = os.path.join(tmp, 'p')
    saved = cli.load_pickle(proj, 'multi_graph')
    assert isinstance(saved, nx.MultiGraph) and 'UNLISTED1' not in saved
    assert saved.number_of_edges() == mg.number_of_edges()
    df = pd.read_csv(os.path.join(proj, 'genewalk_results.csv'))
    assert df['go_id'].notna().sum() > 100 and df['go_name'].notna().sum() == df['go_id'].notna().sum()
