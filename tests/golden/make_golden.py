"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(/root/reference/genewalk) in the authoring container.

    python tests/golden/make_golden.py

gensim and statsmodels are absent here, so the reference is imported through oracle/ref_shim.py
(two stubbed third-party names; everything under /root/reference runs verbatim).  The fixtures are
small and committed; /root/reference does not exist on the GPU box, so the GPU tests read these
files instead.

  walk_structure.json   reference DeepWalk.get_walks on three small MultiGraphs (incl. the graph of
                        genewalk/tests/test_deepwalk.py): len(walks), walks per start node, walk
                        length, and the empirical transition counts of 20k reference steps out of
                        the highest-degree node.
  rand_graph.json       reference get_rand_graph on two MultiGraphs: node labels, degree list,
                        edge count, self-loop / parallel-edge counts averaged over 20 shuffles.
  null_dist.npz         reference get_null_distributions(rg, nv) for a graph with self-loops and
                        parallel edges and fixed random vectors.
  statistics.npz/.json  reference GeneWalk.generate_output (alpha_fdr 1 and 0.1), psim and
                        log_stats for a synthetic gene-GO network, 3 replicates of fixed random
                        vectors and a fixed null distribution.
"""
import json
import os
import random
import sys
import warnings

import numpy as np
import networkx as nx

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim          # noqa: E402
from genewalk_b200 import workloads as graphs             # noqa: E402

warnings.filterwarnings('ignore')


def edges_of(mg):
    return [[u, v] for u, v in mg.edges()]


def graph_spec(mg):
    return {'nodes': [[n, dict(mg.nodes[n])] for n in mg.nodes()], 'edges': edges_of(mg)}


def main():
    dw, nd, ps = ref_shim.import_reference()
    import logging
    logging.getLogger('genewalk').setLevel(logging.WARNING)
    logging.getLogger().setLevel(logging.WARNING)

    # ---- walks -------------------------------------------------------------------------------
    out = []
    path = nx.MultiGraph()
    path.add_edge('a', 'b')
    path.add_edge('b', 'c')
    multi = nx.MultiGraph()
    multi.add_edges_from([('a', 'b'), ('a', 'b'), ('a', 'b'), ('a', 'a'), ('b', 'c'), ('c', 'd'),
                          ('d', 'a'), ('d', 'd'), ('e', 'a')])
    multi.add_node('iso')
    small, _ = graphs.small(seed=3)
    for name, mg, niter, wl in (('path_test_deepwalk', path, 100, 10), ('multi_loops', multi, 50, 7),
                                ('small_gene_go', small, 3, 10)):
        random.seed(12345)
        d = dw.DeepWalk(mg, wl, niter)
        d.get_walks(workers=1)
        starts = {}
        for w in d.walks:
            starts[w[0]] = starts.get(w[0], 0) + 1
        hub = max(mg.nodes(), key=lambda v: len(mg[v]))
        trans = {}
        for _ in range(20000):
            nxt = dw.run_single_walk(hub, mg, 2)[1]
            trans[nxt] = trans.get(nxt, 0) + 1
        out.append({'name': name, 'graph': graph_spec(mg), 'niter': niter, 'walk_length': wl,
                    'n_walks': len(d.walks), 'starts': starts,
                    'lengths': sorted(set(len(w) for w in d.walks)), 'hub': hub,
                    'hub_transitions': trans})
    with open(os.path.join(HERE, 'walk_structure.json'), 'w') as f:
        json.dump(out, f, indent=0)

    # ---- get_rand_graph --------------------------------------------------------------------
    out = []
    for name, mg in (('multi_loops', multi), ('small_gene_go', small)):
        random.seed(777)
        loops, par, nedges = [], [], []
        for _ in range(20):
            rg = nd.get_rand_graph(mg)
            loops.append(nx.number_of_selfloops(rg))
            par.append(rg.number_of_edges() - len(set(tuple(sorted(e)) for e in rg.edges())))
            nedges.append(rg.number_of_edges())
        out.append({'name': name, 'graph': graph_spec(mg), 'nodes': list(rg.nodes()),
                    'degrees': [rg.degree(n) for n in rg.nodes()],
                    'n_edges': nedges[0], 'mean_selfloops': float(np.mean(loops)),
                    'mean_parallel': float(np.mean(par))})
    with open(os.path.join(HERE, 'rand_graph.json'), 'w') as f:
        json.dump(out, f, indent=0)

    # ---- get_null_distributions ----------------------------------------------------------------
    random.seed(99)
    rg = nd.get_rand_graph(small)
    rg.add_edge('n0', 'n0')
    rg.add_edge('n1', 'n2')
    rg.add_edge('n1', 'n2')
    rng = np.random.default_rng(5)
    keys = [k for k in rg.nodes() if len(rg[k]) > 0]
    rng.shuffle(keys)
    vec = rng.standard_normal((len(keys), 8)).astype(np.float32)
    nv = ref_shim.ShimKeyedVectors(keys, vec)
    srd = nd.get_null_distributions(rg, nv)
    # adjacency lists in networkx order (list(rg[node])): the output order depends on it
    adj_ptr = np.cumsum([0] + [len(rg[n]) for n in rg.nodes()])
    adj = np.array([w for n in rg.nodes() for w in rg[n]])
    np.savez(os.path.join(HERE, 'null_dist.npz'), nodes=np.array(list(rg.nodes())),
             adj_ptr=adj_ptr, adj=adj, keys=np.array(keys), vectors=vec,
             srd=np.asarray(srd, dtype=np.float32))

    # ---- statistics ------------------------------------------------------------------------------
    mg, genes = graphs.gene_go_network(60, 25, 150, 200, 30, seed=8)
    genes = genes + [{'ID': 'NOT_IN_GRAPH'}]
    mg.add_node('G_NO_GO')            # gene in graph without GO connections
    genes.append({'ID': 'G_NO_GO'})
    rng = np.random.default_rng(42)
    vocab = [n for n in mg.nodes() if len(mg[n]) > 0]
    nvs, vecs, key_lists = [], [], []
    for r in range(3):
        keys = list(vocab)
        rng.shuffle(keys)
        base = rng.standard_normal((len(keys), 8)).astype(np.float32)
        # make neighbours similar so that some pairs are significant
        idx = {k: i for i, k in enumerate(keys)}
        for u, v in mg.edges():
            if u in idx and v in idx:
                base[idx[v]] += np.float32(0.35) * base[idx[u]]
        nvs.append(ref_shim.ShimKeyedVectors(keys, base))
        vecs.append(base)
        key_lists.append(keys)
    null = np.sort((rng.beta(2, 2, size=4000) * 2 - 1).astype(np.float32))
    GW = ps.GeneWalk(mg, genes, nvs, null, gene_id_type='custom')
    spec = {'graph': graph_spec(mg), 'genes': genes, 'keys': key_lists, 'outputs': {}}
    for alpha in (1, 0.1):
        df = GW.generate_output(alpha_fdr=alpha)
        spec['outputs'][str(alpha)] = {'columns': list(df.columns),
                                       'dtypes': [str(t) for t in df.dtypes],
                                       'records': json.loads(df.to_json(orient='split',
                                                                        double_precision=15))['data']}
    probe = np.concatenate([null[::400], [-1.0, 1.0, 0.123456]]).astype(np.float32)
    spec['psim'] = {'sims': probe.tolist(), 'pvals': [GW.psim(s) for s in probe]}
    ls_in = [[0.5], [1e-16, 1e-16], [0.01, 0.02, 0.5], [1.0, 1.0, 1.0], [0.3, 1e-8, 0.9, 0.2]]
    spec['log_stats'] = {'vals': ls_in, 'out': [[float(x) for x in GW.log_stats(v)] for v in ls_in]}
    # the other gene id types: same network, genes described the way gene_lists.py does
    # (HGNC_SYMBOL / HGNC plus the species-specific id), first 25 genes + one gene not in the network
    spec['id_types'] = {}
    for id_type, extra in (('hgnc_symbol', None), ('hgnc_id', None), ('mgi_id', 'MGI'),
                           ('rgd_id', 'RGD'), ('ensembl_id', 'ENSEMBL'), ('entrez_human', 'EGID')):
        gl = []
        for k, g in enumerate(genes[:25] + [{'ID': 'NOT_IN_GRAPH'}]):
            d = {'HGNC_SYMBOL': g['ID'], 'HGNC': str(1000 + k)}
            if extra:
                d[extra] = '%s:%d' % (extra, 5000 + k)
            gl.append(d)
        GWt = ps.GeneWalk(mg, gl, nvs, null, gene_id_type=id_type)
        dft = GWt.generate_output(alpha_fdr=1)
        spec['id_types'][id_type] = {'genes': gl, 'columns': list(dft.columns),
                                     'dtypes': [str(t) for t in dft.dtypes],
                                     'records': json.loads(dft.to_json(orient='split',
                                                                       double_precision=15))['data']}
    with open(os.path.join(HERE, 'statistics.json'), 'w') as f:
        json.dump(spec, f, indent=0)
    np.savez(os.path.join(HERE, 'statistics.npz'), null=null, **{'vec%d' % r: vecs[r] for r in range(3)})
    print('golden fixtures written to', HERE)


if __name__ == '__main__':
    main()
