"""Deterministic synthetic GeneWalk networks; shared by tests and bench.py."""
import numpy as np
import networkx as nx

DOMAINS = ('biological_process', 'molecular_function', 'cellular_component')


def gene_go_network(n_genes, n_go, n_gene_edges, n_annot, n_isa, seed, dup_frac=0.1,
                    power_law=False):
    """nx.MultiGraph with gene nodes 'G%05d', GO nodes 'GO:9%06d' (attrs GO/name/domain),
    gene-gene edges (a fraction duplicated with a second rel_type -> parallel edges), gene-GO
    annotation edges and GO-GO is_a edges.  Returns (graph, genes) with genes = [{'ID': ...}]."""
    rng = np.random.default_rng(seed)
    genes = ['G%05d' % i for i in range(n_genes)]
    gos = ['GO:9%06d' % i for i in range(n_go)]
    mg = nx.MultiGraph()
    mg.add_nodes_from(genes)
    for i, g in enumerate(gos):
        mg.add_node(g, GO=g, name='term %d' % i, domain=DOMAINS[i % 3])
    if power_law:
        w = (np.arange(1, n_genes + 1, dtype=np.float64)) ** (-1.0 / 1.5)
        w /= w.sum()
        a = rng.choice(n_genes, size=n_gene_edges, p=w)
        b = rng.choice(n_genes, size=n_gene_edges, p=w)
    else:
        a = rng.integers(0, n_genes, size=n_gene_edges)
        b = rng.integers(0, n_genes, size=n_gene_edges)
    keep = a != b
    a, b = a[keep], b[keep]
    mg.add_edges_from(((genes[x], genes[y], {'label': 'interacts'}) for x, y in zip(a, b)))
    ndup = int(len(a) * dup_frac)
    mg.add_edges_from(((genes[x], genes[y], {'label': 'regulates'})
                       for x, y in zip(a[:ndup], b[:ndup])))
    # annotations: Zipf-like preference over terms
    tw = (np.arange(1, n_go + 1, dtype=np.float64)) ** -1.2
    tw /= tw.sum()
    ga = rng.integers(0, n_genes, size=n_annot)
    gt = rng.choice(n_go, size=n_annot, p=tw)
    mg.add_edges_from(((genes[x], gos[y], {'label': 'GO:annotation'}) for x, y in zip(ga, gt)))
    # is_a: term i -> parent j < i
    child = rng.integers(1, n_go, size=n_isa)
    parent = (rng.random(n_isa) * child).astype(np.int64)
    mg.add_edges_from(((gos[c], gos[p], {'label': 'GO:is_a'}) for c, p in zip(child, parent)))
    return mg, [{'ID': g} for g in genes]


def config_c1(seed=1):
    """C1: 2k genes, 300 GO terms, ~20k edges (BASELINE.json configs[0])."""
    return gene_go_network(2000, 300, 12000, 7000, 1000, seed)


def small(seed=0, n_genes=120, n_go=30):
    return gene_go_network(n_genes, n_go, 400, 300, 40, seed)


def csr_arrays(n, n_edges, seed, power_law=True, self_loops=0):
    """Random distinct-neighbour CSR (numpy, no networkx): for large-size tests and the bench."""
    rng = np.random.default_rng(seed)
    if power_law:
        w = (np.arange(1, n + 1, dtype=np.float64)) ** (-1.0 / 1.5)
        w /= w.sum()
        a = rng.choice(n, size=n_edges, p=w)
        b = rng.choice(n, size=n_edges, p=w)
    else:
        a = rng.integers(0, n, size=n_edges)
        b = rng.integers(0, n, size=n_edges)
    keep = a != b
    a, b = a[keep], b[keep]
    if self_loops:
        s = rng.integers(0, n, size=self_loops)
        a = np.concatenate([a, s]); b = np.concatenate([b, s])
    key = np.unique(np.concatenate([a * n + b, b * n + a]))
    src = (key // n).astype(np.int32)
    dst = (key % n).astype(np.int32)
    row_ptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(row_ptr, src + 1, 1)
    return np.cumsum(row_ptr).astype(np.int32), dst
