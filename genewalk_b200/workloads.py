"""Deterministic synthetic GeneWalk networks: the workloads of bench.py, of the
parity studies and of the tests."""
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


def modular_gene_go_network(n_genes=2000, n_go=300, n_modules=40, n_gene_edges=12000,
                            n_annot=7000, n_isa=1000, p_in=0.8, seed=1, dup_frac=0.1):
    """C1-shaped network WITH planted structure, so that some gene-GO pairs are genuinely
    significant: genes are split into modules; a fraction p_in of gene-gene edges stay inside a
    module; every module owns a few GO terms that annotate its genes (fraction p_in of the
    annotation edges), the rest of the annotations are Zipf noise.  Returns (graph, genes)."""
    rng = np.random.default_rng(seed)
    genes = ['G%05d' % i for i in range(n_genes)]
    gos = ['GO:9%06d' % i for i in range(n_go)]
    mg = nx.MultiGraph()
    mg.add_nodes_from(genes)
    for i, g in enumerate(gos):
        mg.add_node(g, GO=g, name='term %d' % i, domain=DOMAINS[i % 3])
    module = rng.integers(0, n_modules, size=n_genes)
    members = [np.nonzero(module == k)[0] for k in range(n_modules)]
    terms_of = [rng.choice(n_go, size=4, replace=False) for _ in range(n_modules)]
    a = rng.integers(0, n_genes, size=n_gene_edges)
    inside = rng.random(n_gene_edges) < p_in
    b = np.where(inside, [rng.choice(members[module[x]]) for x in a],
                 rng.integers(0, n_genes, size=n_gene_edges))
    keep = a != b
    a, b = a[keep], b[keep]
    mg.add_edges_from(((genes[x], genes[y], {'label': 'interacts'}) for x, y in zip(a, b)))
    ndup = int(len(a) * dup_frac)
    mg.add_edges_from(((genes[x], genes[y], {'label': 'regulates'})
                       for x, y in zip(a[:ndup], b[:ndup])))
    ga = rng.integers(0, n_genes, size=n_annot)
    tw = (np.arange(1, n_go + 1, dtype=np.float64)) ** -1.2
    tw /= tw.sum()
    own = rng.random(n_annot) < p_in
    gt = np.where(own, [rng.choice(terms_of[module[x]]) for x in ga],
                  rng.choice(n_go, size=n_annot, p=tw))
    mg.add_edges_from(((genes[x], gos[y], {'label': 'GO:annotation'}) for x, y in zip(ga, gt)))
    child = rng.integers(1, n_go, size=n_isa)
    parent = (rng.random(n_isa) * child).astype(np.int64)
    mg.add_edges_from(((gos[c], gos[p], {'label': 'GO:is_a'}) for c, p in zip(child, parent)))
    return mg, [{'ID': g} for g in genes]


def human_scale_edges(seed=2, n_genes=20000, n_go=18000, n_gene_edges=700000, n_annot=250000,
                      n_isa=50000, n_modules=500, p_in=0.5):
    """C2/C3 (BASELINE.json configs[1..2]): ~20k genes + ~18k GO terms, ~1M edges.
    Gene-gene edges follow a Chung-Lu model with power-law weights (exponent 2.5) mixed with
    planted modules; gene-GO annotations are half module-specific, half Zipf; GO-GO edges form a
    DAG (child -> parent with smaller index).  Returns (n, u, v, is_go_node) as numpy arrays with
    genes first (ids 0..n_genes-1) and GO terms after."""
    rng = np.random.default_rng(seed)
    w = (np.arange(1, n_genes + 1, dtype=np.float64) + 10.0) ** (-1.0 / 1.5)
    w /= w.sum()
    a = rng.choice(n_genes, size=n_gene_edges, p=w)
    module = rng.integers(0, n_modules, size=n_genes)
    order = np.argsort(module, kind='stable')
    start = np.searchsorted(module[order], np.arange(n_modules))
    size = np.bincount(module, minlength=n_modules)
    inside = rng.random(n_gene_edges) < p_in
    pick = start[module[a]] + (rng.random(n_gene_edges) * size[module[a]]).astype(np.int64)
    b = np.where(inside, order[pick], rng.choice(n_genes, size=n_gene_edges, p=w))
    keep = a != b
    a, b = a[keep], b[keep]
    terms_of = rng.integers(0, n_go, size=(n_modules, 6))
    ga = rng.choice(n_genes, size=n_annot, p=w)
    tw = (np.arange(1, n_go + 1, dtype=np.float64)) ** -1.0
    tw /= tw.sum()
    own = rng.random(n_annot) < 0.5
    gt = np.where(own, terms_of[module[ga], rng.integers(0, 6, size=n_annot)],
                  rng.choice(n_go, size=n_annot, p=tw))
    child = rng.integers(1, n_go, size=n_isa)
    parent = (rng.random(n_isa) * child).astype(np.int64)
    u = np.concatenate([a, ga, n_genes + child]).astype(np.int32)
    v = np.concatenate([b, n_genes + gt, n_genes + parent]).astype(np.int32)
    n = n_genes + n_go
    is_go = np.arange(n) >= n_genes
    return n, u, v, is_go


def edges_to_networkx(n, u, v, is_go, n_genes):
    """The nx.MultiGraph a GeneWalk assembler would hand to the hot path (string ids, GO attrs)."""
    names = ['G%05d' % i for i in range(n_genes)] + ['GO:9%06d' % i for i in range(n - n_genes)]
    mg = nx.MultiGraph()
    mg.add_nodes_from(names[:n_genes])
    for i in range(n_genes, n):
        mg.add_node(names[i], GO=names[i], name='term %d' % (i - n_genes),
                    domain=DOMAINS[i % 3])
    mg.add_edges_from(zip([names[x] for x in u.tolist()], [names[x] for x in v.tolist()]))
    return mg, [{'ID': g} for g in names[:n_genes]]


def csr_from_edges_numpy(n, u, v):
    """Distinct-neighbour CSR (rows sorted by neighbour id) of an undirected multi-edge list."""
    key = np.unique(np.concatenate([u.astype(np.int64) * n + v, v.astype(np.int64) * n + u]))
    src = (key // n).astype(np.int32)
    dst = (key % n).astype(np.int32)
    row_ptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(row_ptr, src + 1, 1)
    return np.cumsum(row_ptr).astype(np.int32), dst


def indra_scale_edges(seed=3, n=50000, m=5000000):
    """C4 (BASELINE.json configs[3]): INDRA-scale dense multigraph -- 50k nodes, 5M multi-edges drawn
    from a Chung-Lu model with Pareto(1.8) weights, so many pairs repeat (they collapse to ~10M
    adjacency entries) and the largest degrees reach ~1e4.  Returns (n, u, v)."""
    rng = np.random.default_rng(seed)
    w = rng.pareto(1.8, size=n) + 1.0
    w /= w.sum()
    u = rng.choice(n, size=m, p=w).astype(np.int32)
    v = rng.choice(n, size=m, p=w).astype(np.int32)
    keep = u != v
    return n, u[keep], v[keep]


def power_law_stress_edges(seed=4, n=1000000, m=5000000):
    """C5 (BASELINE.json configs[4]): 1M-node power-law graph (Barabasi-Albert-like, m = 5:
    ~10M adjacency entries), no GO attributes.  Returns (n, u, v)."""
    rng = np.random.default_rng(seed)
    w = (np.arange(1, n + 1, dtype=np.float64)) ** -0.5
    w /= w.sum()
    u = rng.choice(n, size=m, p=w).astype(np.int32)
    v = rng.choice(n, size=m, p=w).astype(np.int32)
    keep = u != v
    return n, u[keep], v[keep]


def multigraph_degrees_from_edges(n, u, v):
    """``mg.degree`` of the multigraph with edges (u[k], v[k]): parallel edges counted, a
    self-loop counts 2."""
    return (np.bincount(u, minlength=n) + np.bincount(v, minlength=n)).astype(np.int64)
