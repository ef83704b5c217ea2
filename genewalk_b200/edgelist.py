"""GeneWalk network files -> CSR without networkx.

Replaces the graph construction of ``UserNxMgAssembler.add_network_edges``
(/root/reference/genewalk/nx_mg_assembler.py:368-403): ``pd.read_csv(dtype=str, header=None)`` of a
SIF (``nodeA,rel_type,nodeB``) or edge list (``nodeA,nodeB``), ``nx.from_pandas_edgelist(...,
create_using=nx.MultiGraph)``, then removal of every node that is neither a GO term (id starts with
``GO:``) nor in the input gene list.  At human scale building and pickling the ``nx.MultiGraph``
is the serial bottleneck in front of the GPU stages (18 s for ~1 M edges); here the same graph is
produced as arrays with a handful of vectorised numpy passes:

  * node ids in networkx's insertion order (first appearance scanning source, target row by row),
  * distinct-neighbour adjacency in networkx's order (for node v: neighbours by the first row in
    which the edge {v, w} appears), so that ``col_idx[row_ptr[v] + k] == list(graph[v])[k]`` of
    the reference's graph, which is what the walks sample from (deepwalk.py:165,173,197),
  * multigraph degrees (parallel edges counted, a self-loop counts 2) for the randomizer
    (null_distributions.py:23).

``EdgeListNetwork`` exposes the few graph operations the statistics step uses (``v in g``,
``g[v]``, ``g.nodes()``, ``g.nodes[v]``), so it can be passed wherever the assembled
``nx.MultiGraph`` goes: ``DeepWalk``, ``get_rand_graph``, ``GeneWalk`` and ``run_genewalk``.
GO node metadata (name, domain) comes from ``read_obo_terms`` (the reference takes it from
goatools' GODag, nx_mg_assembler.py:414-422).  Adding GO annotations / the ontology to gene-only
files ('sif', 'sif_annot', 'edge_list': nx_mg_assembler.py:404-411) needs the GAF and OBO loaders of
the input-assembly layer, which is out of scope: those formats are read as they are.
"""
import numpy as np

from .csr import GraphCSR

FORMATS = {'edge_list': (0, 1), 'sif': (0, 2), 'sif_annot': (0, 2), 'sif_full': (0, 2)}


class _NodeView(object):
    """``graph.nodes`` of networkx for an EdgeListNetwork: iterable, callable, ``[label]`` -> attrs."""

    def __init__(self, net):
        self._net = net

    def __call__(self):
        return list(self._net.csr.nodes)

    def __iter__(self):
        return iter(self._net.csr.nodes)

    def __len__(self):
        return self._net.csr.n

    def __contains__(self, label):
        return label in self._net.csr.index

    def __getitem__(self, label):
        if label not in self._net.csr.index:
            raise KeyError(label)
        return self._net.node_attrs.setdefault(label, {})


class EdgeListNetwork(object):
    """A GeneWalk network held as CSR arrays.

    Attributes
    ----------
    csr : GraphCSR        nodes in networkx insertion order, neighbours in networkx adjacency order
    degrees : np.ndarray  multigraph degree of every node (``mg.degree``)
    is_go : np.ndarray    bool, node id starts with 'GO:'
    node_attrs : dict     label -> {'GO', 'name', 'domain'} for GO nodes (``annotate_go``)
    n_edges : int         number of (multi-)edges kept
    """

    def __init__(self, csr, degrees, is_go, n_edges):
        self.csr = csr
        self.degrees = degrees
        self.is_go = is_go
        self.n_edges = int(n_edges)
        self.node_attrs = {}
        self.nodes = _NodeView(self)
        self.graph = {'_gw_csr': csr}       # what deepwalk._csr_of looks for on an nx graph

    # -- the graph operations the hot path's callers use ---------------------------------------
    def __contains__(self, label):
        return label in self.csr.index

    def __len__(self):
        return self.csr.n

    def __iter__(self):
        return iter(self.csr.nodes)

    def __getitem__(self, label):
        v = self.csr.index[label]
        nodes = self.csr.nodes
        return [nodes[w] for w in self.csr.col_idx[self.csr.row_ptr[v]:self.csr.row_ptr[v + 1]]]

    def number_of_nodes(self):
        return self.csr.n

    def number_of_edges(self):
        return self.n_edges

    def is_multigraph(self):
        return True

    def go_node_labels(self):
        """Labels carrying a 'GO' attribute (``nx.get_node_attributes(graph, 'GO')``)."""
        return {k for k, a in self.node_attrs.items() if 'GO' in a}

    def annotate_go(self, terms):
        """Attach GO / name / domain to the GO nodes found in ``terms`` (id -> (id, name,
        namespace), e.g. from ``read_obo_terms``): nx_mg_assembler.py:414-422."""
        for v in np.nonzero(self.is_go)[0]:
            label = self.csr.nodes[v]
            t = terms.get(label)
            if t:
                self.node_attrs[label] = {'GO': t[0], 'name': t[1], 'domain': t[2]}
        return self

    def to_networkx(self):
        """The equal ``nx.MultiGraph`` (same node and adjacency order; edge attributes are not
        kept), for callers that insist on networkx."""
        import networkx as nx
        mg = nx.MultiGraph()
        for label in self.csr.nodes:
            mg.add_node(label, **self.node_attrs.get(label, {}))
        mg.add_edges_from(zip(self._edge_labels[0], self._edge_labels[1]))
        return mg


def read_obo_terms(path):
    """id -> (id, name, namespace) of the non-obsolete [Term] stanzas of a GO OBO file; alt_ids map
    to their primary term (what goatools' GODag exposes as ``dag[id].id/.name/.namespace``)."""
    terms, cur, in_term = {}, None, False

    def flush():
        if cur and cur.get('id') and not cur.get('obsolete'):
            t = (cur['id'], cur.get('name', ''), cur.get('namespace', ''))
            terms[cur['id']] = t
            for alt in cur.get('alt', ()):
                terms.setdefault(alt, t)
    with open(path) as f:
        for line in f:
            line = line.strip()
            if line.startswith('['):
                flush()
                in_term = line == '[Term]'
                cur = {} if in_term else None
            elif in_term and ': ' in line:
                key, val = line.split(': ', 1)
                if key in ('id', 'name', 'namespace'):
                    cur.setdefault(key, val)
                elif key == 'alt_id':
                    cur.setdefault('alt', []).append(val)
                elif key == 'is_obsolete' and val.startswith('true'):
                    cur['obsolete'] = True
    flush()
    return terms


def network_from_edges(source, target, genes=None):
    """EdgeListNetwork from parallel arrays of node labels.  ``genes`` (list of gene dicts with
    'ID' or 'HGNC_SYMBOL', as produced by gene_lists.read_gene_list) enables the reference's
    filter: nodes that are neither GO terms nor listed genes are removed with their edges; None
    keeps every node."""
    import pandas as pd
    source = np.asarray(source, dtype=object)
    target = np.asarray(target, dtype=object)
    m = len(source)
    # networkx inserts nodes as add_edge(u, v) meets them: u then v, row by row
    codes, labels = pd.factorize(np.stack([source, target], axis=1).ravel(), sort=False)
    labels = np.asarray(labels, dtype=object)
    u, v = codes[0::2].astype(np.int64), codes[1::2].astype(np.int64)
    is_go = np.fromiter((isinstance(x, str) and x.startswith('GO:') for x in labels), dtype=bool,
                        count=len(labels))
    keep = np.ones(len(labels), dtype=bool)
    if genes is not None:
        listed = {(g['ID'] if 'ID' in g else g['HGNC_SYMBOL']) for g in genes}
        keep = is_go | np.fromiter((x in listed for x in labels), dtype=bool, count=len(labels))
    # remove_nodes_from keeps the order of the surviving nodes and of their surviving neighbours
    new_id = np.cumsum(keep) - 1
    e_keep = keep[u] & keep[v]
    u, v = new_id[u[e_keep]], new_id[v[e_keep]]
    labels, is_go = labels[keep], is_go[keep]
    n = len(labels)
    row = np.nonzero(e_keep)[0]
    # directed entries (a -> b) of every edge in both directions, stamped with the edge's row
    a = np.concatenate([u, v])
    b = np.concatenate([v, u])
    t = np.concatenate([row, row])
    degrees = np.bincount(a, minlength=n).astype(np.int64)        # a self-loop contributes twice
    # first row of every distinct (a, b), then neighbours of a in the order of those first rows
    key = a * n + b
    first = np.lexsort((t, key))
    key_s = key[first]
    head = np.ones(len(key_s), dtype=bool)
    head[1:] = key_s[1:] != key_s[:-1]
    a1, b1, t1 = a[first][head], b[first][head], t[first][head]
    order = np.lexsort((t1, a1))
    a1, b1 = a1[order], b1[order]
    row_ptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(a1, minlength=n), out=row_ptr[1:])
    if row_ptr[-1] >= 2 ** 31:
        raise ValueError('graph has too many adjacency entries for int32 CSR')
    nodes = labels.tolist()
    net = EdgeListNetwork(GraphCSR(nodes, row_ptr.astype(np.int32), b1.astype(np.int32)), degrees,
                          is_go, len(u))
    net._edge_labels = (labels[u], labels[v])
    return net


def read_network(filepath, gwn_format='sif_full', genes=None, obo=None):
    """Load a user-provided GeneWalk network file (nx_mg_assembler.py:368-403).

    ``gwn_format``: 'edge_list' (nodeA,nodeB[,attributes]) or 'sif' / 'sif_annot' / 'sif_full'
    (nodeA,rel_type,nodeB); no header.  ``genes``: the input gene list (see
    ``network_from_edges``).  ``obo``: path of a GO OBO file for the GO node metadata."""
    import pandas as pd
    if gwn_format not in FORMATS:
        raise ValueError('%s is not a valid GeneWalk network format' % gwn_format)
    df = pd.read_csv(filepath, dtype=str, header=None)
    cs, ct = FORMATS[gwn_format]
    net = network_from_edges(df[cs].to_numpy(dtype=object), df[ct].to_numpy(dtype=object), genes)
    if obo is not None:
        net.annotate_go(read_obo_terms(obo))
    return net
