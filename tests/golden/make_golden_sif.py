"""Golden fixture for network ingestion (tests/golden/sif_networks.json): the UNMODIFIED reference's
UserNxMgAssembler (/root/reference/genewalk/nx_mg_assembler.py:342-422, goatools stubbed by
oracle/ref_shim.py) run on the reference's own test networks; for every case the node order, every
node's neighbour list (``list(graph[v])``), multigraph degrees, edge count and GO node attributes.

    python tests/golden/make_golden_sif.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402

RES = '/root/reference/genewalk/tests/resources'


class Resources(object):
    def get_go_obo(self):
        return os.path.join(RES, 'go.obo')


def main():
    asm = ref_shim.import_assembler()
    hgnc = [{'HGNC_SYMBOL': s, 'HGNC': str(i), 'UP': ''} for i, s in
            enumerate(['KRAS', 'BRAF', 'MAP2K2', 'MAPK1', 'PIK3CA', 'AKT1'])]
    custom = [{'ID': s} for s in ('CUSTOM:ABC', 'CUSTOM:XYZ', 'CUSTOM:ZZZ')]
    cases = []
    for name, fname, genes in (('sif_full_hgnc', 'test_sif_full.sif', hgnc),
                               ('sif_full_hgnc_subset', 'test_sif_full.sif', hgnc[:3]),
                               ('sif_custom_filter', 'test_sif_custom.sif', custom)):
        path = os.path.join(RES, fname)
        mga = asm.UserNxMgAssembler(genes, Resources(), path, gwn_format='sif_full')
        g = mga.graph
        with open(path) as f:
            rows = [line.strip().split(',') for line in f if line.strip()]
        cases.append({'name': name, 'rows': rows, 'genes': genes, 'nodes': list(g.nodes()),
                      'adjacency': {v: list(g[v]) for v in g.nodes()},
                      'degrees': [g.degree(v) for v in g.nodes()],
                      'n_edges': g.number_of_edges(),
                      'attrs': {v: dict(g.nodes[v]) for v in g.nodes() if g.nodes[v]}})
    with open(os.path.join(HERE, 'sif_networks.json'), 'w') as f:
        # the GO terms the stub GODag saw, re-serialised as a minimal OBO (id / name / namespace)
        dag = ref_shim.ShimGODag(os.path.join(RES, 'go.obo'))
        obo = ''.join('[Term]\nid: %s\nname: %s\nnamespace: %s\n\n' % (t.id, t.name, t.namespace)
                      for k, t in sorted(dag.items()) if k == t.id)
        obo += '[Term]\nid: GO:9999999\nname: gone\nnamespace: biological_process\nis_obsolete: true\n'
        json.dump({'obo': obo, 'cases': cases}, f, indent=1)
    for c in cases:
        print(c['name'], len(c['nodes']), 'nodes', c['n_edges'], 'edges', len(c['attrs']), 'annotated')


if __name__ == '__main__':
    main()
