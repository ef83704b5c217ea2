"""Gene-GO similarity statistics on the GPU behind genewalk's ``GeneWalk`` class.

Mirrors /root/reference/genewalk/perform_statistics.py: ``GeneWalk(graph, genes, nvs, null_dist,
gene_id_type)`` with ``generate_output(alpha_fdr)``, ``psim(sim)``, ``log_stats(vals)``,
``global_fdr``.  All arithmetic -- cosine similarities (:102), empirical p-values (:247-258),
per-gene and global Benjamini-Hochberg (:112,:273), geometric mean / CI (:117-124), mean / SEM of
the similarities (:180-184) -- runs batched in libgenewalk_b200.so; what stays in Python is the row
bookkeeping (which pairs exist, labels, DataFrame assembly and sorting), kept line-for-line
compatible with the reference's output columns, order and dtypes.
"""
import logging

import networkx as nx
import numpy as np
import pandas as pd

from . import _lib

logger = logging.getLogger('genewalk.perform_statistics')


def _dev(arr, dtype):
    import torch
    return torch.from_numpy(np.ascontiguousarray(arr, dtype=dtype)).cuda()


class GeneWalk(object):
    """GeneWalk object that generates the final output list of significant GO
    terms for each gene in the input list with genes of interest from an
    experiment, for example differentially expressed genes or CRISPR screen
    hits.

    Parameters
    ----------
    graph : networkx.MultiGraph
        GeneWalk network for which the statistics are calculated.
    genes : list of dict
        List of gene references for relevant genes.
    nvs : list of node-vector containers
        Node vectors for nodes in the graph (``NodeVectors`` or any object with ``vectors``,
        ``index_to_key`` / ``key_to_index``), one per real-network replicate.
    null_dist : np.array
        Similarity random (null) distribution, ascending.
    gene_id_type : Optional[str]
        The type of gene IDs that were the basis of doing the analysis. Default: hgnc_symbol
    """

    def __init__(self, graph, genes, nvs, null_dist, gene_id_type='hgnc_symbol'):
        self.graph = graph
        self.genes = genes
        self.nvs = nvs
        self.srd = null_dist
        self.gene_id_type = gene_id_type
        self.go_nodes = set(nx.get_node_attributes(self.graph, 'GO'))
        self.gene_nodes = self.get_gene_nodes()
        self._d_srd = None

    # -- identical helpers ------------------------------------------------------------------------
    def get_gene_nodes(self):
        if self.gene_id_type == 'custom':
            return {g['ID'] for g in self.genes}
        else:
            return {g['HGNC_SYMBOL'] for g in self.genes}

    def get_gene_node_id(self, gene):
        return gene['HGNC_SYMBOL'] if self.gene_id_type != 'custom' \
            else gene['ID']

    def get_gene_attribs(self, gene):
        """Return an attribute dict for a given gene."""
        gene_id = self.get_gene_node_id(gene)
        if gene_id in self.graph:
            ncon_gene = len(self.graph[gene_id])
        else:
            ncon_gene = np.nan
        hgnc_symbol = gene['HGNC_SYMBOL'] if self.gene_id_type != 'custom' \
            else None
        hgnc_id = gene['HGNC'] if self.gene_id_type != 'custom' else None
        custom_id = gene['ID'] if self.gene_id_type == 'custom' else None
        return {
            'hgnc_symbol': hgnc_symbol,
            'hgnc_id': hgnc_id,
            'custom_id': custom_id,
            'ncon_gene': ncon_gene
        }

    def _prefix(self, gene):
        if self.gene_id_type == 'mgi_id':
            return [gene.get('MGI', '')]
        elif self.gene_id_type == 'rgd_id':
            return [gene.get('RGD', '')]
        elif self.gene_id_type == 'ensembl_id':
            return [gene.get('ENSEMBL', '')]
        elif self.gene_id_type in {'entrez_mouse', 'entrez_human'}:
            return [gene.get('EGID', '')]
        elif self.gene_id_type == 'custom':
            return [gene.get('ID', '')]
        return []

    def add_empty_row(self, gene, gene_attribs):
        row = [gene_attribs['hgnc_symbol'],
               gene_attribs['hgnc_id'],
               '', '', '',
               gene_attribs['ncon_gene'],
               np.nan, np.nan, np.nan, np.nan,
               np.nan, np.nan, np.nan, np.nan, np.nan]
        row.extend([np.nan for i in range(getattr(self, '_nreps', len(self.nvs)))])
        return self._prefix(gene) + row

    # -- device kernels ---------------------------------------------------------------------------
    def _device_null(self):
        if self._d_srd is None:
            _lib.require_cuda()
            self._d_srd = _dev(np.asarray(self.srd), np.float32)
        return self._d_srd

    def psim(self, sim):
        """Determine the p-value of the experimental similarity by determining
        its percentile, i.e. the normalized rank, in the null distribution
        with random similarity values (gw_psim)."""
        return float(self._psim_batch(np.asarray([sim], dtype=np.float32))[0])

    def _psim_batch(self, sims):
        import torch
        srd = self._device_null()
        d_s = sims if hasattr(sims, 'data_ptr') else _dev(sims, np.float32)
        out = torch.empty(d_s.numel(), dtype=torch.float64, device=srd.device)
        _lib.call('gw_psim', _lib.ptr(srd), srd.numel(), _lib.ptr(d_s), d_s.numel(), None,
                  _lib.ptr(out), _lib.stream_ptr())
        return out if hasattr(sims, 'data_ptr') else out.cpu().numpy()

    @staticmethod
    def _bh(pvals, seg_ptr):
        """pvals: torch float64 [m] on device; seg_ptr: torch int64 [nseg+1] on device."""
        import torch
        q = torch.empty_like(pvals)
        _lib.call('gw_bh_segmented', _lib.ptr(pvals), pvals.numel(), _lib.ptr(seg_ptr),
                  seg_ptr.numel() - 1, _lib.ptr(q), _lib.stream_ptr())
        return q

    @staticmethod
    def _log_stats_rows(vals):
        """vals: torch float64 [rows, nreps] on device -> [rows, 3] (g_mean, low, upp)."""
        import torch
        vals = vals.contiguous()
        out = torch.empty((vals.shape[0], 3), dtype=torch.float64, device=vals.device)
        _lib.call('gw_log_stats', _lib.ptr(vals), vals.shape[0], vals.shape[1], _lib.ptr(out),
                  _lib.stream_ptr())
        return out

    def log_stats(self, vals):
        import torch
        _lib.require_cuda()
        v = torch.as_tensor(np.asarray(vals, dtype=np.float64)).cuda().reshape(1, -1)
        g, lo, up = self._log_stats_rows(v)[0].cpu().numpy()
        return g, lo, up

    def _row_index_of_nodes(self, nv, node_index):
        """row_of_node[node id] = row of that node in nv.vectors (-1 when absent)."""
        n = len(node_index)
        row_of_node = np.full(n, -1, dtype=np.int64)
        ids = getattr(nv, 'node_ids', None)
        keys = nv.index_to_key
        if ids is not None and len(ids) == len(keys) and len(keys) > 0 and \
                int(np.max(ids)) < n and node_index.get(keys[0]) == int(ids[0]) and \
                node_index.get(keys[-1]) == int(ids[-1]):
            row_of_node[np.asarray(ids, dtype=np.int64)] = np.arange(len(keys))
        else:
            for r, k in enumerate(keys):
                i = node_index.get(k)
                if i is not None:
                    row_of_node[i] = r
        return row_of_node

    # -- main -------------------------------------------------------------------------------------
    def collect_pairs(self):
        """Bookkeeping of perform_statistics.py:75-97: which (gene, GO) pairs exist.

        Returns a dict with the gene attribute dicts, the node-id arrays ``pair_gene`` /
        ``pair_go`` (grouped by gene, GO terms in adjacency order), ``seg_ptr`` (one segment per
        gene that has GO connections) and ``gene_of_seg`` (index into ``self.genes``)."""
        graph = self.graph
        node_index = {v: i for i, v in enumerate(graph.nodes())}
        go_nodes = self.go_nodes
        gene_attr_list = []
        pair_gene, pair_go, seg_ptr = [], [], [0]
        gene_of_seg = []
        for gi, gene in enumerate(self.genes):
            ga = self.get_gene_attribs(gene)
            gene_attr_list.append(ga)
            if isinstance(ga['ncon_gene'], float) and np.isnan(ga['ncon_gene']):
                continue
            gene_node_id = ga['hgnc_symbol'] if self.gene_id_type != 'custom' else ga['custom_id']
            connected = [w for w in graph[gene_node_id] if w in go_nodes]
            if connected:
                g_idx = node_index[gene_node_id]
                pair_gene.extend([g_idx] * len(connected))
                pair_go.extend(node_index[w] for w in connected)
                seg_ptr.append(len(pair_go))
                gene_of_seg.append(gi)
        return {'node_index': node_index, 'gene_attribs': gene_attr_list,
                'pair_gene': np.asarray(pair_gene, dtype=np.int64),
                'pair_go': np.asarray(pair_go, dtype=np.int64),
                'seg_ptr': np.asarray(seg_ptr, dtype=np.int64), 'gene_of_seg': gene_of_seg}

    def pair_similarities(self, nv, pairs, out=None):
        """Cosine similarity of every (gene, GO) pair for one replicate's vectors (device
        float32 [m]); gw_cosine_pairs replaces nv.similarity at perform_statistics.py:102."""
        import torch
        m = len(pairs['pair_go'])
        if out is None:
            out = torch.empty(m, dtype=torch.float32, device='cuda')
        if m == 0:
            return out
        row_of_node = self._row_index_of_nodes(nv, pairs['node_index'])
        a = row_of_node[pairs['pair_gene']]
        b = row_of_node[pairs['pair_go']]
        if (a < 0).any() or (b < 0).any():
            bad = pairs['pair_gene'][a < 0] if (a < 0).any() else pairs['pair_go'][b < 0]
            raise KeyError("Key '%s' not present" % (list(self.graph.nodes())[int(bad[0])],))
        vec = _dev(nv.vectors, np.float32)
        d_a, d_b = _dev(a, np.int32), _dev(b, np.int32)
        _lib.call('gw_cosine_pairs', int(vec.shape[1]), _lib.ptr(vec), m, _lib.ptr(d_a),
                  _lib.ptr(d_b), _lib.ptr(out), _lib.stream_ptr())
        return out

    def generate_output(self, alpha_fdr=1, sims=None, pairs=None):
        """Main function of GeneWalk object that generates the final
        GeneWalk output table (in csv format).

        Parameters
        ----------
        alpha_fdr : Optional[float]
            Significance level for FDR [0,1] (default=1, i.e. all GO
            terms and their statistics are output). If set to a lower value,
            only connected GO terms with mean padj < alpha_FDR are output.
        sims : Optional[torch.Tensor]
            Precomputed device float32 [nreps, m] gene-GO similarities in ``collect_pairs`` order
            (what the multi-GPU driver all-gathers); default: computed here from ``self.nvs``.
        """
        import torch
        _lib.require_cuda()
        graph = self.graph
        if pairs is None:
            pairs = self.collect_pairs()
        gene_attr_list = pairs['gene_attribs']
        pair_go, seg_ptr, gene_of_seg = pairs['pair_go'], pairs['seg_ptr'], pairs['gene_of_seg']
        m = len(pair_go)
        nreps = len(self.nvs) if sims is None else int(sims.shape[0])
        self._nreps = nreps

        # ---- numeric core on the device ---------------------------------------------------------
        if m > 0:
            dev = torch.device('cuda')
            st = _lib.stream_ptr()
            if sims is None:
                sims = torch.empty((nreps, m), dtype=torch.float32, device=dev)
                for r, nv in enumerate(self.nvs):
                    self.pair_similarities(nv, pairs, out=sims[r])
            sims = sims.to(dev).contiguous()
            pvals = self._psim_batch(sims.reshape(-1)).reshape(nreps, m)
            # per-gene BH for every replicate in one call: segments = (replicate, gene)
            seg_all = np.concatenate([seg_ptr[:-1] + r * m for r in range(nreps)] + [[nreps * m]])
            qvals = self._bh(pvals.reshape(-1), _dev(seg_all, np.int64)).reshape(nreps, m)
            q_stats = self._log_stats_rows(qvals.t())       # [m, 3]
            p_stats = self._log_stats_rows(pvals.t())
            mean_sim = torch.empty(m, dtype=torch.float32, device=dev)
            sem_sim = torch.empty(m, dtype=torch.float64, device=dev)
            sims_t = sims.t().contiguous()
            _lib.call('gw_mean_sem_f32', _lib.ptr(sims_t), m, nreps, _lib.ptr(mean_sim),
                      _lib.ptr(sem_sim), st)
            q_stats = q_stats.cpu().numpy()
            p_stats = p_stats.cpu().numpy()
            mean_sim = mean_sim.cpu().numpy()
            sem_sim = sem_sim.cpu().numpy()
            pvals_h = pvals.t().cpu().numpy()               # [m, nreps]
            self.sims = sims.cpu().numpy()                  # kept for inspection / tests
        # ---- rows (perform_statistics.py:158-218) -----------------------------------------------
        nodes = list(graph.nodes())
        node_data = graph.nodes
        adj = graph.adj
        rows = []
        seg_of_gene = {gi: s for s, gi in enumerate(gene_of_seg)}
        for gi, gene in enumerate(self.genes):
            ga = gene_attr_list[gi]
            in_graph = not (isinstance(ga['ncon_gene'], float) and np.isnan(ga['ncon_gene']))
            if in_graph and gi in seg_of_gene:
                s = seg_of_gene[gi]
                prefix = self._prefix(gene)
                for k in range(int(seg_ptr[s]), int(seg_ptr[s + 1])):
                    gene_padj = q_stats[k, 0]
                    if gene_padj < alpha_fdr or alpha_fdr == 1:
                        go_id = nodes[pair_go[k]]
                        row = [ga['hgnc_symbol'], ga['hgnc_id'],
                               node_data[go_id]['name'], go_id,
                               node_data[go_id]['domain'].replace('_', ' '),
                               ga['ncon_gene'], len(adj[go_id]),
                               gene_padj, p_stats[k, 0],
                               mean_sim[k], sem_sim[k],
                               q_stats[k, 1], q_stats[k, 2],
                               p_stats[k, 1], p_stats[k, 2]]
                        row.extend(pvals_h[k].tolist())
                        rows.append(prefix + row)
            elif alpha_fdr == 1:  # case: no GO connections / not in graph
                rows.append(self.add_empty_row(gene, ga))
        header = ['hgnc_symbol', 'hgnc_id',
                  'go_name', 'go_id', 'go_domain',
                  'ncon_gene', 'ncon_go',
                  'gene_padj', 'pval',
                  'sim', 'sem_sim',
                  'cilow_gene_padj', 'ciupp_gene_padj',
                  'cilow_pval', 'ciupp_pval']
        header.extend(['pval_rep' + str(i) for i in range(nreps)])
        if self.gene_id_type in {'mgi_id', 'rgd_id', 'ensembl_id',
                                 'entrez_human', 'entrez_mouse', 'custom'}:
            header = [self.gene_id_type] + header

        df = pd.DataFrame.from_records(rows, columns=header)
        df = self.global_fdr(df, alpha_fdr)
        df.drop(['pval_rep' + str(i) for i in range(nreps)], axis=1, inplace=True)
        df[self.gene_id_type] = pd.Categorical(
            df[self.gene_id_type], categories=df[self.gene_id_type].unique(), ordered=True)
        df[['ncon_gene', 'ncon_go']] = df[['ncon_gene', 'ncon_go']].astype('str')
        df = df.sort_values(by=[self.gene_id_type, 'global_padj', 'gene_padj',
                                'sim', 'go_domain', 'go_name'],
                            ascending=[True, True, True, False, True, True])
        if self.gene_id_type == 'custom':
            df.drop(['hgnc_symbol', 'hgnc_id'], axis=1, inplace=True)
        return df

    def global_fdr(self, df, alpha_fdr):
        """Determine the global_padj values through FDR multiple testing
        correction over all gene - GO annotation pairs present in the output
        file (one BH segment per replicate, gw_bh_segmented; then gw_log_stats)."""
        import torch
        nreps = getattr(self, '_nreps', len(self.nvs))
        global_stats = {'global_padj': [], 'cilow_global_padj': [],
                        'ciupp_global_padj': []}
        colloc = {'global_padj': 8, 'cilow_global_padj': 4,
                  'ciupp_global_padj': 4}
        ids = df[~df['pval_rep0'].isna()].index
        rows = len(ids)
        if rows > 0:
            P = np.stack([df['pval_rep' + str(i)][ids].to_numpy(dtype=np.float64)
                          for i in range(nreps)])                       # [nreps, rows]
            d_p = _dev(P.reshape(-1), np.float64)
            seg = _dev(np.arange(nreps + 1, dtype=np.int64) * rows, np.int64)
            q = self._bh(d_p, seg).reshape(nreps, rows)
            stats = self._log_stats_rows(q.t()).cpu().numpy()
            global_stats['global_padj'] = stats[:, 0]
            global_stats['cilow_global_padj'] = stats[:, 1]
            global_stats['ciupp_global_padj'] = stats[:, 2]
        for key in global_stats.keys():
            df.insert((len(df.columns) - colloc[key] - nreps), key, np.nan)
            if rows > 0:
                df.loc[ids, key] = global_stats[key]
        return df
