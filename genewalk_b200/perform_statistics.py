"""Gene-GO similarity statistics on the GPU behind genewalk's ``GeneWalk`` class.

Mirrors /root/reference/genewalk/perform_statistics.py: ``GeneWalk(graph, genes, nvs, null_dist,
gene_id_type)`` with ``generate_output(alpha_fdr)``, ``psim(sim)``, ``log_stats(vals)``,
``global_fdr``.  All arithmetic -- cosine similarities (:102), empirical p-values (:247-258),
per-gene and global Benjamini-Hochberg (:112,:273), geometric mean / CI (:117-124), mean / SEM of
the similarities (:180-184), the alpha_fdr row filter (:187) -- runs batched in
libgenewalk_b200.so; what stays on the host is the bookkeeping of which pairs exist and the
column-wise assembly of the output table (same columns, order, dtypes and sort as the reference).
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
        if hasattr(self.graph, 'go_node_labels'):     # edgelist.EdgeListNetwork
            self.go_nodes = self.graph.go_node_labels()
        else:
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
        if hasattr(graph, 'csr'):
            return self._collect_pairs_csr()
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

    def _collect_pairs_csr(self):
        """collect_pairs for a network held as CSR arrays (edgelist.EdgeListNetwork): the same
        pairs in the same order, found with array operations instead of per-gene dict lookups."""
        csr = self.graph.csr
        csr.to_host()
        node_index = csr.index
        go_mask = np.zeros(csr.n, dtype=bool)
        go_mask[[node_index[k] for k in self.go_nodes]] = True
        deg = np.diff(csr.row_ptr).astype(np.int64)
        gid = np.array([node_index.get(self.get_gene_node_id(g), -1) for g in self.genes], dtype=np.int64)
        present = gid >= 0
        ncon = np.where(present, deg[np.maximum(gid, 0)], 0)
        gene_attr_list = []
        custom = self.gene_id_type == 'custom'
        for g, ok, c in zip(self.genes, present, ncon):
            gene_attr_list.append({'hgnc_symbol': None if custom else g['HGNC_SYMBOL'],
                                   'hgnc_id': None if custom else g['HGNC'],
                                   'custom_id': g['ID'] if custom else None,
                                   'ncon_gene': int(c) if ok else np.nan})
        # adjacency entries of the listed genes, gene by gene in list order
        gsel = np.nonzero(present)[0]
        starts = csr.row_ptr[gid[gsel]].astype(np.int64)
        lens = ncon[gsel]
        owner = np.repeat(np.arange(len(gsel)), lens)
        offs = np.arange(int(lens.sum())) - np.repeat(np.cumsum(lens) - lens, lens)
        nbr = csr.col_idx[np.repeat(starts, lens) + offs].astype(np.int64)
        is_pair = go_mask[nbr]
        pair_owner = owner[is_pair]
        counts = np.bincount(pair_owner, minlength=len(gsel))
        with_go = counts > 0
        seg_ptr = np.concatenate([[0], np.cumsum(counts[with_go])]).astype(np.int64)
        return {'node_index': node_index, 'gene_attribs': gene_attr_list,
                'pair_gene': gid[gsel][pair_owner], 'pair_go': nbr[is_pair], 'seg_ptr': seg_ptr,
                'gene_of_seg': gsel[with_go].tolist()}

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
        if pairs is None:
            pairs = self.collect_pairs()
        m = len(pairs['pair_go'])
        nreps = len(self.nvs) if sims is None else int(sims.shape[0])
        self._nreps = nreps
        core = None
        if m > 0:
            if sims is None:
                sims = torch.empty((nreps, m), dtype=torch.float32, device='cuda')
                for r, nv in enumerate(self.nvs):
                    self.pair_similarities(nv, pairs, out=sims[r])
            core = self._numeric_core(sims.to('cuda').contiguous(), pairs['seg_ptr'], alpha_fdr)
        return self.assemble_table(pairs, core, alpha_fdr, nreps)

    def _numeric_core(self, sims, seg_ptr, alpha_fdr):
        """Everything numeric of perform_statistics.py:100-187 and :260-283 for all pairs and
        replicates at once, on the device: psim, per-gene BH (one segment per gene and replicate),
        log_stats of q- and p-values, mean / SEM of the similarities, the row filter
        ``gene_padj < alpha_fdr or alpha_fdr == 1``, and the global BH over the kept rows (one
        segment per replicate) with its log_stats.  Returns host arrays for the kept rows."""
        import torch
        nreps, m = int(sims.shape[0]), int(sims.shape[1])
        dev = sims.device
        pvals = self._psim_batch(sims.reshape(-1)).reshape(nreps, m)
        seg_all = np.concatenate([seg_ptr[:-1] + r * m for r in range(nreps)] + [[nreps * m]])
        qvals = self._bh(pvals.reshape(-1), _dev(seg_all, np.int64)).reshape(nreps, m)
        q_stats = self._log_stats_rows(qvals.t())       # [m, 3]
        p_stats = self._log_stats_rows(pvals.t())
        mean_sim = torch.empty(m, dtype=torch.float32, device=dev)
        sem_sim = torch.empty(m, dtype=torch.float64, device=dev)
        sims_t = sims.t().contiguous()
        _lib.call('gw_mean_sem_f32', _lib.ptr(sims_t), m, nreps, _lib.ptr(mean_sim),
                  _lib.ptr(sem_sim), _lib.stream_ptr())
        if alpha_fdr == 1:
            kept = torch.arange(m, device=dev)
        else:
            kept = torch.nonzero(q_stats[:, 0] < alpha_fdr).reshape(-1)
        rows = int(kept.numel())
        g_stats = torch.empty((rows, 3), dtype=torch.float64, device=dev)
        if rows > 0:
            P = pvals.index_select(1, kept).contiguous()                 # [nreps, rows]
            seg = torch.arange(nreps + 1, dtype=torch.int64, device=dev) * rows
            g_stats = self._log_stats_rows(self._bh(P.reshape(-1), seg).reshape(nreps, rows).t())
        self.sims = sims.cpu().numpy()                                   # kept for inspection / tests
        table = torch.cat([q_stats.index_select(0, kept), p_stats.index_select(0, kept),
                           mean_sim.index_select(0, kept).double().unsqueeze(1),
                           sem_sim.index_select(0, kept).unsqueeze(1), g_stats], dim=1).cpu().numpy()
        return {'kept': kept.cpu().numpy(), 'gene_padj': table[:, 0:3], 'pval': table[:, 3:6],
                'sim': table[:, 6].astype(np.float32), 'sem_sim': table[:, 7],
                'global_padj': table[:, 8:11]}

    def assemble_table(self, pairs, core, alpha_fdr, nreps):
        """The output table of perform_statistics.py:158-245 from the numeric columns of the kept
        (gene, GO) rows: column-wise (no per-row Python), same columns, order, dtypes and sort as
        the reference.  Host-only."""
        genes = self.genes
        graph = self.graph
        ga = pairs['gene_attribs']
        seg_ptr = np.asarray(pairs['seg_ptr'], dtype=np.int64)
        gene_of_seg = np.asarray(pairs['gene_of_seg'], dtype=np.int64)
        kept = core['kept'] if core is not None else np.zeros(0, dtype=np.int64)
        gene_of_pair = np.repeat(gene_of_seg, np.diff(seg_ptr))
        row_gene = gene_of_pair[kept]
        if alpha_fdr == 1:   # genes without GO connections, or absent from the network
            has_rows = np.zeros(len(genes), dtype=bool)
            has_rows[gene_of_seg] = True
            empty = np.nonzero(~has_rows)[0]
        else:
            empty = np.zeros(0, dtype=np.int64)
        nfull, nrow = len(kept), len(kept) + len(empty)
        # the reference emits rows gene by gene: merge the two kinds by gene index (pair rows keep
        # their adjacency order inside a gene)
        all_gene = np.concatenate([row_gene, empty])
        order = np.argsort(all_gene, kind='stable')
        all_gene = all_gene[order]
        is_full = (order < nfull)

        def column(full_vals, empty_val, dtype=None):
            col = np.empty(nrow, dtype=object if dtype is None else dtype)
            col[:nfull] = full_vals
            col[nfull:] = empty_val
            return col[order]

        def numeric(vals):
            return column(vals, np.nan, np.float64)

        nodes = np.array(list(graph.nodes()), dtype=object)
        go_ids = nodes[np.asarray(pairs['pair_go'], dtype=np.int64)[kept]]
        uniq, inv = np.unique(np.asarray(pairs['pair_go'], dtype=np.int64)[kept], return_inverse=True)
        node_data = graph.nodes
        u_name = np.array([node_data[nodes[t]]['name'] for t in uniq], dtype=object)
        u_domain = np.array([node_data[nodes[t]]['domain'].replace('_', ' ') for t in uniq], dtype=object)
        u_ncon = np.array([len(graph[nodes[t]]) for t in uniq], dtype=np.float64)
        by_gene = lambda key: np.array([ga[g][key] for g in all_gene], dtype=object)   # noqa: E731
        ncon_gene = np.array([ga[g]['ncon_gene'] for g in all_gene], dtype=np.float64)
        ncon_go = numeric(u_ncon[inv] if nfull else np.zeros(0))
        data = {}
        if self.gene_id_type in {'mgi_id', 'rgd_id', 'ensembl_id', 'entrez_human', 'entrez_mouse',
                                 'custom'}:
            data[self.gene_id_type] = np.array([self._prefix(genes[g])[0] for g in all_gene], dtype=object)
        data['hgnc_symbol'] = by_gene('hgnc_symbol')
        data['hgnc_id'] = by_gene('hgnc_id')
        data['go_name'] = column(u_name[inv] if nfull else [], '')
        data['go_id'] = column(go_ids, '')
        data['go_domain'] = column(u_domain[inv] if nfull else [], '')
        # from_records infers int64 for a column without NaN and float64 otherwise; astype('str')
        # below then prints '12' or '12.0' accordingly (perform_statistics.py:233-234)
        data['ncon_gene'] = ncon_gene if np.isnan(ncon_gene).any() else ncon_gene.astype(np.int64)
        data['ncon_go'] = ncon_go if np.isnan(ncon_go).any() else ncon_go.astype(np.int64)
        z3 = np.zeros((0, 3))
        gp = core['gene_padj'] if core is not None else z3
        pv = core['pval'] if core is not None else z3
        gl = core['global_padj'] if core is not None else z3
        data['global_padj'] = numeric(gl[:, 0])
        data['gene_padj'] = numeric(gp[:, 0])
        data['pval'] = numeric(pv[:, 0])
        # np.mean of float32 similarities is a float32: the column stays float32 unless an empty
        # row's NaN (a Python float) widens it (perform_statistics.py:180, 126-131)
        data['sim'] = numeric(core['sim'].astype(np.float64) if core is not None else [])
        if len(empty) == 0 and nfull > 0:
            data['sim'] = data['sim'].astype(np.float32)
        data['sem_sim'] = numeric(core['sem_sim'] if core is not None else [])
        data['cilow_global_padj'] = numeric(gl[:, 1])
        data['ciupp_global_padj'] = numeric(gl[:, 2])
        data['cilow_gene_padj'] = numeric(gp[:, 1])
        data['ciupp_gene_padj'] = numeric(gp[:, 2])
        data['cilow_pval'] = numeric(pv[:, 1])
        data['ciupp_pval'] = numeric(pv[:, 2])
        df = pd.DataFrame(data)
        if nrow == 0:   # an empty from_records frame has object columns throughout
            df = df.astype(object)
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
        file (one BH segment per replicate, gw_bh_segmented; then gw_log_stats).
        ``generate_output`` does this on the device before the table exists; this method keeps the
        reference's DataFrame-level entry point (perform_statistics.py:260-283)."""
        nreps = getattr(self, '_nreps', len(self.nvs))
        names = ['global_padj', 'cilow_global_padj', 'ciupp_global_padj']
        ids = df[~df['pval_rep0'].isna()].index
        stats = np.zeros((0, 3))
        if len(ids) > 0:
            P = np.stack([df['pval_rep' + str(i)][ids].to_numpy(dtype=np.float64) for i in range(nreps)])
            seg = _dev(np.arange(nreps + 1, dtype=np.int64) * len(ids), np.int64)
            q = self._bh(_dev(P.reshape(-1), np.float64), seg).reshape(nreps, len(ids))
            stats = self._log_stats_rows(q.t()).cpu().numpy()
        for k, (key, back) in enumerate(zip(names, (8, 4, 4))):
            df.insert(len(df.columns) - back - nreps, key, np.nan)
            if len(ids) > 0:
                df.loc[ids, key] = stats[:, k]
        return df
