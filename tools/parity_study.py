"""Statistical parity study: GPU Hogwild SGNS vs the sequential CPU oracle (north_star criteria:
KS agreement of the null similarity distributions, Jaccard of FDR-significant gene-GO pairs).

  python tools/parity_study.py --side oracle --mode gensim  --out gpurun_out/study   (CPU, here)
  python tools/parity_study.py --side oracle --mode device  --out gpurun_out/study   (CPU, here)
  python tools/parity_study.py --side gpu --tag default --out gpurun_out/study      (B200 box)
  python tools/parity_study.py --side compare --out gpurun_out/study

Both sides regenerate the same corpora from the same seeds (walks and randomized graphs are
bit-exact between the CUDA path and the oracle), so the only differences are the trainer's.
This tool imports oracle/ (it is a checker, like tests/); it is not part of the product.
"""
import argparse
import glob
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from genewalk_b200 import workloads as graphs  # noqa: E402

NITER, WL, D, K, EPOCHS = 100, 10, 8, 5, 5


def mix(seed, kind, rep):
    """splitmix64 of (seed, kind, rep) -> 64-bit stream seed."""
    x = (seed * 0x9E3779B97F4A7C15 + kind * 0xBF58476D1CE4E5B9 + rep * 0x94D049BB133111EB + 1) % 2 ** 64
    x ^= x >> 30; x = (x * 0xBF58476D1CE4E5B9) % 2 ** 64
    x ^= x >> 27; x = (x * 0x94D049BB133111EB) % 2 ** 64
    x ^= x >> 31
    return x


def build_case(name, seed):
    if name == 'c1mod':
        mg, genes = graphs.modular_gene_go_network(seed=100 + seed)
    elif name == 'tiny':
        mg, genes = graphs.modular_gene_go_network(300, 60, 8, 1500, 900, 100, seed=100 + seed)
    elif name == 'c2mod':
        mg, genes = graphs.modular_gene_go_network(20000, 18000, 400, 700000, 250000, 50000,
                                                   seed=100 + seed)
    elif name == 'mid':  # between C1 and C2: n = 8500, ~0.2 M edges (oracle-vs-oracle floor studies)
        mg, genes = graphs.modular_gene_go_network(4500, 4000, 100, 150000, 55000, 11000,
                                                   seed=100 + seed)
    elif name == 'c2':   # the bench workload (BASELINE.json configs[1])
        n, u, v, is_go = graphs.human_scale_edges(seed=2 + seed)
        mg, genes = graphs.edges_to_networkx(n, u, v, is_go, 20000)
    else:
        raise ValueError(name)
    return mg, genes


def pairs_of(mg, genes):
    import networkx as nx
    index = {v: i for i, v in enumerate(mg.nodes())}
    go = set(nx.get_node_attributes(mg, 'GO'))
    pg, po, seg = [], [], [0]
    for g in genes:
        gid = g['ID']
        if gid not in mg:
            continue
        con = [w for w in mg[gid] if w in go]
        if con:
            pg += [index[gid]] * len(con)
            po += [index[w] for w in con]
            seg.append(len(po))
    return np.array(pg), np.array(po), np.array(seg, dtype=np.int64)


# ---------------------------------------------------------------------------------------------
def oracle_vectors(row_ptr, col_idx, walk_seed, sgns_seed, mode, nthreads=1):
    from oracle import lib as olib
    n = len(row_ptr) - 1
    if mode == 'gensim':      # reference restatement: corpus order, cum_table + LCG, job alpha
        tok = olib.walk_uniform(row_ptr, col_idx, NITER, WL, walk_seed, 0)
        itn, syn0, _ = olib.word2vec_gensim_like(tok, n, d=D, K=K, epochs=EPOCHS, nthreads=nthreads,
                                                 inplace=True)
        vec = np.zeros((n, D), dtype=np.float32)
        vec[itn] = syn0
        return vec
    tok = olib.walk_uniform(row_ptr, col_idx, NITER, WL, walk_seed, 0)   # corpus order
    counts = np.bincount(tok.ravel(), minlength=n)
    prob, alias = olib.alias_table(counts)
    order = np.argsort(-counts, kind='stable')
    V = int((counts > 0).sum())
    rows, _ = olib.gensim_init_weights(n, D, seed=1)
    syn0 = np.zeros((n, D), dtype=np.float32)
    syn0[order[:V]] = rows[:V]
    syn1 = np.zeros((n, D), dtype=np.float32)
    olib.sgns_train(tok, n, D, syn0, syn1, sampler=1, K=K, epochs=EPOCHS, alias_prob=prob,
                    alias_idx=alias, seed=sgns_seed, nthreads=nthreads)
    return syn0


def oracle_unit(args):
    case, seed, kind, rep, mode = args
    from oracle import config_model_ref, stats_ref
    from genewalk_b200.csr import GraphCSR, multigraph_degrees
    mg, genes = build_case(case, seed)
    t0 = time.time()
    if kind == 'real':
        csr = GraphCSR.from_networkx(mg)
        vec = oracle_vectors(csr.row_ptr, csr.col_idx, mix(seed, 0, rep), mix(seed, 1, rep), mode)
        pg, po, _ = pairs_of(mg, genes)
        out = stats_ref.cosine_similarity(vec, pg, po)
    else:
        d_seq = config_model_ref.multigraph_degrees_desc(multigraph_degrees(mg))
        u, v = config_model_ref.config_model_edges(d_seq, mix(seed, 2, rep))
        rp, ci = config_model_ref.csr_distinct(len(d_seq), u, v)
        vec = oracle_vectors(rp, ci, mix(seed, 3, rep), mix(seed, 4, rep), mode)
        out = stats_ref.null_similarities(rp, ci, vec)
    return (seed, kind, rep, out, time.time() - t0)


def run_oracle(a):
    import multiprocessing as mp
    jobs = [(a.case, s, kind, r, a.mode) for s in a.seeds for kind in ('real', 'null')
            for r in range(a.nreps)
            if not os.path.exists(os.path.join(a.out, '%s_oracle-%s_s%d_%s%d.npy' % (a.case, a.mode, s, kind, r)))]
    if not jobs:
        return
    with mp.Pool(min(a.procs, len(jobs))) as pool:
        for seed, kind, rep, out, dt in pool.imap_unordered(oracle_unit, jobs):
            f = os.path.join(a.out, '%s_oracle-%s_s%d_%s%d.npy' % (a.case, a.mode, seed, kind, rep))
            np.save(f, out)
            print('oracle %s seed %d %s %d: %d values in %.1fs' % (a.mode, seed, kind, rep, len(out), dt),
                  flush=True)


# ---------------------------------------------------------------------------------------------
def run_gpu(a):
    """GPU side: nreps real + nreps null replicates per seed with the given trainer shape
    (0 = library default); `tools/parity_sweep.py` is the many-shapes-in-one-process variant."""
    import torch
    from genewalk_b200.deepwalk import DeepWalk
    from genewalk_b200.null_distributions import get_rand_graph, get_null_distributions
    from genewalk_b200.perform_statistics import GeneWalk
    from genewalk_b200.csr import GraphCSR
    kw = dict(lanes=a.conc[0], unit_walks=a.chunk, reductions=a.reductions)
    tag = a.tag or 'default'
    for seed in a.seeds:
        mg, genes = build_case(a.case, seed)
        csr = GraphCSR.from_networkx(mg)
        gw = GeneWalk(mg, genes, [], np.zeros(1, np.float32), gene_id_type='custom')
        pairs = gw.collect_pairs()
        for rep in range(a.nreps):
            dw = DeepWalk(csr, WL, NITER, seed=mix(seed, 0, rep))
            dw.get_walks()
            dw.word2vec(vector_size=D, negative=K, epochs=EPOCHS, sgns_seed=mix(seed, 1 + 16 * a.salt, rep), **kw)
            sims = gw.pair_similarities(dw.model.wv, pairs).cpu().numpy()
            np.save(os.path.join(a.out, '%s_gpu-%s_s%d_real%d.npy' % (a.case, tag, seed, rep)), sims)
            rg = get_rand_graph(mg, seed=mix(seed, 2, rep), materialize=False)
            dw = DeepWalk(rg, WL, NITER, seed=mix(seed, 3, rep))
            dw.get_walks()
            dw.word2vec(vector_size=D, negative=K, epochs=EPOCHS, sgns_seed=mix(seed, 4 + 16 * a.salt, rep), **kw)
            null = np.asarray(get_null_distributions(rg, dw.model.wv), dtype=np.float32)
            np.save(os.path.join(a.out, '%s_gpu-%s_s%d_null%d.npy' % (a.case, tag, seed, rep)), null)
            torch.cuda.synchronize()
            print('gpu %s seed %d rep %d done' % (tag, seed, rep), flush=True)


# ---------------------------------------------------------------------------------------------
def load_side(out, case, side, seed, nreps):
    sims, null = [], []
    for r in range(nreps):
        fs = os.path.join(out, '%s_%s_s%d_real%d.npy' % (case, side, seed, r))
        fn = os.path.join(out, '%s_%s_s%d_null%d.npy' % (case, side, seed, r))
        if not (os.path.exists(fs) and os.path.exists(fn)):
            return None
        sims.append(np.load(fs)); null.append(np.load(fn))
    return np.stack(sims), np.sort(np.concatenate(null))


def run_compare(a):
    from scipy.stats import ks_2samp
    from oracle import stats_ref
    sides = sorted({os.path.basename(f).split('_')[1] for f in glob.glob(os.path.join(a.out, a.case + '_*_s*_real0.npy'))})
    print('sides:', sides)
    ref = a.ref
    rows = []
    for side in sides:
        ks, jac, nsig_ref, nsig, dmean = [], [], [], [], []
        for seed in a.seeds:
            A = load_side(a.out, a.case, ref, seed, a.nreps)
            B = load_side(a.out, a.case, side, seed, a.nreps)
            if A is None or B is None:
                continue
            mg, genes = build_case(a.case, seed)
            _, _, seg = pairs_of(mg, genes)
            pa = stats_ref.gene_padj(A[0], A[1], seg) < a.alpha
            pb = stats_ref.gene_padj(B[0], B[1], seg) < a.alpha
            inter, union = (pa & pb).sum(), (pa | pb).sum()
            jac.append(inter / union if union else 1.0)
            nsig_ref.append(int(pa.sum())); nsig.append(int(pb.sum()))
            ks.append(ks_2samp(A[1], B[1]).statistic)
            dmean.append(float(np.abs(A[0].mean(axis=0) - B[0].mean(axis=0)).mean()))
        if ks:
            rows.append((side, np.mean(ks), np.mean(jac), np.mean(nsig_ref), np.mean(nsig), np.mean(dmean)))
    print('%-16s %8s %8s %9s %9s %10s   (vs %s, alpha_fdr=%g, mean over seeds)' %
          ('side', 'KS', 'Jaccard', 'nsig_ref', 'nsig', '|dsim|', ref, a.alpha))
    for r in rows:
        print('%-16s %8.4f %8.4f %9.1f %9.1f %10.4f' % r)
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--side', required=True, choices=['oracle', 'gpu', 'compare'])
    ap.add_argument('--mode', default='gensim', choices=['gensim', 'device'])
    ap.add_argument('--case', default='c1mod')
    ap.add_argument('--seeds', default='0,1,2')
    ap.add_argument('--nreps', type=int, default=2)
    ap.add_argument('--conc', default='0', help='gpu side: workers per training (0 = library default)')
    ap.add_argument('--chunk', type=int, default=0, help='gpu side: walks per unit (0 = library default)')
    ap.add_argument('--tag', default='')
    ap.add_argument('--reductions', default='per_pair')
    ap.add_argument('--salt', type=int, default=0, help='gpu side: perturb only the negative-sampling seeds')
    ap.add_argument('--procs', type=int, default=os.cpu_count())
    ap.add_argument('--out', default='gpurun_out/study')
    ap.add_argument('--ref', default='oracle-gensim')
    ap.add_argument('--alpha', type=float, default=0.1)
    a = ap.parse_args()
    a.seeds = [int(x) for x in a.seeds.split(',')]
    a.conc = [int(x) for x in a.conc.split(',')]
    os.makedirs(a.out, exist_ok=True)
    {'oracle': run_oracle, 'gpu': run_gpu, 'compare': run_compare}[a.side](a)


if __name__ == '__main__':
    main()
