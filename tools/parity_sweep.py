"""GPU side of the parity study in one process: trains ONE real and ONE null replicate of a case
under many trainer shapes and scores each against sequential-oracle outputs shipped in
``parity_ref/`` (produced by ``tools/parity_study.py --side oracle`` on CPU; same seeds, same walks,
same randomized graph).  Prints a table and writes ``<out>/<case>_sweep.json``.

  python tools/parity_sweep.py --case mid --specs a:256:1000:chunk:per_pair,b:1024:5000:burst:per_pair

A spec is tag:lanes:unit_walks:policy:reductions[:extra=val...]; extra keys go to word2vec().
This tool imports oracle/ (a checker, like tests/); it is not part of the product.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tools.parity_study import build_case, pairs_of, mix, NITER, WL, D, K, EPOCHS  # noqa: E402


def parse_spec(s):
    f = s.split(':')
    kw = dict(lanes=int(f[1]), unit_walks=int(f[2]), policy=f[3], reductions=f[4])
    for extra in f[5:]:
        k, v = extra.split('=')
        kw[k] = int(v) if v.lstrip('-').isdigit() else v
    return f[0], kw


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--case', default='mid')
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--specs', required=True)
    ap.add_argument('--ref', default='parity_ref')
    ap.add_argument('--ref-side', default='oracle-gensim')
    ap.add_argument('--out', default='gpurun_out/sweep')
    ap.add_argument('--null-only', action='store_true')
    ap.add_argument('--save', action='store_true', help='also save the similarity arrays')
    a = ap.parse_args()
    import torch
    from scipy.stats import ks_2samp
    from oracle import stats_ref
    from genewalk_b200.deepwalk import DeepWalk
    from genewalk_b200.null_distributions import get_rand_graph, get_null_distributions
    from genewalk_b200.perform_statistics import GeneWalk
    from genewalk_b200.csr import GraphCSR
    os.makedirs(a.out, exist_ok=True)
    seed = a.seed
    mg, genes = build_case(a.case, seed)
    csr = GraphCSR.from_networkx(mg)
    gw = GeneWalk(mg, genes, [], np.zeros(1, np.float32), gene_id_type='custom')
    pairs = gw.collect_pairs()
    _, _, seg = pairs_of(mg, genes)
    ref_null = np.sort(np.load(os.path.join(a.ref, '%s_%s_s%d_null0.npy' % (a.case, a.ref_side, seed))))
    fr = os.path.join(a.ref, '%s_%s_s%d_real0.npy' % (a.case, a.ref_side, seed))
    ref_real = np.load(fr) if os.path.exists(fr) and not a.null_only else None
    ref_sig = stats_ref.gene_padj(ref_real[None, :], ref_null, seg) < 0.1 if ref_real is not None else None
    rows = []
    print('%-14s %7s %7s %7s %8s %8s %8s %8s %8s' % ('tag', 'KS', 'mean', 'sd', 'Jaccard', '|dsim|',
                                                    'maxnorm', 'null_s', 'real_s'), flush=True)
    print('%-14s %7.4f %7.4f %7.4f' % ('ref', 0.0, ref_null.mean(), ref_null.std()), flush=True)
    for spec in a.specs.split(','):
        tag, kw = parse_spec(spec)
        salt = int(kw.pop('salt', 0))
        t0 = time.time()
        rg = get_rand_graph(mg, seed=mix(seed, 2, 0), materialize=False)
        dw = DeepWalk(rg, WL, NITER, seed=mix(seed, 3, 0))
        dw.get_walks()
        dw.word2vec(vector_size=D, negative=K, epochs=EPOCHS, sgns_seed=mix(seed, 4 + 16 * salt, 0), **kw)
        null = np.asarray(get_null_distributions(rg, dw.model.wv), dtype=np.float32)
        torch.cuda.synchronize()
        t_null = time.time() - t0
        maxnorm = float(np.linalg.norm(dw.model.wv.vectors, axis=1).max())
        del dw, rg
        ks = float(ks_2samp(ref_null, null).statistic)
        row = dict(tag=tag, spec=kw, salt=salt, ks=ks, null_mean=float(null.mean()), null_sd=float(null.std()),
                   maxnorm=maxnorm, null_s=t_null)
        if a.save:
            np.save(os.path.join(a.out, '%s_gpu-%s_s%d_null0.npy' % (a.case, tag, seed)), null)
        if ref_real is not None:
            t0 = time.time()
            dw = DeepWalk(csr, WL, NITER, seed=mix(seed, 0, 0))
            dw.get_walks()
            dw.word2vec(vector_size=D, negative=K, epochs=EPOCHS, sgns_seed=mix(seed, 1 + 16 * salt, 0), **kw)
            sims = gw.pair_similarities(dw.model.wv, pairs).cpu().numpy()
            torch.cuda.synchronize()
            row['real_s'] = time.time() - t0
            row['maxnorm_real'] = float(np.linalg.norm(dw.model.wv.vectors, axis=1).max())
            del dw
            sig = stats_ref.gene_padj(sims[None, :], np.sort(null), seg) < 0.1
            inter, union = (sig & ref_sig).sum(), (sig | ref_sig).sum()
            row['jaccard'] = float(inter / union) if union else 1.0
            row['nsig'] = int(sig.sum())
            row['nsig_ref'] = int(ref_sig.sum())
            row['dsim'] = float(np.abs(sims - ref_real).mean())
            if a.save:
                np.save(os.path.join(a.out, '%s_gpu-%s_s%d_real0.npy' % (a.case, tag, seed)), sims)
        rows.append(row)
        print('%-14s %7.4f %7.4f %7.4f %8.4f %8.4f %8.2f %8.2f %8.2f' % (
            tag, ks, row['null_mean'], row['null_sd'], row.get('jaccard', float('nan')),
            row.get('dsim', float('nan')), maxnorm, t_null, row.get('real_s', float('nan'))), flush=True)
        with open(os.path.join(a.out, '%s_sweep.json' % a.case), 'w') as f:
            json.dump(dict(case=a.case, seed=seed, ref_side=a.ref_side, ref_mean=float(ref_null.mean()),
                           ref_sd=float(ref_null.std()), rows=rows), f, indent=1)


if __name__ == '__main__':
    main()
