"""Statistical-parity fixtures of the SGNS trainer at the metric's own setting (nreps 10/10).

The SEQUENTIAL CPU oracle (oracle/gw_oracle.c: gensim's update rule, schedule and -- `gensim`
mode -- its cum_table + LCG negative sampler; `device` mode: the Philox + alias sampler the CUDA
kernel uses) trains nreps_graph = nreps_null = 10 replicates of the C1-shaped network for three
seeds; this script packs what the acceptance test needs into tests/golden/parity_<case>_s<seed>.npz:

  null_q        8193 quantiles (k/8192) of the pooled, sorted null similarities   [gensim mode]
  null_mean/sd, n_null
  sims_mean     per gene-GO pair: mean over the 10 real replicates                [gensim mode]
  gene_padj     per pair, perform_statistics.py:100-187 (psim -> per-gene BH -> gmean over reps)
  floor_*       the same comparison oracle(device sampler) vs oracle(gensim sampler): what two
                SEQUENTIAL runs that differ only in the negative-sampling RNG score -- the floor of
                any comparison against the oracle (KS, Jaccard at alpha_fdr = 0.1, |dsim|)

Step 1 (CPU, ~6 min per replicate and core; 120 replicates):
  python tools/parity_study.py --side oracle --mode gensim --case c1mod --nreps 10 --out gpurun_out/parity10
  python tools/parity_study.py --side oracle --mode device --case c1mod --nreps 10 --out gpurun_out/parity10
Step 2:
  python tests/golden/make_parity_fixture.py --case c1mod --nreps 10 --src gpurun_out/parity10

Both sides of the later comparison regenerate the same walks and randomized graphs from the same
seeds (walk and configuration-model kernels are bit-exact against the oracle), so the differences
are the trainer's alone.  The generator imports oracle/ (test infrastructure).
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

NQ = 8192


def quantiles(sorted_vals):
    idx = np.round(np.linspace(0, len(sorted_vals) - 1, NQ + 1)).astype(np.int64)
    return sorted_vals[idx].astype(np.float32)


def ks_against_quantiles(q, sample):
    """Two-sample KS statistic between a sample and a distribution given by its k/NQ quantiles
    (exact up to 1/NQ)."""
    sample = np.sort(np.asarray(sample))
    # ECDF of the sample at the quantile points vs k/NQ, and the reverse
    f_s = np.searchsorted(sample, q, side='right') / len(sample)
    f_q = np.arange(len(q)) / (len(q) - 1)
    return float(np.abs(f_s - f_q).max())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--case', default='c1mod')
    ap.add_argument('--seeds', default='0,1,2')
    ap.add_argument('--nreps', type=int, default=10)
    ap.add_argument('--src', default='gpurun_out/parity10')
    ap.add_argument('--dst', default=os.path.join(ROOT, 'tests', 'golden'))
    ap.add_argument('--compact', action='store_true',
                    help='large networks: store the significant set as packed bits and the mean '
                         'similarities as float16, no pair index arrays (the test recomputes the pairs)')
    a = ap.parse_args()
    from scipy.stats import ks_2samp
    from oracle import stats_ref
    from tools.parity_study import build_case, pairs_of, load_side
    for seed in [int(x) for x in a.seeds.split(',')]:
        mg, genes = build_case(a.case, seed)
        pg, po, seg = pairs_of(mg, genes)
        G = load_side(a.src, a.case, 'oracle-gensim', seed, a.nreps)
        D = load_side(a.src, a.case, 'oracle-device', seed, a.nreps)
        assert G is not None, 'oracle-gensim outputs missing for seed %d' % seed
        padj_g = stats_ref.gene_padj(G[0], G[1], seg)
        if a.compact:
            out = dict(case=a.case, seed=seed, nreps=a.nreps, m=len(po), null_q=quantiles(G[1]),
                       null_mean=float(G[1].mean()), null_sd=float(G[1].std()), n_null=len(G[1]),
                       sims_mean=G[0].mean(axis=0).astype(np.float16),
                       sig_bits=np.packbits(padj_g < 0.1))
        else:
            out = dict(case=a.case, seed=seed, nreps=a.nreps, pair_gene=pg.astype(np.int32),
                       pair_go=po.astype(np.int32), seg_ptr=seg, null_q=quantiles(G[1]),
                       null_mean=float(G[1].mean()), null_sd=float(G[1].std()), n_null=len(G[1]),
                       sims_mean=G[0].mean(axis=0).astype(np.float32), gene_padj=padj_g)
        if D is not None:
            padj_d = stats_ref.gene_padj(D[0], D[1], seg)
            sg, sd = padj_g < 0.1, padj_d < 0.1
            out.update(floor_ks=float(ks_2samp(G[1], D[1]).statistic),
                       floor_jaccard=float((sg & sd).sum() / max(1, (sg | sd).sum())),
                       floor_dsim=float(np.abs(G[0].mean(axis=0) - D[0].mean(axis=0)).mean()),
                       floor_null_mean=float(D[1].mean()), floor_nsig=int(sd.sum()))
        out['nsig'] = int((padj_g < 0.1).sum())
        f = os.path.join(a.dst, 'parity_%s_s%d.npz' % (a.case, seed))
        np.savez_compressed(f, **out)
        print(f, {k: v for k, v in out.items() if np.isscalar(v)},
              'ks(self, quantised) %.5f' % ks_against_quantiles(out['null_q'], G[1]))


if __name__ == '__main__':
    main()
