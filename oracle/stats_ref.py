"""numpy restatement of the similarity statistics (oracle; TEST INFRASTRUCTURE ONLY).

Follows, line by line:
  cosine_similarity      gensim KeyedVectors.similarity as called at
                         /root/reference/genewalk/perform_statistics.py:102
                         (dot(unitvec(a), unitvec(b)), float32)
  null_similarities      /root/reference/genewalk/null_distributions.py:33-48
                         (1 - KeyedVectors.distances(node, other_words=neighbours != node), float32)
  psim                   /root/reference/genewalk/perform_statistics.py:247-258
  fdrcorrection          statsmodels.stats.multitest.fdrcorrection(method='indep') as called at
                         perform_statistics.py:112,273 (statsmodels is absent here; the restatement is
                         pinned against scipy.stats.false_discovery_control in tests/)
  log_stats              /root/reference/genewalk/perform_statistics.py:117-124
gensim/statsmodels are third-party and absent: their formulas are restated from the published
source.
"""
import numpy as np


def cosine_similarity(vectors, a_idx, b_idx):
    """float32 dot(unitvec(a), unitvec(b)) for index pairs."""
    v = np.asarray(vectors, dtype=np.float32)
    a = v[np.asarray(a_idx)]
    b = v[np.asarray(b_idx)]
    na = np.sqrt((a * a).sum(axis=1, dtype=np.float32)).astype(np.float32)
    nb = np.sqrt((b * b).sum(axis=1, dtype=np.float32)).astype(np.float32)
    ua = a * (np.float32(1.0) / na).astype(np.float32)[:, None]
    ub = b * (np.float32(1.0) / nb).astype(np.float32)[:, None]
    return (ua * ub).sum(axis=1, dtype=np.float32).astype(np.float32)


def null_similarities(row_ptr, col_idx, vectors):
    """For every node u and every distinct neighbour v != u (CSR order):
    1 - (1 - dot(V[v], V[u]) / (|V[u]| * |V[v]|)), float32 (null_distributions.py:39-47)."""
    row_ptr = np.asarray(row_ptr)
    col_idx = np.asarray(col_idx)
    n = len(row_ptr) - 1
    src = np.repeat(np.arange(n), np.diff(row_ptr))
    keep = src != col_idx
    u = src[keep]
    v = col_idx[keep]
    vec = np.asarray(vectors, dtype=np.float32)
    norms = np.sqrt((vec * vec).sum(axis=1, dtype=np.float32)).astype(np.float32)
    dots = (vec[v] * vec[u]).sum(axis=1, dtype=np.float32).astype(np.float32)
    sims = dots / (norms[u] * norms[v])
    dist = np.float32(1.0) - sims
    return (np.float32(1.0) - dist).astype(np.float32)


def psim_rank(srd, sims):
    return np.searchsorted(np.asarray(srd), np.asarray(sims), side='left')


def psim(srd, sims):
    """max(1 - rank/len(srd), 1e-16) in float64 (perform_statistics.py:252-257)."""
    rank = psim_rank(srd, sims)
    p = 1.0 - rank.astype(np.float64) / float(len(srd))
    return np.maximum(p, 1e-16)


def fdrcorrection(pvals):
    """Benjamini-Hochberg adjusted p-values, statsmodels fdrcorrection(method='indep')[1]."""
    pvals = np.asarray(pvals, dtype=np.float64)
    n = len(pvals)
    if n == 0:
        return pvals.copy()
    order = np.argsort(pvals, kind='stable')
    ps = pvals[order]
    ecdf = np.arange(1, n + 1) / float(n)
    raw = ps / ecdf
    corr = np.minimum.accumulate(raw[::-1])[::-1]
    corr[corr > 1] = 1
    out = np.empty_like(corr)
    out[order] = corr
    return out


def fdrcorrection_segmented(pvals, seg_ptr):
    out = np.empty(len(pvals), dtype=np.float64)
    for s in range(len(seg_ptr) - 1):
        a, b = int(seg_ptr[s]), int(seg_ptr[s + 1])
        out[a:b] = fdrcorrection(pvals[a:b])
    return out


def log_stats(vals):
    """perform_statistics.py:117-124 with scipy gmean/gstd spelled out:
    gmean = exp(mean(log v)); gstd = exp(std(log v, ddof=1))."""
    eps = 1e-16
    vals = np.asarray(vals, dtype=np.float64) + eps
    nreps = len(vals)
    logs = np.log(vals)
    g_mean = np.exp(logs.mean()) - eps
    if nreps > 1:
        g_std = np.exp(np.sqrt(((logs - logs.mean()) ** 2).sum() / (nreps - 1)))
    else:
        g_std = eps
    return (g_mean, g_mean * (g_std ** (-1.96 / np.sqrt(nreps))),
            g_mean * (g_std ** (1.96 / np.sqrt(nreps))))


def gene_padj(sims, srd_sorted, seg_ptr):
    """Per-pair gene_padj of perform_statistics.py:100-187 from raw similarities:
    sims [nreps, m] -> psim -> per-gene BH per replicate -> log_stats()[0] across replicates."""
    sims = np.atleast_2d(np.asarray(sims, dtype=np.float32))
    nreps, m = sims.shape
    q = np.empty((nreps, m))
    for r in range(nreps):
        q[r] = fdrcorrection_segmented(psim(srd_sorted, sims[r]), seg_ptr)
    eps = 1e-16
    return np.exp(np.log(q + eps).mean(axis=0)) - eps
