"""Pure-Python statement of the SGNS update in both reduction orders (small cases only).

TEST INFRASTRUCTURE ONLY, like everything under oracle/.  A third, independent statement next to
oracle/gw_oracle.c and the CUDA kernel, written from the definitions rather than from either of
them: used by tests/test_oracle.py to pin the C oracle's two orders bit for bit.

  per_pair  gensim w2v_fast_sentence_sg_neg (word2vec_inner.pyx) as reached from
            /root/reference/genewalk/deepwalk.py:135-139: every (centre, context) pair updates the
            tables before the next pair is looked at.
  merged    the kernel's Hogwild order (include/genewalk_b200.h, GW_SGNS_MERGED): per step (= both
            pairs of a centre) ONE addition to syn1neg[centre]; the update of syn0[walk[j]] computed
            in step j-1 is held back and added, together with the one computed in step j+1, in step
            j+1; a held-back update is part of the row its owner computes on.
Negatives: Philox4x32-10 + alias table (sampler 1 of the C oracle).
"""
import numpy as np

from . import philox

F = np.float32


def _dot(a, b):
    """Summation order of the kernel / gw_oracle.c:dot_rows."""
    d = len(a)
    G = 1
    if d >= 4 and d & (d - 1) == 0:
        G = min(32, d // 2)
    E = d // G
    c = []
    for g in range(G):
        s = F(a[g * E] * b[g * E])
        for k in range(1, E):
            s = F(s + F(a[g * E + k] * b[g * E + k]))
        c.append(s)
    m = 1
    while m < G:
        for g in range(0, G, 2 * m):
            c[g] = F(c[g] + c[g + m])
        m *= 2
    return c[0]


def _sigmoid(f, table):
    if f <= F(-6.0) or f >= F(6.0):
        return None
    i = int(float(F(f + F(6.0))) * (1000.0 / 6.0 / 2.0))
    return table[min(i, 999)]


def _negatives(p, q, ep, K, V, prob, alias, seed):
    out, words = [], None
    for t in range(K):
        w = 2 * t
        if w % 4 == 0:
            r = philox.philox4x32_10(np.uint32(p & 0xffffffff), np.uint32(p >> 32),
                                     np.uint32((q << 4) | (w >> 2)), np.uint32((ep << 8) | 0x53),
                                     np.uint32(seed & 0xffffffff), np.uint32(seed >> 32))
            words = [int(x) for x in r]
        x0, x1 = words[w % 4], words[w % 4 + 1]
        slot = (x0 * V) >> 32
        thr = min(65535, int(np.floor(F(prob[slot]) * F(65536.0))))
        out.append(slot if (x1 >> 16) < thr else int(alias[slot]))
    return out


def _alpha(alpha0, alpha_min, ep, epochs, pushed, W):
    progress = (ep + pushed / W) / epochs
    return F(max(alpha_min, alpha0 - (alpha0 - alpha_min) * progress))


def _pair(ctx, centre, centre_row, negs, alpha, table, syn1neg, on_centre):
    """One pair on the context row `ctx` (a copy).  Negative rows are updated in syn1neg at once;
    the centre's update is handed to on_centre(delta).  Returns the update of the context row."""
    d = len(ctx)
    work = np.zeros(d, F)
    for t in range(len(negs) + 1):
        if t == 0:
            row, label = centre_row, F(1.0)
        else:
            if negs[t - 1] == centre:
                continue
            row, label = syn1neg[negs[t - 1]], F(0.0)
        sig = _sigmoid(_dot(ctx, row), table)
        if sig is None:
            continue
        g = F(F(label - sig) * alpha)
        delta = np.array([F(g * ctx[k]) for k in range(d)], F)
        for k in range(d):
            work[k] = F(work[k] + F(g * row[k]))
        if t == 0:
            on_centre(delta)
        else:
            for k in range(d):
                row[k] = F(row[k] + delta[k])
    return work


def train(tokens, V, syn0, syn1neg, prob, alias, seed, table, K=5, epochs=1, alpha0=0.025,
          alpha_min=0.0001, job_walks=1000, merged=False):
    L, W = tokens.shape
    d = syn0.shape[1]
    for ep in range(epochs):
        for p in range(W):
            walk = [int(x) for x in tokens[:, p]]
            if -1 in walk:
                walk = walk[:walk.index(-1)]
            alpha = _alpha(alpha0, alpha_min, ep, epochs, (p // job_walks) * job_walks, W)
            held = {}     # position j -> update of syn0[walk[j]] held back since step j-1
            for i, c in enumerate(walk):
                if not merged:
                    for j in (i - 1, i + 1):
                        if 0 <= j < len(walk):
                            negs = _negatives(p, 2 * i + (j > i), ep, K, V, prob, alias, seed)
                            crow = syn1neg[c]

                            def on_centre(delta, crow=crow):
                                for k in range(d):
                                    crow[k] = F(crow[k] + delta[k])
                            ctx = syn0[walk[j]]
                            work = _pair(ctx.copy(), c, crow, negs, alpha, table, syn1neg, on_centre)
                            for k in range(d):
                                ctx[k] = F(ctx[k] + work[k])
                    continue
                # merged order
                crow = syn1neg[c].copy()       # the worker's running copy of the centre row
                total = np.zeros(d, F)
                touched = []

                def on_centre(delta):
                    for k in range(d):
                        crow[k] = F(crow[k] + delta[k])
                        total[k] = F(total[k] + delta[k])
                    touched.append(1)
                left = None
                if i - 1 >= 0:
                    a = walk[i - 1]
                    left = syn0[a].copy()
                    pend = held.pop(i - 1, None)
                    if pend is not None:
                        left = np.array([F(left[k] + pend[k]) for k in range(d)], F)
                    work = _pair(left, c, crow, _negatives(p, 2 * i, ep, K, V, prob, alias, seed),
                                 alpha, table, syn1neg, on_centre)
                    left = np.array([F(left[k] + work[k]) for k in range(d)], F)
                    if pend is not None:
                        work = np.array([F(pend[k] + work[k]) for k in range(d)], F)
                    for k in range(d):
                        syn0[a][k] = F(syn0[a][k] + work[k])
                if i + 1 < len(walk):
                    b = walk[i + 1]
                    right = left.copy() if (left is not None and b == walk[i - 1]) else syn0[b].copy()
                    held[i + 1] = _pair(right, c, crow,
                                        _negatives(p, 2 * i + 1, ep, K, V, prob, alias, seed),
                                        alpha, table, syn1neg, on_centre)
                if touched:
                    for k in range(d):
                        syn1neg[c][k] = F(syn1neg[c][k] + total[k])
            for j in sorted(held):       # the walk ended before the held-back updates came due
                for k in range(d):
                    syn0[walk[j]][k] = F(syn0[walk[j]][k] + held[j][k])
