"""Order-only emulation of the trainer's scheduling policies on the CPU (round 2 parity study).

The sequential oracle trains the null replicate of a case with its walks re-ordered the way `lanes`
concurrent workers would visit them (one walk per worker per tick, workers pull units from a queue
in corpus order).  Every update still sees all earlier ones, so the difference to the canonical run
is the ORDER effect of a policy alone; the GPU run of the same policy adds the Hogwild staleness.

  python tools/order_study.py --case mid --policy chunk:1024:1000 --out gpurun_out/order
  python tools/order_study.py --case mid --policy burst:1024:0     (unit = a start node's walks)
  python tools/order_study.py --case mid --policy burst:1024:5000  (bursts cut into <=5000 walks)

This tool imports oracle/ (a checker, like tests/); it is not part of the product.
"""
import argparse
import heapq
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tools.parity_study import build_case, mix, NITER, WL, D, K, EPOCHS  # noqa: E402


def order_of_units(first, count, lanes):
    """Visit order of the walks when `lanes` workers pull the units (first[i], count[i]) in turn and
    train one walk per tick."""
    free = [(0, l) for l in range(lanes)]
    heapq.heapify(free)
    start = np.empty(len(first), np.int64)
    lane = np.empty(len(first), np.int64)
    for i in range(len(first)):
        t, l = heapq.heappop(free)
        start[i], lane[i] = t, l
        heapq.heappush(free, (t + int(count[i]), l))
    W = int(count.sum())
    walk = np.empty(W, np.int64)
    stamp = np.empty(W, np.int64)
    o = 0
    for i in range(len(first)):
        c = int(count[i])
        walk[o:o + c] = np.arange(first[i], first[i] + c)
        stamp[o:o + c] = (start[i] + np.arange(c)) * lanes + lane[i]
        o += c
    return walk[np.argsort(stamp, kind='stable')]


def units_of(policy, row_ptr, W):
    kind, lanes, size = policy.split(':')[:3]
    lanes, size = int(lanes), int(size)
    if kind == 'chunk':
        first = np.arange(0, W, size, dtype=np.int64)
        count = np.minimum(size, W - first)
    elif kind == 'burst':
        deg = np.diff(row_ptr).astype(np.int64)
        b0 = row_ptr[:-1].astype(np.int64) * NITER
        blen = deg * NITER
        first, count = [], []
        for s, ln in zip(b0, blen):
            if ln == 0:
                continue
            if size <= 0 or ln <= size:
                first.append(s); count.append(ln)
            else:
                k = -(-ln // size)
                for j in range(k):
                    first.append(s + j * size); count.append(min(size, ln - j * size))
        first, count = np.array(first, np.int64), np.array(count, np.int64)
    elif kind == 'burstcap':   # burstcap:lanes:unit:max_pieces -- start-node units, at most max_pieces per node
        cap = int(policy.split(':')[3])
        deg = np.diff(row_ptr).astype(np.int64)
        b0 = row_ptr[:-1].astype(np.int64) * NITER
        blen = deg * NITER
        first, count = [], []
        for s, ln in zip(b0, blen):
            if ln == 0:
                continue
            k = min(-(-ln // size), cap)
            for j in range(k):
                a, b = s + (ln * j) // k, s + (ln * (j + 1)) // k
                first.append(a); count.append(b - a)
        first, count = np.array(first, np.int64), np.array(count, np.int64)
    else:
        raise ValueError(policy)
    return first, count, lanes


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--case', default='mid')
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--policy', required=True)
    ap.add_argument('--out', default='gpurun_out/order')
    ap.add_argument('--ref', default='parity_ref')
    a = ap.parse_args()
    from scipy.stats import ks_2samp
    from oracle import lib as olib
    from oracle import config_model_ref, stats_ref
    from genewalk_b200.csr import multigraph_degrees
    os.makedirs(a.out, exist_ok=True)
    seed = a.seed
    mg, genes = build_case(a.case, seed)
    d_seq = config_model_ref.multigraph_degrees_desc(multigraph_degrees(mg))
    u, v = config_model_ref.config_model_edges(d_seq, mix(seed, 2, 0))
    rp, ci = config_model_ref.csr_distinct(len(d_seq), u, v)
    n = len(rp) - 1
    tok = olib.walk_uniform(rp, ci, NITER, WL, mix(seed, 3, 0), 0)
    W = tok.shape[1]
    t0 = time.time()
    if a.policy != 'canon':
        first, count, lanes = units_of(a.policy, rp, W)
        perm = order_of_units(first, count, lanes)
        assert len(perm) == W and np.array_equal(np.sort(perm[:100000]), np.sort(perm[:100000]))
        tok = np.ascontiguousarray(tok[:, perm])
        print('policy %s: %d units, order built in %.1fs' % (a.policy, len(first), time.time() - t0), flush=True)
    counts = np.bincount(tok.ravel(), minlength=n)
    prob, alias = olib.alias_table(counts)
    order = np.argsort(-counts, kind='stable')
    V = int((counts > 0).sum())
    rows, _ = olib.gensim_init_weights(n, D, seed=1)
    syn0 = np.zeros((n, D), dtype=np.float32)
    syn0[order[:V]] = rows[:V]
    syn1 = np.zeros((n, D), dtype=np.float32)
    t0 = time.time()
    olib.sgns_train(tok, n, D, syn0, syn1, sampler=1, K=K, epochs=EPOCHS, alias_prob=prob,
                    alias_idx=alias, seed=mix(seed, 4, 0), nthreads=1)
    null = stats_ref.null_similarities(rp, ci, syn0)
    tag = a.policy.replace(':', '-')
    np.save(os.path.join(a.out, '%s_order-%s_s%d_null0.npy' % (a.case, tag, seed)), null)
    msg = 'policy %s: trained in %.0fs, null mean %.4f sd %.4f maxnorm %.2f' % (
        a.policy, time.time() - t0, null.mean(), null.std(), np.linalg.norm(syn0, axis=1).max())
    for side in ('oracle-device', 'oracle-gensim'):
        f = os.path.join(a.ref, '%s_%s_s%d_null0.npy' % (a.case, side, seed))
        if os.path.exists(f):
            msg += '  KS vs %s %.4f' % (side, ks_2samp(np.load(f), null).statistic)
    print(msg, flush=True)


if __name__ == '__main__':
    main()
