"""Pin the SGNS oracle against REAL gensim (the unpin procedure of DESIGN.md section 2).

gensim (`gensim>=4.0.0`, /root/reference/setup.py:11; call site deepwalk.py:135-139) is absent from
the reference tree, from this image and from its offline wheelhouse, so oracle/gw_oracle.c restates
gensim's published algorithm and is marked "parity unpinned".  On ANY box that has gensim 4.x and a
C compiler, run

    python tools/pin_gensim.py [--out profiles/gensim_pin.json]

It trains  Word2Vec(sentences, sg=1, vector_size=8, window=1, min_count=1, negative=5, workers=1,
sample=0)  (exactly the reference's call) on a 40-node walk corpus and compares, bit for bit, with
oracle.lib.word2vec_gensim_like (sampler 0: cum_table + 48-bit LCG, per-job seeds from
RandomState(1)):

  check                     what must hold for the pin
  ------------------------  -----------------------------------------------------------------
  index_to_key              same vocabulary order (descending count, ties first-seen)
  cum_table                 np.array_equal(model.cum_table, oracle.gensim_cum_table(counts))
  init_vectors              Word2Vec before training == oracle.gensim_init_weights (default_rng(1))
  vectors / syn1neg         array_equal after 5 epochs, for ONE of the two sigmoid index scales
  exp_scale                 which scale matched: 83.333... (true division of the DEF constants)
                            or 83.0 (C-style folding) -- the open question of ADVICE.md round 1
  job_seeds                 the first job's next_random reproduces the first negative draws
                            (implied by vectors matching; reported separately for diagnosis)
  exp_table                 EXP_TABLE[0, 500, 999] of the oracle (gensim's is not exported; a
                            mismatch shows up in `vectors`)
  index_clamp               whether an index of 1000 (f just below 6) ever occurred in the run: the
                            one place where the oracle deliberately deviates (clamps to 999)

Output schema (JSON): {"gensim_version", "numpy_version", "corpus": {"n", "W", "L"},
"index_to_key_equal", "cum_table_equal", "init_vectors_equal",
"scales": {"83.3333": {"vectors_equal", "syn1neg_equal", "max_abs_diff"}, "83.0": {...}},
"exp_scale_matched": 83.3333 | 83.0 | null, "exp_table_probe": [..3 floats..], "pinned": bool}.
If "pinned" is true, set the matched scale as the default in oracle/gw_oracle.c AND
genewalk_b200/csrc/gw_sgns_train.cu (sigmoid_index), commit the JSON under profiles/ and drop the
"parity unpinned" notes.  This tool imports oracle/ (a checker); it is not part of the product.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def corpus(n=40, m=120, niter=5, L=10, seed=15):
    from genewalk_b200.workloads import csr_arrays
    from oracle import lib as olib
    row_ptr, col_idx = csr_arrays(n, m, seed=seed, self_loops=2)
    tok = olib.walk_uniform(row_ptr, col_idx, niter, L, 4242, 0)     # [L, W], corpus order
    return n, tok


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=os.path.join(ROOT, 'profiles', 'gensim_pin.json'))
    a = ap.parse_args()
    try:
        import gensim
        from gensim.models import Word2Vec
    except ImportError:
        print('gensim is not importable here: nothing pinned (run this on a box with gensim 4.x)')
        return 2
    from oracle import lib as olib
    n, tok = corpus()
    L, W = tok.shape
    sentences = [[str(t) for t in tok[:, p]] for p in range(W)]
    kw = dict(sg=1, vector_size=8, window=1, min_count=1, negative=5, workers=1, sample=0)
    # untrained model: vocabulary, cum_table, initial weights
    m0 = Word2Vec(**kw)
    m0.build_vocab(sentences)
    index_to_node, node_to_index, counts = olib.gensim_vocab(tok, n)
    res = {'gensim_version': gensim.__version__, 'numpy_version': np.__version__,
           'corpus': {'n': n, 'W': int(W), 'L': int(L)}}
    res['index_to_key_equal'] = [int(k) for k in m0.wv.index_to_key] == [int(x) for x in index_to_node]
    res['cum_table_equal'] = bool(np.array_equal(np.asarray(m0.cum_table, dtype=np.uint32),
                                                 olib.gensim_cum_table(counts)))
    syn0_init, _ = olib.gensim_init_weights(len(index_to_node), 8, seed=1)
    res['init_vectors_equal'] = bool(np.array_equal(m0.wv.vectors, syn0_init))
    # the reference's one-call form (deepwalk.py:135-139)
    m = Word2Vec(sentences, **kw)
    res['scales'] = {}
    res['exp_scale_matched'] = None
    for name, scale in (('83.3333', 1000.0 / 6.0 / 2.0), ('83.0', 83.0)):
        olib.set_exp_scale(scale)
        itn, syn0, syn1 = olib.word2vec_gensim_like(tok, n)
        ok0 = bool(np.array_equal(m.wv.vectors, syn0))
        ok1 = bool(np.array_equal(m.syn1neg, syn1))
        res['scales'][name] = {'vectors_equal': ok0, 'syn1neg_equal': ok1,
                               'max_abs_diff': float(np.abs(m.wv.vectors - syn0).max())}
        if ok0 and ok1:
            res['exp_scale_matched'] = float(name)
    olib.set_exp_scale(1000.0 / 6.0 / 2.0)
    et = olib.exp_table()
    res['exp_table_probe'] = [float(et[0]), float(et[500]), float(et[999])]
    res['pinned'] = bool(res['index_to_key_equal'] and res['cum_table_equal'] and
                         res['init_vectors_equal'] and res['exp_scale_matched'] is not None)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, 'w') as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))
    return 0 if res['pinned'] else 1


if __name__ == '__main__':
    sys.exit(main())
