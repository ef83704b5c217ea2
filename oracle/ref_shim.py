"""Import the UNMODIFIED reference modules from /root/reference in the authoring container
(oracle; TEST INFRASTRUCTURE ONLY -- used by tests/golden/make_golden.py and by tests that are
skipped when /root/reference is absent, as it is on the GPU box).

gensim and statsmodels are not installed here, so the two names the reference imports from them
are stubbed in sys.modules before import:
  gensim.models.Word2Vec                       -> oracle restatement (oracle/lib.py)
  statsmodels.stats.multitest.fdrcorrection    -> oracle restatement (oracle/stats_ref.py)
Everything else (deepwalk.get_walks/run_single_walk, null_distributions.*, perform_statistics.GeneWalk)
then runs verbatim from /root/reference.
"""
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = '/root/reference'


class ShimKeyedVectors(object):
    """The three gensim KeyedVectors methods the reference calls, restated
    (perform_statistics.py:102, null_distributions.py:46)."""

    def __init__(self, index_to_key, vectors):
        self.index_to_key = list(index_to_key)
        self.key_to_index = {k: i for i, k in enumerate(self.index_to_key)}
        self.vectors = np.asarray(vectors, dtype=np.float32)

    def get_vector(self, key):
        return self.vectors[self.key_to_index[key]]

    __getitem__ = get_vector

    def similarity(self, w1, w2):
        def unitvec(v):
            n = np.sqrt((v * v).sum(dtype=np.float32))
            return v * np.float32(1.0 / n) if n > 0 else v
        return np.dot(unitvec(self.get_vector(w1)), unitvec(self.get_vector(w2)))

    def distances(self, word_or_vector, other_words=()):
        v = self.get_vector(word_or_vector)
        if not other_words:
            others = self.vectors
        else:
            others = self.vectors[[self.key_to_index[w] for w in other_words]]
        norm = np.linalg.norm(v)
        all_norms = np.linalg.norm(others, axis=1)
        return 1 - np.dot(others, v) / (norm * all_norms)


class ShimWord2Vec(object):
    """gensim.models.Word2Vec stand-in backed by the C oracle's gensim-faithful trainer."""

    def __init__(self, sentences=None, sg=1, vector_size=8, window=1, min_count=1, negative=5,
                 workers=1, sample=0, **kw):
        from . import lib as olib
        assert sg == 1 and window == 1 and min_count == 1 and sample == 0
        keys = {}
        for s in sentences:
            for t in s:
                if t not in keys:
                    keys[t] = len(keys)
        names = list(keys)
        L = len(sentences[0])
        tok = np.array([[keys[t] for t in s] for s in sentences], dtype=np.int32).T.copy()
        itn, syn0, _ = olib.word2vec_gensim_like(tok, len(names), d=vector_size, K=negative,
                                                 nthreads=1)
        self.wv = ShimKeyedVectors([names[i] for i in itn], syn0)


def fdrcorrection(pvals, alpha=0.05, method='indep', is_sorted=False):
    from .stats_ref import fdrcorrection as bh
    q = bh(np.asarray(pvals, dtype=np.float64))
    return q <= alpha, q


class ShimGOTerm(object):
    def __init__(self, id, name, namespace):
        self.id, self.name, self.namespace = id, name, namespace


class ShimGODag(dict):
    """goatools.obo_parser.GODag stand-in (goatools is absent here): id -> term with .id, .name,
    .namespace for the non-obsolete [Term] stanzas of an OBO file, alt_ids included -- the part of
    GODag that nx_mg_assembler.py:414-422 reads."""

    def __init__(self, obo_file, *args, **kwargs):
        super().__init__()
        block = None
        with open(obo_file) as f:
            for line in list(f) + ['[End]']:
                line = line.strip()
                if line.startswith('['):
                    if block and 'id' in block and not block.get('is_obsolete'):
                        term = ShimGOTerm(block['id'], block.get('name'), block.get('namespace'))
                        for key in [block['id']] + block.get('alt_id', []):
                            self.setdefault(key, term)
                    block = {} if line == '[Term]' else None
                elif block is not None and ': ' in line:
                    k, v = line.split(': ', 1)
                    if k == 'alt_id':
                        block.setdefault('alt_id', []).append(v)
                    elif k == 'is_obsolete':
                        block[k] = v.startswith('true')
                    else:
                        block.setdefault(k, v)


def import_assembler():
    """The reference's nx_mg_assembler module (UserNxMgAssembler) with goatools stubbed."""
    if not available():
        raise RuntimeError('/root/reference is not present (GPU box): golden fixtures are used')
    if 'goatools' not in sys.modules:
        g = types.ModuleType('goatools')
        go = types.ModuleType('goatools.obo_parser')
        go.GODag = ShimGODag
        g.obo_parser = go
        sys.modules['goatools'] = g
        sys.modules['goatools.obo_parser'] = go
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib
    return importlib.import_module('genewalk.nx_mg_assembler')


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'genewalk'))


def import_reference():
    """Returns (deepwalk, null_distributions, perform_statistics) reference modules."""
    if not available():
        raise RuntimeError('/root/reference is not present (GPU box): golden fixtures are used')
    if 'gensim' not in sys.modules:
        g = types.ModuleType('gensim')
        gm = types.ModuleType('gensim.models')
        gm.Word2Vec = ShimWord2Vec
        g.models = gm
        sys.modules['gensim'] = g
        sys.modules['gensim.models'] = gm
    if 'statsmodels' not in sys.modules:
        s = types.ModuleType('statsmodels')
        ss = types.ModuleType('statsmodels.stats')
        sm = types.ModuleType('statsmodels.stats.multitest')
        sm.fdrcorrection = fdrcorrection
        s.stats = ss
        ss.multitest = sm
        sys.modules['statsmodels'] = s
        sys.modules['statsmodels.stats'] = ss
        sys.modules['statsmodels.stats.multitest'] = sm
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib
    dw = importlib.import_module('genewalk.deepwalk')
    nd = importlib.import_module('genewalk.null_distributions')
    ps = importlib.import_module('genewalk.perform_statistics')
    return dw, nd, ps
