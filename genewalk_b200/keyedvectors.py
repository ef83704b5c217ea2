"""KeyedVectors-compatible node-vector container.

The reference persists and consumes ``DW.model.wv`` (a gensim ``KeyedVectors``):
``copy.deepcopy`` + pickle (/root/reference/genewalk/cli.py:208-210,227-229), ``nv.similarity(a, b)``
(perform_statistics.py:102) and ``nv.distances(node, other_words=...)`` (null_distributions.py:46).
``NodeVectors`` offers that surface.  The arithmetic of ``similarity`` / ``distances`` runs in the
CUDA library (gw_cosine_pairs); there is no CPU implementation, so calling them without a CUDA device
raises.  Unknown keys (e.g. isolated nodes, which never enter the vocabulary) raise ``KeyError`` as
gensim does.
"""
import numpy as np

from . import _lib


class NodeVectors(object):
    def __init__(self, index_to_key, vectors, counts=None):
        self.index_to_key = list(index_to_key)
        self.key_to_index = {k: i for i, k in enumerate(self.index_to_key)}
        self.vectors = np.ascontiguousarray(vectors, dtype=np.float32)
        if self.vectors.shape[0] != len(self.index_to_key):
            raise ValueError('vectors and index_to_key disagree')
        self.vector_size = int(self.vectors.shape[1]) if self.vectors.ndim == 2 else 0
        self.counts = None if counts is None else np.asarray(counts, dtype=np.int64)
        self._dev = None

    # -- pickling: only host data crosses stage boundaries --------------------------------------
    def __getstate__(self):
        st = dict(self.__dict__)
        st['_dev'] = None
        return st

    def __len__(self):
        return len(self.index_to_key)

    def __contains__(self, key):
        return key in self.key_to_index

    def get_index(self, key):
        try:
            return self.key_to_index[key]
        except KeyError:
            raise KeyError("Key '%s' not present" % (key,))

    def get_vector(self, key, norm=False):
        v = self.vectors[self.get_index(key)]
        if norm:
            v = v / np.linalg.norm(v)
        return v

    __getitem__ = get_vector

    def get_vecattr(self, key, attr):
        if attr != 'count' or self.counts is None:
            raise KeyError(attr)
        return int(self.counts[self.get_index(key)])

    # -- device-side arithmetic -------------------------------------------------------------------
    def _device_vectors(self):
        import torch
        _lib.require_cuda()
        if self._dev is None:
            self._dev = torch.from_numpy(self.vectors).cuda()
        return self._dev

    def similarities(self, a_idx, b_idx):
        """Batched float32 cosine similarity dot(unitvec(a), unitvec(b)) for row-index arrays."""
        import torch
        vec = self._device_vectors()
        a = torch.as_tensor(np.asarray(a_idx, dtype=np.int32)).cuda()
        b = torch.as_tensor(np.asarray(b_idx, dtype=np.int32)).cuda()
        out = torch.empty(a.numel(), dtype=torch.float32, device=vec.device)
        _lib.call('gw_cosine_pairs', self.vector_size, _lib.ptr(vec), a.numel(), _lib.ptr(a),
                  _lib.ptr(b), _lib.ptr(out), _lib.stream_ptr())
        return out.cpu().numpy()

    def similarity(self, w1, w2):
        return self.similarities([self.get_index(w1)], [self.get_index(w2)])[0]

    def distances(self, word_or_vector, other_words=()):
        """1 - cosine similarity to ``other_words`` (all keys when empty, as gensim)."""
        if isinstance(word_or_vector, np.ndarray):
            raise NotImplementedError('distances() from a raw vector is not supported')
        i = self.get_index(word_or_vector)
        if not other_words:
            others = np.arange(len(self), dtype=np.int32)
        else:
            others = np.fromiter((self.get_index(w) for w in other_words), dtype=np.int32)
        sims = self.similarities(np.full(len(others), i, dtype=np.int32), others)
        return np.float32(1.0) - sims


class Word2VecModel(object):
    """What ``DeepWalk.model`` holds: ``.wv`` plus the training hyper-parameters and the output
    layer, mirroring the attributes of gensim's Word2Vec that survive into GeneWalk's pickles."""

    def __init__(self, wv, syn1neg=None, **params):
        self.wv = wv
        self.syn1neg = syn1neg
        self.__dict__.update(params)
