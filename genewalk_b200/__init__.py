"""genewalk_b200 -- B200-native (sm_100a CUDA) implementation of GeneWalk's representation-learning
hot path behind genewalk's own Python API.

Drop-in modules (same names, signatures and error behaviour as /root/reference/genewalk):
  genewalk_b200.deepwalk            DeepWalk, run_walks, get_start_nodes, run_walks_for_node, ...
  genewalk_b200.null_distributions  get_rand_graph, get_null_distributions
  genewalk_b200.perform_statistics  GeneWalk
All arithmetic runs in libgenewalk_b200.so (C ABI: include/genewalk_b200.h); no CPU fallback.
"""
import logging

__version__ = '0.1.0'

logging.getLogger('genewalk').addHandler(logging.NullHandler())

from . import _lib                                    # noqa: E402,F401
from .deepwalk import DeepWalk, run_walks             # noqa: E402,F401
from .null_distributions import get_rand_graph, get_null_distributions   # noqa: E402,F401
from .perform_statistics import GeneWalk              # noqa: E402,F401
from .keyedvectors import NodeVectors                 # noqa: E402,F401
