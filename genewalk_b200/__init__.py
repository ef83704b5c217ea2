"""genewalk_b200 -- B200-native (sm_100a CUDA) implementation of GeneWalk's representation-learning
hot path behind genewalk's own Python API.

Drop-in modules (same names, signatures and error behaviour as /root/reference/genewalk):
  genewalk_b200.deepwalk            DeepWalk, run_walks, get_start_nodes, run_walks_for_node, ...
  genewalk_b200.null_distributions  get_rand_graph, get_null_distributions
  genewalk_b200.perform_statistics  GeneWalk
All arithmetic runs in libgenewalk_b200.so (C ABI: include/genewalk_b200.h); no CPU fallback.
"""
import logging
import os

__version__ = '0.2.0'

# The replicate driver keeps up to 24 independent trainings in flight on as many CUDA streams.  A
# CUDA context has 8 hardware work queues by default; streams beyond that share queues, and a
# seconds-long training kernel then holds back the kernels of the streams queued behind it
# (measured: 12 streams ran as 8 + 4).  32 is the maximum; it must be set before the CUDA context
# is created, so import this package before the first CUDA call of the process (or export it).
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')

logging.getLogger('genewalk').addHandler(logging.NullHandler())

from . import _lib                                    # noqa: E402,F401
from .deepwalk import DeepWalk, run_walks             # noqa: E402,F401
from .null_distributions import get_rand_graph, get_null_distributions   # noqa: E402,F401
from .perform_statistics import GeneWalk              # noqa: E402,F401
from .keyedvectors import NodeVectors                 # noqa: E402,F401
