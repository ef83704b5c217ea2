"""Track table norms epoch by epoch (diagnostic for Hogwild stability)."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from genewalk_b200 import _lib
from genewalk_b200.csr import GraphCSR
from genewalk_b200.deepwalk import DeepWalk
from genewalk_b200.null_distributions import get_rand_graph
from tests import graphs
import tools.parity_study as ps

case, lanes, chunk, streams = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
tries = int(sys.argv[5]) if len(sys.argv) > 5 else 3
K = int(sys.argv[6]) if len(sys.argv) > 6 else 5
if case == 'c2':
    n, u, v, is_go = graphs.human_scale_edges(seed=2)
    rp, ci = graphs.csr_from_edges_numpy(n, u, v)
    g = GraphCSR(['v%d' % i for i in range(n)], rp, ci).to_device()
else:
    mg, genes = ps.build_case('c1mod', 2)
    g = get_rand_graph(mg, seed=ps.mix(2, 2, 0), materialize=False)
for t in range(tries):
    dw = DeepWalk(g, 10, 100, seed=ps.mix(2, 3, 0) + t)
    dw.get_walks()
    stats = []
    def hook(ep, when):
        pass
    import ctypes
    # run epoch by epoch by calling train_device internals
    nodes, n, tokens = dw._device_corpus()
    dv = None
    for ep_end in range(1, 6):
        pass
    dv = dw.train_device(lanes=lanes, chunk_walks=chunk, streams=streams, epochs=5, negative=K,
                         epoch_hook=lambda ep, when: stats.append((ep, when, None)))
    torch.cuda.synchronize()
    s0, s1 = dv['syn0'], dv['syn1neg']
    n0 = s0.norm(dim=1); n1 = s1.norm(dim=1)
    fin = bool(torch.isfinite(s0).all() and torch.isfinite(s1).all())
    cnt = dv['counts'].float()
    top = torch.argsort(n1, descending=True)[:5]
    print('try', t, 'finite', fin, 'max|syn0| %.3g max|syn1| %.3g  mean %.3g %.3g' % (n0.max().item(), n1.max().item(), n0.mean().item(), n1.mean().item()),
          'top syn1 rows (node,count,norm):', [(int(i), int(cnt[i]), round(float(n1[i]), 2)) for i in top], flush=True)
