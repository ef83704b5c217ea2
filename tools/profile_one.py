"""One node_vectors replicate of the bench workload (BASELINE.json configs[1]) with the shipped
defaults, cut down to what an ncu capture needs: CSR upload, the vocabulary pass (walk_count
kernel), and `epochs` SGNS epoch launches (sgns_train_kernel, walks regenerated in-kernel).

  python tools/profile_one.py [epochs=1] [null]        ('null': train on a randomized network)
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch                                         # noqa: E402
from genewalk_b200 import workloads                  # noqa: E402
from genewalk_b200.csr import GraphCSR               # noqa: E402
from genewalk_b200.deepwalk import DeepWalk          # noqa: E402

epochs = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n, u, v, is_go = workloads.human_scale_edges(seed=2)
if len(sys.argv) > 2 and sys.argv[2] == 'null':
    from genewalk_b200.null_distributions import get_rand_graph
    graph = get_rand_graph(None, seed=5, materialize=False,
                           degrees=workloads.multigraph_degrees_from_edges(n, u, v))
else:
    rp, ci = workloads.csr_from_edges_numpy(n, u, v)
    graph = GraphCSR(['v%d' % i for i in range(n)], rp, ci).to_device()
dw = DeepWalk(graph, 10, 100, seed=77)
dw.get_walks()
ev = []


def hook(stage, when):
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    ev.append((stage, when, e))


t0 = time.time()
dv = dw.train_device(epochs=5, epoch_hook=hook) if epochs >= 5 else dw.train_device(epochs=epochs, epoch_hook=hook)
torch.cuda.synchronize()
W = dv['W']
for i in range(0, len(ev), 2):
    ms = ev[i][2].elapsed_time(ev[i + 1][2])
    if ev[i][0] == 'count':
        print('vocabulary pass: %.1f ms  (%.3g walk steps/s)' % (ms, W * 9 / ms * 1e3))
    else:
        print('epoch %s: %.1f ms  (%.3g pairs/s, %d workers, units of %d walks)' % (
            ev[i][0], ms, W * 18 / ms * 1e3, dv['lanes'], dv['unit_walks']))
print('finite:', bool(torch.isfinite(dv['syn0']).all()), ' wall %.1fs' % (time.time() - t0))
