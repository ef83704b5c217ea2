"""Aggregate SGNS throughput with R replicates trained concurrently on one GPU (R CUDA streams).

    python tools/concurrent_replicates.py [--sorted|--shuffled] R:lanes[:reductions] [R:lanes[:reductions] ...]

--sorted relabels the nodes by descending degree (the labelling of a get_rand_graph network),
--shuffled by a random permutation; default is the generator's labelling.  --null trains on
configuration-model randomizations of the network (one per replicate) instead.
"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from genewalk_b200.csr import GraphCSR
from genewalk_b200.deepwalk import DeepWalk
from genewalk_b200 import workloads as graphs
n, u, v, is_go = graphs.human_scale_edges(seed=2)
mode = [a for a in sys.argv[1:] if a.startswith('--')]
sys.argv = [a for a in sys.argv if not a.startswith('--')]
if mode and mode[0] != '--null':
    deg = np.bincount(np.concatenate([u, v]), minlength=n)
    order = np.argsort(-deg, kind='stable') if mode[0] == '--sorted' else np.random.default_rng(0).permutation(n)
    new_id = np.empty(n, np.int64); new_id[order] = np.arange(n)
    u, v = new_id[u], new_id[v]
    print('labelling:', mode[0], flush=True)
rp, ci = graphs.csr_from_edges_numpy(n, u, v)
csr = GraphCSR(['v%d' % i for i in range(n)], rp, ci).to_device()
RMAX = max(int(a.split(':')[0]) for a in sys.argv[1:])
dws = []
degrees = np.bincount(np.concatenate([u, v]), minlength=n)
for r in range(RMAX):
    g = csr
    if mode and mode[0] == '--null':
        import networkx as nx
        from genewalk_b200.null_distributions import get_rand_graph
        g = get_rand_graph(nx.MultiGraph(), seed=100 + r, materialize=False, degrees=degrees)
    dw = DeepWalk(g, 10, 100, seed=r + 1); dw.get_walks(); dws.append(dw)
torch.cuda.synchronize()
streams = [torch.cuda.Stream() for _ in range(RMAX)]
W = len(dws[0].walks)
for spec in sys.argv[1:]:
    f = spec.split(':')
    R, lanes = int(f[0]), int(f[1])
    red = f[2] if len(f) > 2 else 'auto'
    for trial in range(2):
        t0 = time.perf_counter()
        outs = []
        for r in range(R):
            with torch.cuda.stream(streams[r]):
                outs.append(dws[r].train_device(epochs=2, lanes=lanes, concurrent=R, reductions=red))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    nrm = float(outs[0]['syn0'].norm(dim=1).max())
    print('R=%d lanes=%s streams=%s red=%s: %.3f s -> %.3g pairs/s aggregate (max row norm %.2f)' %
          (R, outs[0]['lanes'], outs[0]['streams'], red, dt, R * 2 * W * 18 / dt, nrm), flush=True)
    del outs
