"""Aggregate SGNS throughput with R replicates trained concurrently on one GPU (R CUDA streams)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from genewalk_b200.csr import GraphCSR
from genewalk_b200.deepwalk import DeepWalk
from tests import graphs
R = int(sys.argv[1]); lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n, u, v, is_go = graphs.human_scale_edges(seed=2)
rp, ci = graphs.csr_from_edges_numpy(n, u, v)
csr = GraphCSR(['v%d' % i for i in range(n)], rp, ci).to_device()
dws = []
for r in range(R):
    dw = DeepWalk(csr, 10, 100, seed=r + 1); dw.get_walks(); dws.append(dw)
torch.cuda.synchronize()
streams = [torch.cuda.Stream() for _ in range(R)]
W = len(dws[0].walks)
for trial in range(2):
    t0 = time.perf_counter()
    outs = []
    for r in range(R):
        with torch.cuda.stream(streams[r]):
            outs.append(dws[r].train_device(epochs=1, lanes=lanes, shared_sm=R > 1))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print('R=%d lanes=%s: %.3f s for %d epoch(s) -> %.3g pairs/s aggregate' % (R, outs[0]['lanes'], dt, R, R * W * 18 / dt), flush=True)
