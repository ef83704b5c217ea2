"""Throughput of the SGNS kernel vs the number of concurrent workers (lanes) on the human-scale
synthetic network: one epoch per setting, CUDA-event timed.  python tools/sgns_sweep.py [scale]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from genewalk_b200 import _lib                     # noqa: E402
from genewalk_b200.csr import GraphCSR             # noqa: E402
from genewalk_b200.deepwalk import DeepWalk        # noqa: E402
from genewalk_b200 import workloads as graphs                           # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
lanes_list = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else \
    [2048, 4096, 9472, 18944]
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 0
stride = int(sys.argv[4]) if len(sys.argv) > 4 else 0
streams = int(sys.argv[5]) if len(sys.argv) > 5 else 0
n, u, v, is_go = graphs.human_scale_edges(seed=2, n_genes=int(20000 * scale), n_go=int(18000 * scale),
                                          n_gene_edges=int(700000 * scale), n_annot=int(250000 * scale),
                                          n_isa=int(50000 * scale), n_modules=max(4, int(500 * scale)))
rp, ci = graphs.csr_from_edges_numpy(n, u, v)
csr = GraphCSR(['v%d' % i for i in range(n)], rp, ci).to_device()
dw = DeepWalk(csr, 10, 100, seed=1)
from genewalk_b200.csr import coprime_stride
dw.get_walks(stride=coprime_stride(csr.nnz) if stride < 0 else stride)
W = len(dw.walks)
pairs = W * 18
out = {'n': n, 'A': csr.nnz, 'W': W, 'pairs_per_epoch': pairs, 'lanes': {}}
for lanes in lanes_list:
    ev = []

    def hook(ep, when):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        ev.append(e)
    dv = dw.train_device(epochs=5, lanes=lanes, chunk_walks=chunk, streams=streams, epoch_hook=hook)
    torch.cuda.synchronize()
    ms = [ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(5)]
    ok = bool(torch.isfinite(dv['syn0']).all().item())
    out['lanes'][lanes] = {'ms_per_epoch': ms, 'pairs_per_s': pairs / (np.mean(ms[1:]) / 1e3),
                           'finite': ok}
    print(lanes, dv['chunk_walks'], dv['streams'], ['%.1f' % m for m in ms], '%.3g pairs/s' % (pairs / (np.mean(ms[1:]) / 1e3)), ok,
          flush=True)
print(json.dumps(out))
