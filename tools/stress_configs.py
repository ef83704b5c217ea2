"""BASELINE.json configs[3] (INDRA-scale dense multigraph) and configs[4] (1M-node power-law stress):
walk + SGNS throughput on one GPU.  python tools/stress_configs.py c4|c5 [epochs]"""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from genewalk_b200.csr import GraphCSR
from genewalk_b200.deepwalk import DeepWalk
from genewalk_b200 import workloads as graphs

which = sys.argv[1]
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(3 if which == 'c4' else 4)
if which == 'c4':     # 50k nodes, 5M multi-edges, Pareto(1.8) weights: many repeated pairs, max degree ~1e4
    n, m = 50000, 5000000
    w = rng.pareto(1.8, size=n) + 1.0
else:                 # 1M nodes, ~5M edges, power-law weights (BA-like, m = 5)
    n, m = 1000000, 5000000
    w = (np.arange(1, n + 1, dtype=np.float64)) ** -0.5
w /= w.sum()
u = rng.choice(n, size=m, p=w).astype(np.int32)
v = rng.choice(n, size=m, p=w).astype(np.int32)
keep = u != v
rp, ci = graphs.csr_from_edges_numpy(n, u[keep], v[keep])
csr = GraphCSR(['v%d' % i for i in range(n)], rp, ci).to_device()
deg = np.diff(rp)
print(which, 'n', n, 'multi-edges', int(keep.sum()), 'A', csr.nnz, 'max deg', int(deg.max()), 'isolated', int((deg == 0).sum()), flush=True)
dw = DeepWalk(csr, 10, 100, seed=1)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); dw.get_walks(); e1.record(); torch.cuda.synchronize()
W = len(dw.walks)
walk_ms = e0.elapsed_time(e1)
ev = []
dv = dw.train_device(epochs=5, epoch_hook=lambda ep, when: (ev.append(torch.cuda.Event(enable_timing=True)), ev[-1].record())) if epochs >= 5 else None
if dv is None:
    # time `epochs` epochs of the 5-epoch schedule
    import ctypes
    from genewalk_b200 import _lib
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    hooks = []
    def hook(ep, when):
        e = torch.cuda.Event(enable_timing=True); e.record(); hooks.append(e)
    # run the first `epochs` epochs only by monkey-patching range of epochs: use the public knob
    dv = dw.train_device(epochs=epochs, epoch_hook=hook)
    torch.cuda.synchronize()
    ms = [hooks[2 * i].elapsed_time(hooks[2 * i + 1]) for i in range(epochs)]
else:
    torch.cuda.synchronize()
    ms = [ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(5)]
out = {'config': which, 'n': n, 'A': csr.nnz, 'walks': W, 'tokens_GB': W * 10 * 4 / 1e9,
       'walk_ms': walk_ms, 'walk_steps_per_s': W * 9 / (walk_ms / 1e3),
       'sgns_ms_per_epoch': ms, 'sgns_pairs_per_s': W * 18 / (np.mean(ms) / 1e3),
       'shape': {k: dv[k] for k in ('lanes', 'chunk_walks', 'streams')},
       'finite': bool(torch.isfinite(dv['syn0']).all().item())}
print(json.dumps(out))
