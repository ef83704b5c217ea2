import sys, os, ctypes, struct
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from genewalk_b200 import _lib
from genewalk_b200.deepwalk import DeepWalk
from genewalk_b200.null_distributions import get_rand_graph
import tools.parity_study as ps
lanes, chunk, streams, tries = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
mg, genes = ps.build_case('c1mod', 2)
g = get_rand_graph(mg, seed=ps.mix(2, 2, 0), materialize=False)
lib = _lib.load()
lib.gw_debug_read.argtypes = [ctypes.POINTER(ctypes.c_uint64)]
def f32(u): return struct.unpack('f', struct.pack('I', u & 0xffffffff))[0]
for t in range(tries):
    dw = DeepWalk(g, 10, 100, seed=ps.mix(2, 3, 0) + t); dw.get_walks()
    buf = (ctypes.c_uint64 * 16)()
    lib.gw_debug_read(buf)
    recs = []
    def hook(ep, when):
        if when == 'after':
            lib.gw_debug_read(buf)
            if buf[0]: recs.append((ep, [int(x) for x in buf[:8]]))
    dv = dw.train_device(lanes=lanes, chunk_walks=chunk, streams=streams, epochs=5, epoch_hook=hook)
    torch.cuda.synchronize()
    cnt = dv['counts']
    print('try', t, 'max', float(dv['syn0'].abs().max()), float(dv['syn1neg'].abs().max()))
    for ep, b in recs[:2]:
        print('  first record in epoch', ep, 'site', b[1], 'centre', b[2], 'cnt', int(cnt[b[2]]), 'target', b[3], 'cnt', int(cnt[b[3]]),
              'vals', f32(b[4]), f32(b[5]), f32(b[6]), 'thread', b[7])
