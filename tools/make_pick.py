"""Row-popularity sample of the C2 workload for tools/l2_mixbench.cu: 2^22 row ids, 3/13 drawn ~ degree
(centre and context rows of a step) and 10/13 ~ degree^0.75 (negatives), shuffled."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from tests import graphs
n, u, v, is_go = graphs.human_scale_edges(seed=2)
rp, ci = graphs.csr_from_edges_numpy(n, u, v)
deg = np.diff(rp).astype(np.float64)
rng = np.random.default_rng(0)
N = 1 << 22
f = deg / deg.sum(); fn = f ** 0.75; fn /= fn.sum()
k = N * 3 // 13
pick = np.concatenate([rng.choice(n, size=k, p=f), rng.choice(n, size=N - k, p=fn)]).astype(np.uint32)
rng.shuffle(pick)
pick.tofile(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/pick_c2.bin')
print('top row share', np.bincount(pick, minlength=n).max() / N)
