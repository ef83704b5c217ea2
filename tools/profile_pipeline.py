"""Wall-clock breakdown of replicates.run_genewalk on the human-scale network (host + device)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from genewalk_b200 import workloads as graphs
from genewalk_b200.csr import GraphCSR
from genewalk_b200.perform_statistics import GeneWalk
from genewalk_b200 import replicates as R

nreps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
t = time.perf_counter
n, u, v, is_go = graphs.human_scale_edges(seed=2)
t0 = t(); mg, genes = graphs.edges_to_networkx(n, u, v, is_go, 20000); print('build nx graph %.2fs' % (t() - t0))
t0 = t(); csr = GraphCSR.from_networkx(mg).to_device(); torch.cuda.synchronize(); print('CSR from networkx %.2fs' % (t() - t0))
gw = GeneWalk(mg, genes, [], np.zeros(1, np.float32), gene_id_type='custom')
t0 = t(); pairs = gw.collect_pairs(); print('collect_pairs %.2fs  (m=%d)' % (t() - t0, len(pairs['pair_go'])))
sims, nulls = [], []
for r in range(nreps):
    t0 = t(); dw = R.real_replicate(csr, 1, r); torch.cuda.synchronize(); t1 = t()
    s = gw.pair_similarities(dw.model.wv, pairs); torch.cuda.synchronize(); t2 = t()
    print('real replicate %d: run_walks %.2fs, pair_similarities %.2fs' % (r, t1 - t0, t2 - t1)); sims.append(s)
for r in range(nreps):
    t0 = t(); nl, _ = R.null_replicate(mg, 1, r); torch.cuda.synchronize(); print('null replicate %d: %.2fs' % (r, t() - t0)); nulls.append(nl)
t0 = t(); pooled = torch.cat(nulls).contiguous()
from genewalk_b200 import _lib
_lib.call('gw_sort_f32', _lib.ptr(pooled), pooled.numel(), _lib.stream_ptr()); torch.cuda.synchronize(); print('pool+sort null %.2fs (%d values)' % (t() - t0, pooled.numel()))
gw.srd = pooled.cpu().numpy(); gw._d_srd = pooled; gw.nvs = [None] * nreps
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
t0 = t(); df = gw.generate_output(alpha_fdr=0.1, sims=torch.stack(sims), pairs=pairs); print('generate_output %.2fs (%d rows)' % (t() - t0, len(df)))
pr.disable(); pstats.Stats(pr).sort_stats('cumulative').print_stats(12)
