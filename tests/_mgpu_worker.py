"""Worker of tests/test_multigpu.py: one process per GPU (torchrun), NCCL."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from genewalk_b200.replicates import run_genewalk, plan_units   # noqa: E402
from genewalk_b200 import workloads as graphs                                         # noqa: E402


def main():
    out = sys.argv[1]
    local = int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    rank, ws = dist.get_rank(), dist.get_world_size()
    mg, genes = graphs.modular_gene_go_network(300, 60, 8, 1500, 900, 100, seed=5)
    df, parts = run_genewalk(mg, genes, nreps_graph=3, nreps_null=3, alpha_fdr=1,
                             gene_id_type='custom', base_seed=11, return_parts=True)
    mine = plan_units(3, 3, ws, rank)
    # the same run in sequential SGNS mode (bit-reproducible), sharded over the ranks and, on every
    # rank, unsharded: the number of GPUs must not change a single bit, nor the shape of a training
    kw = dict(nreps_graph=2, nreps_null=2, alpha_fdr=1, gene_id_type='custom', base_seed=5, niter=8,
              return_parts=True, lanes=1)
    _, sh = run_genewalk(mg, genes, **kw)
    _, un = run_genewalk(mg, genes, shard=False, **kw)
    _, auto = run_genewalk(mg, genes, nreps_graph=2, nreps_null=2, alpha_fdr=1, gene_id_type='custom',
                           base_seed=5, niter=8, return_parts=True, shard=False)
    np.savez(os.path.join(out, 'rank%d.npz' % rank), sims=parts['sims'].cpu().numpy(),
             null=parts['null'], padj=df['gene_padj'].to_numpy(dtype=np.float64),
             n_units=len(mine), n_nvs=len(parts['nvs']),
             seq_equal=bool(np.array_equal(sh['sims'].cpu().numpy(), un['sims'].cpu().numpy()) and
                            np.array_equal(sh['null'], un['null'])),
             shapes_sharded=np.array([list(x[:2]) for x in parts['train_shapes_local']]),
             shapes_unsharded=np.array([list(x[:2]) for x in auto['train_shapes_local']]))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
