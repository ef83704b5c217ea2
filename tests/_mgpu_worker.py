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
    np.savez(os.path.join(out, 'rank%d.npz' % rank), sims=parts['sims'].cpu().numpy(),
             null=parts['null'], padj=df['gene_padj'].to_numpy(dtype=np.float64),
             n_units=len(mine), n_nvs=len(parts['nvs']))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
