"""Two-GPU run of the replicate driver over NCCL (skipped with fewer than 2 GPUs): every rank ends
with the same similarity matrix, pooled null distribution and results table."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_two_gpu_pipeline(tmp_path):
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', '29533',
           os.path.join(ROOT, 'tests', '_mgpu_worker.py'), str(tmp_path)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    r0 = np.load(os.path.join(tmp_path, 'rank0.npz'))
    r1 = np.load(os.path.join(tmp_path, 'rank1.npz'))
    assert int(r0['n_units']) == 3 and int(r1['n_units']) == 3          # 6 units over 2 ranks
    assert int(r0['n_nvs']) + int(r1['n_nvs']) == 3                      # real replicates split 2 + 1
    assert r0['sims'].shape[0] == 3 and np.array_equal(r0['sims'], r1['sims'])
    assert np.array_equal(r0['null'], r1['null']) and np.all(np.diff(r0['null']) >= 0)
    assert np.array_equal(r0['padj'], r1['padj'], equal_nan=True)
    assert np.isfinite(r0['sims']).all() and len(r0['null']) > 3 * 3000
    # sequential SGNS mode: the 2-GPU run equals the 1-GPU run bit for bit
    assert bool(r0['seq_equal']) and bool(r1['seq_equal'])
    # default mode: every training has the same shape (workers, unit) on 1 and on 2 GPUs
    shapes = {tuple(x) for r in (r0, r1) for k in ('shapes_sharded', 'shapes_unsharded') for x in r[k]}
    assert len(shapes) == 1, shapes
