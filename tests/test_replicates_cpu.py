"""CPU tests of the multi-GPU host logic: replicate sharding plan and the all-gather exchange of
similarity arrays, run over gloo with world_size 2 (the GPU path uses the same code over NCCL)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from genewalk_b200 import replicates as R


def test_plan_units_covers_everything_once():
    for ng, nn in ((3, 3), (10, 10), (1, 0), (0, 2), (5, 2)):
        for ws in (1, 2, 4, 8):
            seen = []
            for rank in range(ws):
                units = R.plan_units(ng, nn, ws, rank)
                assert all(R.owner_of(k, r, ng, ws) == rank for k, r in units)
                seen += units
            assert sorted(seen) == sorted([(R.KIND_REAL, r) for r in range(ng)] +
                                          [(R.KIND_NULL, r) for r in range(nn)])
            # balanced: 20 units on 8 GPUs -> 3 waves
            sizes = [len(R.plan_units(ng, nn, ws, rank)) for rank in range(ws)]
            assert max(sizes) - min(sizes) <= 1


def test_mix_seed_streams_are_distinct():
    seeds = {R.mix_seed(b, s, r) for b in (0, 1, 2) for s in range(5) for r in range(10)}
    assert len(seeds) == 150 and all(0 <= x < 2 ** 64 for x in seeds)


def test_exchange_single_process_is_identity():
    sims = {0: torch.arange(5, dtype=torch.float32), 1: torch.ones(5)}
    null = {0: torch.tensor([0.5, 0.25]), 1: torch.zeros(0), 2: torch.tensor([1.0])}
    s, n = R.exchange_results(sims, null, 2, 3, 5, torch.device('cpu'))
    assert torch.equal(s[0], sims[0]) and torch.equal(s[1], sims[1])
    assert [x.tolist() for x in n] == [[0.5, 0.25], [], [1.0]]


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, ng, nn, m, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=ws)
    try:
        # each rank fabricates the results of its own units: values encode (kind, rep)
        local_sims, local_null = {}, {}
        for kind, rep in R.plan_units(ng, nn, ws, rank):
            if kind == R.KIND_REAL:
                local_sims[rep] = torch.full((m,), float(rep)) + torch.arange(m) * 1e-3
            else:
                local_null[rep] = torch.full((10 + 3 * rep,), 100.0 + rep)
        sims, nulls = R.exchange_results(local_sims, local_null, ng, nn, m, torch.device('cpu'))
        ok = sims.shape == (ng, m)
        for rep in range(ng):
            ok &= bool(torch.allclose(sims[rep], torch.full((m,), float(rep)) + torch.arange(m) * 1e-3))
        for rep in range(nn):
            ok &= nulls[rep].numel() == 10 + 3 * rep and bool((nulls[rep] == 100.0 + rep).all())
        parts = R.gather_varlen(torch.arange(rank + 1, dtype=torch.float32))
        ok &= [p.numel() for p in parts] == list(range(1, ws + 1))
        out[rank] = int(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('ng,nn', [(3, 3), (1, 4), (5, 0)])
def test_exchange_over_gloo_world_size_2(ng, nn):
    ws = 2
    ctx = mp.get_context('spawn')
    out = ctx.Array('i', [0] * ws)
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, ng, nn, 7, out)) for r in range(ws)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert list(out) == [1] * ws


def test_rolling_schedule_order_and_shares(monkeypatch):
    """_run_rolling: unit i runs on stream i % nstreams, unit i - nstreams is finished right before
    unit i is enqueued (no barrier between rounds), units finish in order, and the last, smaller
    round is told how many units share the GPU.  CUDA streams are replaced by dummies."""
    import contextlib
    import torch
    from genewalk_b200 import replicates as R

    class FakeStream(object):
        def __init__(self, name):
            self.name, self.waited = name, []

        def wait_stream(self, other):
            self.waited.append(other.name)
    cur = FakeStream('cur')
    side = [FakeStream('s%d' % k) for k in range(8)]
    active = []

    @contextlib.contextmanager
    def fake_ctx(stream):
        active.append(stream.name)
        yield
        active.pop()
    monkeypatch.setattr(torch.cuda, 'current_stream', lambda: cur)
    monkeypatch.setattr(torch.cuda, 'stream', fake_ctx)
    monkeypatch.setattr(R, '_side_streams', lambda count: side[:count])
    log = []
    R._run_rolling(8, 3, lambda i: log.append(('enq', i, active[-1])) or ('job', i),
                   lambda i, job: log.append(('fin', i, job, active[-1])))
    want = [('enq', 0, 's0'), ('enq', 1, 's1'), ('enq', 2, 's2'),
            ('fin', 0, ('job', 0), 's0'), ('enq', 3, 's0'),
            ('fin', 1, ('job', 1), 's1'), ('enq', 4, 's1'),
            ('fin', 2, ('job', 2), 's2'), ('enq', 5, 's2'),
            ('fin', 3, ('job', 3), 's0'), ('enq', 6, 's0'),
            ('fin', 4, ('job', 4), 's1'), ('enq', 7, 's1'),
            ('fin', 5, ('job', 5), 's2'), ('fin', 6, ('job', 6), 's0'), ('fin', 7, ('job', 7), 's1')]
    assert log == want
    assert all(s.waited == ['cur'] for s in side[:3]) and cur.waited == ['s0', 's1', 's2']
    # a single stream degenerates to the serial loop on the current stream
    log.clear()
    R._run_rolling(3, 1, lambda i: log.append(('enq', i, active[-1])) or i,
                   lambda i, job: log.append(('fin', i, job, active[-1])))
    assert log == [('enq', 0, 'cur'), ('fin', 0, 0, 'cur'), ('enq', 1, 'cur'),
                   ('fin', 1, 1, 'cur'), ('enq', 2, 'cur'), ('fin', 2, 2, 'cur')]
    # more streams than units
    log.clear()
    R._run_rolling(2, 6, lambda i: log.append(('enq', i)) or i,
                   lambda i, job: log.append(('fin', i)))
    assert log == [('enq', 0), ('enq', 1), ('fin', 0), ('fin', 1)]


def test_fit_concurrent_caps_by_free_memory(monkeypatch):
    """fit_concurrent: as many replicates in flight as asked for -- walks are never stored, a
    replicate holds its tables, vocabulary and plan (64 MiB allowance + n * (8 d + 64) bytes) --
    unless even that exceeds 80 % of the free HBM; never below one."""
    import torch
    from genewalk_b200 import replicates as R
    state = {'free': 170 << 30}
    monkeypatch.setattr(torch.cuda, 'mem_get_info', lambda: (state['free'], 180 << 30))
    assert R.fit_concurrent(24, 38000) == 24
    assert R.fit_concurrent(24, 1000000, 128) == 24
    state['free'] = 1 << 30
    assert R.fit_concurrent(24, 38000) == 11
    state['free'] = 1 << 20
    assert R.fit_concurrent(24, 38000) == 1
    assert R.fit_concurrent(1, 10) == 1
