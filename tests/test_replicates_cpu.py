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


def test_rolling_schedule_order(monkeypatch):
    """_run_rolling: at most nstreams units in flight, one per stream; a stream whose unit has
    completed on the device gets the next unit (no barrier between rounds, no waiting on a busy
    stream while another is idle).  CUDA streams and events are replaced by dummies; a fake clock
    decides which units are complete."""
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
    done = set()

    class FakeEvent(object):
        unit = None

        def record(self):
            self.unit = current_unit[0]

        def query(self):
            return self.unit in done

    current_unit = [None]

    @contextlib.contextmanager
    def fake_ctx(stream):
        active.append(stream.name)
        yield
        active.pop()
    monkeypatch.setattr(torch.cuda, 'current_stream', lambda: cur)
    monkeypatch.setattr(torch.cuda, 'stream', fake_ctx)
    monkeypatch.setattr(R, '_side_streams', lambda count: side[:count])
    monkeypatch.setattr(R, '_new_event', lambda timing=False: FakeEvent())
    log = []
    # unit durations: unit 1 is slow, so stream s1 is skipped over while s0 and s2 take new units
    finish_order = iter([0, 2, 3, 4, 1, 5, 6, 7])

    def enqueue(i):
        current_unit[0] = i
        log.append(('enq', i, active[-1]))
        return ('job', i)

    def fake_sleep(dt):
        done.add(next(finish_order))
    import time
    monkeypatch.setattr(time, 'sleep', fake_sleep)
    R._run_rolling(8, 3, enqueue, lambda i, job: log.append(('fin', i, job, active[-1])))
    want = [('enq', 0, 's0'), ('enq', 1, 's1'), ('enq', 2, 's2'),
            ('fin', 0, ('job', 0), 's0'), ('enq', 3, 's0'),
            ('fin', 2, ('job', 2), 's2'), ('enq', 4, 's2'),
            ('fin', 3, ('job', 3), 's0'), ('enq', 5, 's0'),
            ('fin', 4, ('job', 4), 's2'), ('enq', 6, 's2'),
            ('fin', 1, ('job', 1), 's1'), ('enq', 7, 's1'),
            ('fin', 5, ('job', 5), 's0'), ('fin', 6, ('job', 6), 's2'), ('fin', 7, ('job', 7), 's1')]
    assert log == want
    assert all(s.waited == ['cur'] for s in side[:3]) and cur.waited == ['s0', 's1', 's2']
    # a single stream degenerates to the serial loop on the current stream
    log.clear()
    done.clear()
    finish_order = iter([0, 1, 2])
    R._run_rolling(3, 1, enqueue, lambda i, job: log.append(('fin', i, job, active[-1])))
    assert log == [('enq', 0, 'cur'), ('fin', 0, ('job', 0), 'cur'), ('enq', 1, 'cur'),
                   ('fin', 1, ('job', 1), 'cur'), ('enq', 2, 'cur'), ('fin', 2, ('job', 2), 'cur')]
    # more streams than units
    log.clear()
    done.clear()
    finish_order = iter([0, 1])
    R._run_rolling(2, 6, enqueue, lambda i, job: log.append(('fin', i)))
    assert log == [('enq', 0, 's0'), ('enq', 1, 's1'), ('fin', 0), ('fin', 1)]


def test_fit_concurrent_caps_by_residency_and_memory(monkeypatch):
    """fit_concurrent: as many replicates in flight as asked for, but never more trainings than are
    resident on the GPU with their full worker count (so a replicate's result cannot depend on what
    runs next to it), and never more than fit 80 % of the free HBM (walks are never stored: a
    replicate holds its tables, vocabulary and plan, 64 MiB allowance + n * (8 d + 64) bytes)."""
    import torch
    from genewalk_b200 import _lib
    from genewalk_b200 import replicates as R

    class FakeLib(object):
        @staticmethod
        def gw_sgns_resident_trainings(d, lanes):
            return max(1, 148 * 512 // (lanes * (d // 2)))
    monkeypatch.setattr(_lib, 'load', lambda: FakeLib)

    def fake_call(name, n, W, job, hub, lanes, unit, policy):
        assert name == 'gw_sgns_auto_shape'
        lanes._obj.value = max(16, min(1536, n // 24))
    monkeypatch.setattr(_lib, 'call', fake_call)
    state = {'free': 170 << 30}
    monkeypatch.setattr(torch.cuda, 'mem_get_info', lambda: (state['free'], 180 << 30))
    assert R.fit_concurrent(24, 38000) == 148 * 128 // 1536          # 12 trainings of 1536 workers
    assert R.fit_concurrent(24, 2300) == 24                            # 95 workers each: all fit
    assert R.fit_concurrent(24, 38000, lanes=512) == 24
    assert R.fit_concurrent(24, 1000000, 128) == 1                     # 1536 workers of 32 threads
    state['free'] = 1 << 30
    assert R.fit_concurrent(24, 2300) == 12
    state['free'] = 1 << 20
    assert R.fit_concurrent(24, 2300) == 1
    assert R.fit_concurrent(1, 10) == 1
