#!/usr/bin/env python
"""bench.py -- GeneWalk hot-path benchmark (contract: see the task prompt / DESIGN.md section 6).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--config c2|c4|c5]

BASELINE.json's metric is "end-to-end GeneWalk s (nreps=10) at 1/2/4/8 B200; walk steps/s; SGNS
pairs/s".  A "step" is the representation-learning work of ONE GeneWalk run at the metric's setting
on the human-scale synthetic gene-GO network (BASELINE.json configs[1]/[2]: ~20k genes + ~18k GO
terms, ~1M edges; reference defaults walk_length=10, niter=100, vector_size=8, window=1, negative=5,
epochs=5): nreps_graph = 10 real-network replicates (cli.py:200-214) and nreps_null = 10
randomized-network replicates (cli.py:219-237) -- per replicate the vocabulary pass over its
W = 100*A walks, the 5 SGNS epochs over them (5 * 18 * W pairs; the walks are regenerated inside the
trainer, there is no stored corpus) and its similarity output (gene-GO cosines / null similarities).
The 20 replicates are independent; each is enqueued on its own CUDA stream and they train
concurrently, every one with the worker count its network calls for (gw_sgns_auto_shape: never a
function of GPU fill or GPU count).

  value     SGNS (centre, context) pairs per second of whole-step time; the CSR is resident in HBM
            and the results stay on the device
  e2e       the same metric through the public API with HOST inputs and outputs:
            run_genewalk(nx.MultiGraph, genes, nreps 10/10) -> results DataFrame (CSR extraction,
            H2D, all replicates, D2H of similarities, statistics)
  pipeline  BASELINE's first metric: seconds of node_vectors + null_distribution + statistics at
            nreps 10/10 with the 20 replicates SHARDED over the N ranks (strong scaling), plus a
            checksum of the pooled null distribution that must agree on all ranks (it exercises the
            variable-length NCCL exchange)
  N > 1     value / e2e: every rank runs its own step (weak scaling, no data-path collective)
  --impl reference   the reference's CPU path for the same step restated in C (oracle/, gensim-
            faithful SGNS, OpenMP Hogwild on all host cores) on a bounded sample of the same corpus
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')   # one hardware queue per replicate stream

NITER, WL, D, K, EPOCHS = 100, 10, 8, 5, 5
PAIRS_PER_WALK = 2 * (WL - 1) * EPOCHS
BYTES_PER_PAIR = 4 * D * (2 * K + 4) + 8 + 8 * K      # 496 at d=8, K=5
BYTES_PER_STEP = 16                                     # walk step
METRIC = 'SGNS pairs/s over the replicates of a GeneWalk run (nreps_graph = nreps_null = 10)'


# ---------------------------------------------------------------------------------------------
class ClockMonitor(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons every 200 ms while the timed region runs."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._halt = index, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                      '--format=csv,noheader,nounits'], capture_output=True,
                                     text=True, timeout=5).stdout.strip()
                f = [x.strip() for x in out.split(',')]
                if len(f) >= 7:
                    self.samples.append(f)
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm = [float(s[0]) for s in self.samples if s[0].replace('.', '').isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[i] for s in self.samples for i in range(4) if s[3 + i] == 'Active'})
        return {'sm_mhz': float(np.median(sm)) if sm else None,
                'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
                'samples': len(self.samples)}


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def build_inputs(args):
    """(config name, n, u, v, is_go | None, n_genes | None)."""
    from genewalk_b200 import workloads
    if args.config == 'c4':
        return (workloads,) + workloads.indra_scale_edges(seed=3) + (None, None)
    if args.config == 'c5':
        return (workloads,) + workloads.power_law_stress_edges(seed=4) + (None, None)
    s = args.scale
    ng, ngo = int(20000 * s), int(18000 * s)
    n, u, v, is_go = workloads.human_scale_edges(seed=2, n_genes=ng, n_go=ngo,
                                                 n_gene_edges=int(700000 * s), n_annot=int(250000 * s),
                                                 n_isa=int(50000 * s), n_modules=max(4, int(500 * s)))
    return workloads, n, u, v, is_go, ng


WORKLOADS = {
    'c2': 'human-scale synthetic gene-GO network (BASELINE.json configs[1]/[2]): one GeneWalk run, '
          '%d real + %d null node_vectors replicates in flight on the GPU',
    'c4': 'INDRA-scale dense synthetic multigraph (BASELINE.json configs[3]: 50k nodes, 5M multi-edges, '
          'Pareto degrees): %d real + %d null node_vectors replicates in flight on the GPU',
    'c5': '1M-node power-law stress graph (BASELINE.json configs[4]): %d real + %d null node_vectors '
          'replicates in flight on the GPU'}


def workload_config(args, n, n_edges, A):
    return {'workload': WORKLOADS[args.config] % (args.real, args.null), 'config': args.config,
            'nodes': n, 'edges': n_edges, 'adjacency_entries': A, 'walks_per_real_replicate': NITER * A,
            'real_replicates_per_step': args.real, 'null_replicates_per_step': args.null,
            'walk_length': WL, 'niter': NITER, 'vector_size': D, 'window': 1, 'negative': K,
            'epochs': EPOCHS, 'scale': args.scale,
            'l2': 'L2 is flushed between timed steps (a 512 MB buffer is written); every step rebuilds '
                  'its %d table pairs from scratch and draws new walks (new seeds).  Within a step the '
                  'tables are L2-resident BY DESIGN (that residency is what the kernel is built on, '
                  'not a carry-over between iterations)' % (args.real + args.null)}


# ---------------------------------------------------------------------------------------------
def cpu_sample_rate(row_ptr, col_idx, threads, target_s=15.0):
    """Oracle (gensim-faithful C restatement, OpenMP) on a bounded sample of the workload's corpus:
    the first `m` walks of the reference's corpus order, all 5 epochs.  Returns (pairs/s, text)."""
    from oracle import lib as olib
    n = len(row_ptr) - 1
    W = NITER * int(row_ptr[-1])

    def run(m):
        t0 = time.perf_counter()
        tok = olib.walk_uniform(row_ptr, col_idx, NITER, WL, 12345, 0, 0, m)
        olib.word2vec_gensim_like(tok, n, d=D, K=K, epochs=EPOCHS, nthreads=threads)
        return time.perf_counter() - t0
    pilot = min(W, 20000 * max(1, threads))
    dt = run(pilot)
    rate = pilot * PAIRS_PER_WALK / dt
    m = int(min(W, max(pilot, rate * target_s / PAIRS_PER_WALK)))
    if m > pilot:
        dt = run(m)
    else:
        m = pilot
    pairs = m * PAIRS_PER_WALK
    return pairs / dt, ('first %d of %d walks of one replicate\'s corpus (reference order), %d epochs = '
                        '%.1f M pairs in %.1f s' % (m, W, EPOCHS, pairs / 1e6, dt))


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    workloads, n, u, v, is_go, ng = build_inputs(args)
    row_ptr, col_idx = workloads.csr_from_edges_numpy(n, u, v)
    threads = host_threads()   # torchrun exports OMP_NUM_THREADS=1: ask the OS instead
    vals, sample = [], ''
    t_all = time.perf_counter()
    for i in range(args.warmup + args.steps):
        rate, sample = cpu_sample_rate(row_ptr, col_idx, threads,
                                       target_s=min(15.0, 120.0 / max(1, args.warmup + args.steps)))
        if i >= args.warmup:
            vals.append(rate)
    value = float(np.mean(vals))
    A = int(row_ptr[-1])
    # a null replicate's corpus is ~1.28x a real one's; the extrapolation uses the real size for all
    pairs_step = (args.real + args.null) * NITER * A * PAIRS_PER_WALK
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'pairs/s',
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': 1e3 * pairs_step / value, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args, n, len(u), A),
            'cpu_baseline': {'value': value, 'unit': 'pairs/s', 'cores': threads, 'kind': 'port',
                             'sample': sample},
            'e2e': {'value': value, 'unit': 'pairs/s', 'h2d_bytes_per_step': 0,
                    'd2h_bytes_per_step': 0},
            'note': 'ms_per_step is the full-step time extrapolated linearly from the sample; '
                    'gensim is not installable here, the C oracle restates it (oracle/gw_oracle.c)',
            'wall_s': time.perf_counter() - t_all}
    emit(line)


# ---------------------------------------------------------------------------------------------
def live_l2_peak(rows):
    """Build tools/l2_mixbench.cu with nvcc and measure, on this GPU and in this run, the rate of the
    SGNS step's access pattern (13 gathers + 14 reductions of 32-byte rows per step, groups of 4
    lanes) over `rows` uniformly drawn rows -- the L2 request-rate ceiling for that mix.  Returns
    (steps/s, gathers-only steps/s, source text) or (None, None, why)."""
    exe = os.path.join(ROOT, 'gpurun_out', 'l2_mixbench_r02')
    try:
        os.makedirs(os.path.dirname(exe), exist_ok=True)
        nvcc = '/usr/local/cuda/bin/nvcc' if os.path.exists('/usr/local/cuda/bin/nvcc') else 'nvcc'
        subprocess.run([nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-o', exe,
                        os.path.join(ROOT, 'tools', 'l2_mixbench.cu')], check=True, capture_output=True,
                       timeout=300)
        out = subprocess.run([exe, '--peak', str(int(rows))], check=True, capture_output=True, text=True,
                             timeout=300).stdout
        recs = [json.loads(x) for x in out.splitlines() if x.startswith('{')]
        mix = max(r['steps_per_s'] for r in recs if r['reds'] == 14)
        gat = max(r['steps_per_s'] for r in recs if r['reds'] == 0)
        return mix, gat, 'measured in this run (tools/l2_mixbench.cu --peak %d)' % rows
    except Exception as exc:   # no nvcc on this box, or the tool failed
        return None, None, 'live measurement failed (%s)' % (str(exc)[:80],)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from genewalk_b200 import _lib
    from genewalk_b200.csr import GraphCSR
    from genewalk_b200.replicates import (_NullJob, _run_rolling, fit_concurrent, preload_kernels,
                                          real_replicate, run_genewalk)
    from genewalk_b200.perform_statistics import GeneWalk

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    _lib.require_cuda()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    workloads, n, u, v, is_go, ng = build_inputs(args)
    mg = genes = None
    if is_go is not None:
        mg, genes = workloads.edges_to_networkx(n, u, v, is_go, ng)
        csr = GraphCSR.from_networkx(mg).to_device()
        from genewalk_b200.csr import multigraph_degrees
        degrees = multigraph_degrees(mg)
        gw = GeneWalk(mg, genes, [], np.zeros(1, np.float32), gene_id_type='custom')
        pairs = gw.collect_pairs()
        d_pg = torch.from_numpy(pairs['pair_gene'].astype(np.int32)).cuda()
        d_po = torch.from_numpy(pairs['pair_go'].astype(np.int32)).cuda()
    else:   # C4 / C5: no GO attributes; the replicate's product is its table
        rp, ci = workloads.csr_from_edges_numpy(n, u, v)
        csr = GraphCSR(['v%d' % i for i in range(n)], rp, ci).to_device()
        degrees = workloads.multigraph_degrees_from_edges(n, u, v)
        d_pg = d_po = None
    m_pairs = int(d_po.numel()) if d_po is not None else 0
    A = csr.nnz
    R_real, R_null = args.real, args.null
    R = R_real + R_null
    w2v = dict(lanes=args.lanes) if args.lanes else {}
    if args.reductions != 'per_pair':
        w2v['reductions'] = args.reductions
    preload_kernels(D, **w2v)
    in_flight = fit_concurrent(R, n, D, NITER * A, args.lanes)
    train_events, count_events = [], []
    l2_flush = torch.empty(512 << 20, dtype=torch.uint8, device='cuda')

    def device_step(step, timed):
        """R independent replicates, one per CUDA stream; returns the walks trained."""
        walks = []
        base = (1 + step) * 1000003 + rank

        def hook_for(unit):
            def hook(stage, when):
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                (count_events if stage == 'count' else train_events).append((unit, stage, when, e))
            return hook if timed else None

        def enqueue(i):   # longest first: the null replicates (1.28x the walks), then the real ones
            if i >= R_null:
                dw = real_replicate(csr, base, i - R_null, D, wait=False, epoch_hook=hook_for(i), **w2v)
                if m_pairs:   # the replicate's product in the pipeline: gene-GO similarities
                    sims = torch.empty(m_pairs, dtype=torch.float32, device='cuda')
                    _lib.call('gw_cosine_pairs', D, _lib.ptr(dw._pending['syn0']), m_pairs, _lib.ptr(d_pg),
                              _lib.ptr(d_po), _lib.ptr(sims), _lib.stream_ptr())
                walks.append(dw._pending['W'])
                return dw
            job = _NullJob(None, base, i, D, WL, NITER, dict(epoch_hook=hook_for(i), **w2v),
                           degrees=degrees)
            walks.append(len(job.dw.walks))
            return job

        def finish(i, job):
            if i < R_null:
                job.finish()           # the null similarities of the replicate (device)
            else:
                job._pending = None
        _run_rolling(R, in_flight, enqueue, finish)
        return walks, (1 + step)

    # ---- value: device-resident ------------------------------------------------------------
    for i in range(args.warmup):
        device_step(i, False)
    barrier()
    mon = ClockMonitor(local)
    mon.start()
    launches0 = _lib.launch_count()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    walks_timed = 0
    for i in range(args.steps):
        l2_flush.fill_(i)          # nothing of the previous step stays in L2
        wk, _ = device_step(args.warmup + i, True)
        walks_timed += sum(wk)
    t1.record()
    barrier()
    launches = _lib.launch_count() - launches0
    clocks = mon.stop()
    ms_max = max_over_ranks(t0.elapsed_time(t1))
    pairs_timed_all = sum_over_ranks(walks_timed * PAIRS_PER_WALK)
    value = pairs_timed_all / (ms_max / 1e3)
    # dominant kernel: one SGNS epoch launch.  The launches of the replicates of a step overlap on
    # the GPU, so the rate is (launches' pairs) / (time during which at least one SGNS launch was in
    # flight), from the per-launch CUDA events on the launching streams.
    # (events of different steps share (unit, stage) keys: they are paired in recording order)
    spans, k_ms, open_ev = [], [], {}
    for uid, st, when, e in train_events:
        if when == 'before':
            open_ev[(uid, st)] = e
        else:
            b = open_ev.pop((uid, st))
            k_ms.append(b.elapsed_time(e))
            spans.append((t0.elapsed_time(b), t0.elapsed_time(e)))
    if os.environ.get('GW_BENCH_SPANS'):   # diagnosis: (unit, stage, begin ms, end ms) of every launch
        rec, op = [], {}
        for uid, st, when, e in count_events + train_events:
            if when == 'before':
                op[(uid, st)] = e
            else:
                rec.append([uid, st, t0.elapsed_time(op.pop((uid, st))), t0.elapsed_time(e)])
        with open(os.environ['GW_BENCH_SPANS'], 'w') as f:
            json.dump(rec, f)
    spans.sort()
    busy_ms, hi = 0.0, -1.0
    for lo, up in spans:
        if up > hi:
            busy_ms += up - max(lo, hi)
            hi = up
    n_launch = len(spans)
    k_avg = float(np.mean(k_ms)) if k_ms else float('nan')
    k_eff = busy_ms / max(n_launch, 1)            # GPU time per launch with the overlap credited
    pairs_launch_avg = walks_timed * 2 * (WL - 1) / max(n_launch / EPOCHS, 1)
    achieved = BYTES_PER_PAIR * walks_timed * PAIRS_PER_WALK / (busy_ms / 1e3) / 1e9
    c_open, c_ms = {}, []
    for uid, st, when, e in count_events:
        if when == 'before':
            c_open[uid] = e
        else:
            c_ms.append(c_open.pop(uid).elapsed_time(e))
    peaks = {}
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    traffic, traffic_src = None, None
    try:
        if args.config != 'c2' or args.scale != 1.0:
            raise ValueError('the capture is of the c2 workload')
        with open(os.path.join(ROOT, 'profiles', 'r02_sgns_train_traffic.json')) as f:
            tj = json.load(f)
            traffic, traffic_src = tj.get('dram_bytes_per_launch'), tj.get('source')
    except Exception:
        pass
    dv_shape = None

    # ---- pipeline (strong scaling) and e2e (public API, host inputs) -------------------------
    pipeline = e2e = None
    if mg is not None and not args.no_pipeline:
        run_genewalk(mg, genes, nreps_graph=1, nreps_null=1, alpha_fdr=0.1, gene_id_type='custom',
                     base_seed=99, shard=False, **w2v)        # host-side warm-up of the public path
        barrier()
        p0 = time.perf_counter()
        df, parts = run_genewalk(mg, genes, nreps_graph=args.real, nreps_null=args.null, alpha_fdr=0.1,
                                 gene_id_type='custom', base_seed=2024, return_parts=True, **w2v)
        torch.cuda.synchronize()
        local_s = time.perf_counter() - p0
        barrier()
        p_s = max_over_ranks(local_s)
        p_pairs = sum_over_ranks(parts['walks_trained_local'] * PAIRS_PER_WALK)
        null = parts['null']
        chk = float(np.float64(null.astype(np.float64).sum())) if len(null) else 0.0
        chk_t = torch.tensor([chk, float(len(null)), float(len(df))], dtype=torch.float64, device='cuda')
        chks = [torch.empty_like(chk_t) for _ in range(world)]
        if world > 1:
            dist.all_gather(chks, chk_t)
        else:
            chks = [chk_t]
        same = all(bool(torch.equal(c, chks[0])) for c in chks)
        pipeline = {'nreps_graph': args.real, 'nreps_null': args.null, 'seconds': p_s,
                    'scaling': 'strong', 'pairs': p_pairs, 'pairs_per_s': p_pairs / p_s,
                    'rows_significant': int(len(df)), 'alpha_fdr': 0.1,
                    'pooled_null': {'n': int(len(null)), 'sum': chk, 'equal_on_all_ranks': bool(same)},
                    'timings_rank0': parts['timings'],
                    'timeline_rank0_s': [[u[0], u[1], k, round(b / 1e3, 1), round(e / 1e3, 1)]
                                         for u, k, b, e in parts['timeline']],
                    'limiter': 'replicates are independent and run concurrently; a training uses the '
                               'worker count its network calls for (fidelity), so with fewer replicates '
                               'per GPU the per-worker latency, not the GPU, bounds the time'}
        # e2e: weak scaling like `value` -- every rank runs a whole GeneWalk run of its own
        # CSR + (gene, GO) index arrays + gensim's initial rows (uploaded once, shared by all
        # replicates) + the degree sequence of every randomization
        h2d = csr.row_ptr.nbytes + csr.col_idx.nbytes + 2 * 4 * m_pairs + n * D * 4 + R_null * 4 * n
        if world == 1:
            e_s, e_pairs, d2h = p_s, p_pairs, int(parts['sims'].numel() * 4 + null.nbytes)
        else:
            barrier()
            e0 = time.perf_counter()
            df2, parts2 = run_genewalk(mg, genes, nreps_graph=args.real, nreps_null=args.null,
                                       alpha_fdr=0.1, gene_id_type='custom', base_seed=3000 + rank,
                                       return_parts=True, shard=False, **w2v)
            torch.cuda.synchronize()
            local_e = time.perf_counter() - e0
            barrier()
            e_s = max_over_ranks(local_e)
            e_pairs = sum_over_ranks(parts2['walks_trained_local'] * PAIRS_PER_WALK)
            d2h = int(parts2['sims'].numel() * 4 + parts2['null'].nbytes)
        e2e = {'value': e_pairs / e_s, 'unit': 'pairs/s', 'h2d_bytes_per_step': int(h2d),
               'd2h_bytes_per_step': int(d2h), 'ms_per_step': 1e3 * e_s,
               'what': 'run_genewalk(nx.MultiGraph, genes, nreps_graph=%d, nreps_null=%d) -> DataFrame, '
                       'one whole run per rank' % (args.real, args.null)}
    elif mg is None:
        # C4 / C5: no GeneWalk table to make; e2e = the same step from HOST CSR arrays to host vectors
        from genewalk_b200.replicates import run_walks_many
        host_csr = GraphCSR(csr.nodes, csr.row_ptr, csr.col_idx)
        barrier()
        e0 = time.perf_counter()
        dws = run_walks_many(host_csr, R, base_seed=555 + rank, **w2v)
        torch.cuda.synchronize()
        local_e = time.perf_counter() - e0
        barrier()
        e_s = max_over_ranks(local_e)
        e_pairs = sum_over_ranks(sum(len(dw.walks) for dw in dws) * PAIRS_PER_WALK)
        e2e = {'value': e_pairs / e_s, 'unit': 'pairs/s',
               'h2d_bytes_per_step': int(csr.row_ptr.nbytes + csr.col_idx.nbytes + n * D * 4),
               'd2h_bytes_per_step': int(sum(2 * dw.model.wv.vectors.nbytes for dw in dws)),
               'ms_per_step': 1e3 * e_s,
               'what': 'run_walks_many(host CSR, %d replicates) -> host vectors' % R}
        del dws

    line = None
    if rank == 0:
        import ctypes
        a_l, a_u, a_p = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int32(0)
        _lib.call('gw_sgns_auto_shape', n, NITER * A, 1000, NITER * int(csr.max_degree or 0),
                  ctypes.byref(a_l), ctypes.byref(a_u), ctypes.byref(a_p))
        mix, gat, src = live_l2_peak(2 * in_flight * n) if not args.no_l2_peak else (None, None, 'skipped')
        steps_per_s = walks_timed * PAIRS_PER_WALK / 2 / (busy_ms / 1e3)   # step = 2 pairs (13 gathers + 14 reductions)
        if mix is None:
            try:
                with open(os.path.join(ROOT, 'profiles', 'l2_pattern_peaks.json')) as f:
                    mix = json.load(f)['uniform_rows']
                    src += '; fallback: profiles/l2_pattern_peaks.json (round 1, one 38k-row table)'
            except Exception:
                pass
        peak_gbs = mix * 2 * BYTES_PER_PAIR / 1e9 if mix else None
        line = {'metric': METRIC, 'value': value, 'unit': 'pairs/s', 'n_gpus': world,
                'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_max / args.steps,
                'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
                'data': 'synthetic', 'config': workload_config(args, n, len(u), A),
                'clocks': clocks, 'e2e': e2e, 'gpu_launches': int(launches),
                'roofline': {'kernel': 'sgns_train_kernel<8,4,5> (one epoch of one replicate per launch)',
                             'bound': 'l2', 'achieved': achieved, 'peak': peak_gbs, 'unit': 'GB/s',
                             'frac': achieved / peak_gbs if peak_gbs else None, 'traffic': traffic,
                             'traffic_source': traffic_src,
                             'peak_source': src,
                             'peak_pattern_steps_per_s': mix, 'peak_gathers_only_steps_per_s': gat,
                             'achieved_steps_per_s': steps_per_s,
                             'frac_of_hbm_copy_bw': achieved / hbm_peak,
                             'hbm_peak_gbs': hbm_peak,
                             'algorithmic_bytes_per_pair': BYTES_PER_PAIR,
                             'pairs_per_launch': pairs_launch_avg, 'avg_launch_ms': k_eff,
                             'launches': n_launch, 'replicates_in_flight': in_flight,
                             'launch_ms_on_its_stream': k_avg, 'sgns_busy_ms': busy_ms,
                             'note': 'the embedding tables are L2-resident by design, so the 496 B/pair '
                                     'are L2 sector traffic (HBM carries almost nothing: see traffic); '
                                     'peak = 2*496 B x the measured rate of the step\'s own request mix '
                                     '(13 row gathers + 14 row reductions) over uniformly drawn rows; '
                                     'avg_launch_ms = (time with >= 1 SGNS launch in flight) / launches '
                                     'because the launches of a step overlap'},
                'walk': {'kernel': 'walk_count_smem_kernel (vocabulary pass: every walk drawn, counted, '
                                   'not stored)',
                         'steps_per_s': (NITER * A * (WL - 1) / (float(np.mean(c_ms)) / 1e3)) if c_ms else None,
                         'ms_on_its_stream': float(np.mean(c_ms)) if c_ms else None,
                         'note': 'timed on its stream while the other replicates of the step train; the '
                                 'trainer draws the same walks again in every epoch inside the SGNS '
                                 'kernel (12 algorithmic bytes per step, nothing stored)'},
                'sgns_kernel_pairs_per_s': walks_timed * PAIRS_PER_WALK / (busy_ms / 1e3),
                'sgns_shape': {'lanes': a_l.value if not args.lanes else args.lanes,
                               'unit_walks': a_u.value, 'policy': ('chunk', 'burst')[a_p.value],
                               'job_walks': 1000, 'per_replicate': True}}
        if pipeline is not None:
            line['pipeline'] = pipeline
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        csr.to_host()
        threads = host_threads()
        rate, sample = cpu_sample_rate(csr.row_ptr, csr.col_idx, threads)
        line['cpu_baseline'] = {'value': rate, 'unit': 'pairs/s', 'cores': threads, 'kind': 'port',
                                'sample': sample}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def quiet_stdout():
    """Libraries (NCCL's version banner, torchrun) print to stdout; the contract is ONE JSON line
    there.  Everything written to fd 1 from here on goes to stderr, emit() writes the line."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    text = json.dumps(line) + '\n'
    if _REAL_STDOUT is None:
        sys.stdout.write(text)
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, text.encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=2)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='c2', choices=['c2', 'c4', 'c5'],
                    help='c2: the headline workload (default); c4 / c5: BASELINE.json configs[3] / [4]')
    ap.add_argument('--scale', type=float, default=1.0, help='shrink the c2 workload (tests only)')
    ap.add_argument('--lanes', type=int, default=0, help='workers per training (0: library default)')
    ap.add_argument('--reductions', default='per_pair', choices=['per_pair', 'merged'],
                    help='reduction order of the trainer (library default: per_pair, gensim\'s order)')
    ap.add_argument('--real', type=int, default=None, help='real-network replicates per step')
    ap.add_argument('--null', type=int, default=None, help='randomized-network replicates per step')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-pipeline', action='store_true')
    ap.add_argument('--no-l2-peak', action='store_true')
    args = ap.parse_args()
    default_reps = {'c2': (10, 10), 'c4': (6, 6), 'c5': (2, 2)}[args.config]
    if args.real is None:
        args.real = default_reps[0]
    if args.null is None:
        args.null = default_reps[1]
    if args.warmup < 3 and args.impl == 'ours' and args.scale == 1.0:
        print('warning: fewer than 3 warm-up steps', file=sys.stderr)
    quiet_stdout()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
