#!/usr/bin/env python
"""bench.py -- GeneWalk hot-path benchmark (contract: see the task prompt / DESIGN.md section 6).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--scale S]

A "step" is one batch of 6 node_vectors replicates (cli.py:200-214; flag --replicates) -- the batch
the replicate driver trains at once per GPU; a default GeneWalk run holds 3 + 3 independent
trainings, the nreps=10 run of BASELINE's metric 10 + 10 -- on the human-scale synthetic gene-GO network (BASELINE.json configs[1]: ~20k genes +
~18k GO terms, ~1M edges; reference defaults walk_length=10, niter=100, vector_size=8, window=1,
negative=5, epochs=5): per replicate, random walks from every node (W = 100*A walks) followed by
the 5 SGNS epochs over them (5 * 18 * W pairs).  The replicates of a step are independent; they are
enqueued on separate CUDA streams and train concurrently on the GPU.
metric = SGNS (centre, context) pairs trained per second of whole-stage time (walk generation,
vocabulary, alias table and init included).

  value   inputs (CSR) already resident in HBM, results left on the device
  e2e     the same batch through the public API run_walks_many(nx.MultiGraph, 6): CSR extraction
          from networkx, host->device copies, training, device->host copy of the vectors
  N > 1   every rank runs its own replicates (the path shards by replicate); one NCCL
          all-gather of the per-replicate gene-GO similarity vectors closes each step (weak scaling)
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

NITER, WL, D, K, EPOCHS = 100, 10, 8, 5, 5
BYTES_PER_PAIR = 4 * D * (2 * K + 4) + 8 + 8 * K      # 496 at d=8, K=5
BYTES_PER_STEP = 16                                     # walk step
METRIC = 'SGNS pairs/s over the node_vectors stage (walks + 5 SGNS epochs per replicate)'


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


def build_inputs(scale):
    from genewalk_b200 import workloads as graphs
    ng, ngo = int(20000 * scale), int(18000 * scale)
    n, u, v, is_go = graphs.human_scale_edges(seed=2, n_genes=ng, n_go=ngo,
                                              n_gene_edges=int(700000 * scale),
                                              n_annot=int(250000 * scale), n_isa=int(50000 * scale),
                                              n_modules=max(4, int(500 * scale)))
    return graphs, n, u, v, is_go, ng


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
    rate = pilot * 2 * (WL - 1) * EPOCHS / dt
    m = int(min(W, max(pilot, rate * target_s / (2 * (WL - 1) * EPOCHS))))
    if m > pilot:
        dt = run(m)
    else:
        m = pilot
    pairs = m * 2 * (WL - 1) * EPOCHS
    return pairs / dt, ('first %d of %d walks of the corpus (reference order), %d epochs = %.1f M '
                        'pairs in %.1f s' % (m, W, EPOCHS, pairs / 1e6, dt))


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    graphs, n, u, v, is_go, ng = build_inputs(args.scale)
    row_ptr, col_idx = graphs.csr_from_edges_numpy(n, u, v)
    from oracle import lib as olib
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
    pairs_step = args.replicates * NITER * A * 2 * (WL - 1) * EPOCHS
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'pairs/s',
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': 1e3 * pairs_step / value, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(n, len(u), A, args.scale, args.replicates),
            'cpu_baseline': {'value': value, 'unit': 'pairs/s', 'cores': threads, 'kind': 'port',
                             'sample': sample},
            'e2e': {'value': value, 'unit': 'pairs/s', 'h2d_bytes_per_step': 0,
                    'd2h_bytes_per_step': 0},
            'note': 'ms_per_step is the full-stage time extrapolated linearly from the sample; '
                    'gensim is not installable here, the C oracle restates it (oracle/gw_oracle.c)',
            'wall_s': time.perf_counter() - t_all}
    emit(line)


def workload_config(n, n_edges, A, scale, reps):
    return {'workload': 'human-scale synthetic gene-GO network, batch of %d node_vectors replicates '
                        'trained concurrently (BASELINE.json configs[1])' % reps,
            'nodes': n, 'edges': n_edges, 'adjacency_entries': A, 'walks_per_replicate': NITER * A,
            'replicates_per_step': reps,
            'walk_length': WL, 'niter': NITER, 'vector_size': D, 'window': 1, 'negative': K,
            'epochs': EPOCHS, 'pairs_per_step': reps * NITER * A * 2 * (WL - 1) * EPOCHS,
            'scale': scale,
            'l2': 'token streams (%.1f GB per replicate) >> 126 MB L2; embedding tables are rebuilt '
                  'every step' % (NITER * A * WL * 4 / 1e9)}


# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from genewalk_b200 import _lib
    from genewalk_b200.csr import GraphCSR
    from genewalk_b200.deepwalk import DeepWalk
    from genewalk_b200.replicates import run_walks_many
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

    graphs, n, u, v, is_go, ng = build_inputs(args.scale)
    mg, genes = graphs.edges_to_networkx(n, u, v, is_go, ng)
    csr = GraphCSR.from_networkx(mg).to_device()
    A = csr.nnz
    W = NITER * A
    R = max(1, int(args.replicates))
    pairs_step = R * W * 2 * (WL - 1) * EPOCHS
    gw = GeneWalk(mg, genes, [], np.zeros(1, np.float32), gene_id_type='custom')
    pairs = gw.collect_pairs()
    d_pg = torch.from_numpy(pairs['pair_gene'].astype(np.int32)).cuda()
    d_po = torch.from_numpy(pairs['pair_go'].astype(np.int32)).cuda()
    m_pairs = int(d_po.numel())
    gathered = [torch.empty((R, m_pairs), dtype=torch.float32, device='cuda') for _ in range(world)]
    sims_all = torch.empty((R, m_pairs), dtype=torch.float32, device='cuda')
    cur = torch.cuda.current_stream()
    side = [torch.cuda.Stream() for _ in range(R)] if R > 1 else [cur]
    train_events, walk_events = [], []

    def device_step(step, timed):
        """R independent replicates, one per CUDA stream, then the exchange of their products."""
        dvs = []
        for k in range(R):
            if R > 1:
                side[k].wait_stream(cur)
            with torch.cuda.stream(side[k]):
                def hook(ep, when):
                    e = torch.cuda.Event(enable_timing=True)
                    e.record()
                    train_events.append((len(dvs), ep, when, e))
                seed = ((1 + step) * 1000003 + rank) * 16 + k
                dw = DeepWalk(csr, WL, NITER, seed=seed)
                w0 = torch.cuda.Event(enable_timing=True); w1 = torch.cuda.Event(enable_timing=True)
                w0.record()
                dw.get_walks()
                w1.record()
                if timed:
                    walk_events.append((w0, w1))
                dv = dw.train_device(vector_size=D, negative=K, epochs=EPOCHS, lanes=args.lanes,
                                     concurrent=R, epoch_hook=hook if timed else None)
                # the replicate's product in the pipeline: gene-GO similarities (fixed length)
                _lib.call('gw_cosine_pairs', D, _lib.ptr(dv['syn0']), m_pairs, _lib.ptr(d_pg),
                          _lib.ptr(d_po), _lib.ptr(sims_all[k]), _lib.stream_ptr())
                dvs.append(dv)
        if R > 1:
            for k in range(R):
                cur.wait_stream(side[k])
        if world > 1:
            dist.all_gather(gathered, sims_all)
        return dvs

    # ---- value: device-resident ------------------------------------------------------------
    for i in range(args.warmup):
        device_step(i, False)
    barrier()
    mon = ClockMonitor(local)
    mon.start()
    launches0 = _lib.launch_count()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(args.steps):
        dv = device_step(args.warmup + i, True)[0]
    t1.record()
    barrier()
    launches = _lib.launch_count() - launches0
    clocks = mon.stop()
    ms = t0.elapsed_time(t1)
    t_ms = torch.tensor([ms], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    value = world * pairs_step * args.steps / (ms_max / 1e3)
    # dominant kernel: one SGNS epoch launch.  With R > 1 the launches of different replicates
    # overlap on the GPU, so the rate is (launches' pairs) / (time during which at least one SGNS
    # launch was in flight), from the per-launch CUDA events on the launching streams.
    befores = [e for _, ep, when, e in train_events if when == 'before']
    afters = [e for _, ep, when, e in train_events if when == 'after']
    k_ms = [b.elapsed_time(a) for b, a in zip(befores, afters)]
    k_avg = float(np.mean(k_ms)) if k_ms else float('nan')
    spans = sorted((t0.elapsed_time(b), t0.elapsed_time(a)) for b, a in zip(befores, afters))
    busy_ms, hi = 0.0, -1.0
    for lo, up in spans:
        if up > hi:
            busy_ms += up - max(lo, hi)
            hi = up
    n_launch = len(spans)
    walk_ms = float(np.mean([a.elapsed_time(b) for a, b in walk_events]))
    pairs_launch = W * 2 * (WL - 1)
    k_eff = busy_ms / max(n_launch, 1)            # GPU time per launch with the overlap credited
    achieved = BYTES_PER_PAIR * pairs_launch / (k_eff / 1e3) / 1e9
    peaks = {}
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    traffic = None
    try:
        with open(os.path.join(ROOT, 'profiles', 'sgns_train_traffic.json')) as f:
            traffic = json.load(f).get('dram_bytes_per_launch')
    except Exception:
        pass

    l2_peaks = {}
    try:
        with open(os.path.join(ROOT, 'profiles', 'l2_pattern_peaks.json')) as f:
            l2_peaks = json.load(f)
    except Exception:
        pass

    # ---- e2e: public API with host inputs ---------------------------------------------------
    barrier()
    d2h = 0
    h2d = csr.row_ptr.nbytes + csr.col_idx.nbytes + R * n * D * 4
    for i in range(min(args.warmup, 1)):
        run_walks_many(mg, R, base_seed=7 + i, concurrent=R, lanes=args.lanes)
    barrier()
    e0 = time.perf_counter()
    for i in range(args.steps):
        dws = run_walks_many(mg, R, base_seed=100 + i * world + rank, concurrent=R, lanes=args.lanes)
        d2h = 0
        for dw in dws:
            wv = dw.model.wv
            d2h += (wv.vectors.nbytes + dw.model.syn1neg.nbytes + wv.counts.nbytes +
                    wv.node_ids.nbytes)
        del dws
    barrier()
    e_s = time.perf_counter() - e0
    t_e = torch.tensor([e_s], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = world * pairs_step * args.steps / float(t_e.item())

    line = None
    if rank == 0:
        line = {'metric': METRIC, 'value': value, 'unit': 'pairs/s', 'n_gpus': world,
                'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_max / args.steps,
                'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
                'data': 'synthetic', 'config': workload_config(n, len(u), A, args.scale, R),
                'clocks': clocks,
                'e2e': {'value': e2e_value, 'unit': 'pairs/s', 'h2d_bytes_per_step': int(h2d),
                        'd2h_bytes_per_step': int(d2h), 'ms_per_step': 1e3 * float(t_e.item()) / args.steps},
                'gpu_launches': int(launches),
                'roofline': {'kernel': 'sgns_train_kernel<8,4,5> (one epoch per launch)',
                             'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                             'frac': achieved / peak, 'traffic': traffic,
                             'peak_source': 'measured (MEASURED_PEAKS.json hbm_gbs)' if peaks else 'fallback',
                             'algorithmic_bytes_per_pair': BYTES_PER_PAIR,
                             'pairs_per_launch': pairs_launch, 'avg_launch_ms': k_eff,
                             'launches': n_launch, 'concurrent_launches': R,
                             'launch_ms_on_its_stream': k_avg, 'sgns_busy_ms': busy_ms,
                             'note': 'avg_launch_ms = (time with >= 1 SGNS launch in flight) / '
                                     'launches: the launches of the R replicates of a step overlap. '
                                     'Tables are L2-resident by design: the algorithmic bytes are '
                                     'L2 sector traffic; HBM carries only the token streams'},
                'roofline_l2': {
                    'what': 'SGNS steps/s (2 pairs) vs the measured rate of the same gather+reduce '
                            'pattern in isolation (tools/l2_mixbench.cu, profiles/l2_pattern_peaks.json)',
                    'achieved_steps_per_s': pairs_launch / 2 / (k_eff / 1e3),
                    'peak_uniform_rows': l2_peaks.get('uniform_rows'),
                    'peak_c2_row_popularity': l2_peaks.get('c2_popularity'),
                    'frac_of_uniform': (pairs_launch / 2 / (k_eff / 1e3) / l2_peaks['uniform_rows'])
                    if l2_peaks.get('uniform_rows') else None},
                'walk': {'steps_per_s': W * (WL - 1) / (walk_ms / 1e3), 'ms': walk_ms,
                         'achieved_gbs': BYTES_PER_STEP * W * (WL - 1) / (walk_ms / 1e3) / 1e9,
                         'frac_of_l2_random_sector_rate':
                             (W * (WL - 1) / (walk_ms / 1e3) / l2_peaks['random_sector_gathers_per_s'])
                             if l2_peaks.get('random_sector_gathers_per_s') else None,
                         'note': 'timed on its stream while the other replicates of the step train; '
                                 'one random L2 sector (the neighbour) per step is the bound'},
                'sgns_kernel_pairs_per_s': pairs_launch / (k_eff / 1e3),
                'sgns_shape': {'lanes': dv['lanes'], 'chunk_walks': dv['chunk_walks'], 'streams': dv['streams'],
                               'job_walks': dv['job_walks']}}
    if args.pipeline > 0:
        # BASELINE metric, first item: end-to-end seconds of node_vectors + null_distribution +
        # statistics (cli.py:194-252) with nreps_graph = nreps_null = --pipeline, replicates sharded
        from genewalk_b200.replicates import run_genewalk
        barrier()
        p0 = time.perf_counter()
        df, parts = run_genewalk(mg, genes, nreps_graph=args.pipeline, nreps_null=args.pipeline,
                                 alpha_fdr=0.1, gene_id_type='custom', base_seed=2024,
                                 lanes=args.lanes, concurrent=R, return_parts=True)
        barrier()
        t_p = torch.tensor([time.perf_counter() - p0], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(t_p, op=dist.ReduceOp.MAX)
        if rank == 0:
            line['pipeline'] = {'nreps_graph': args.pipeline, 'nreps_null': args.pipeline,
                                'seconds': float(t_p.item()), 'rows_significant': int(len(df)),
                                'alpha_fdr': 0.1, 'timings_rank0': parts['timings'],
                                'timeline_rank0_ms': [[u[0], u[1], k, round(b), round(e)]
                                                      for u, k, b, e in parts['timeline']]}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import lib as olib
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
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--scale', type=float, default=1.0, help='shrink the workload (tests only)')
    ap.add_argument('--lanes', type=int, default=0, dest='lanes')
    ap.add_argument('--replicates', type=int, default=6,
                    help='replicates per step, trained concurrently on one GPU')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--pipeline', type=int, default=0,
                    help='also time the full pipeline with this many real and null replicates')
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == 'ours' and args.scale == 1.0:
        print('warning: fewer than 3 warm-up steps', file=sys.stderr)
    quiet_stdout()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
