#!/usr/bin/env python
"""bench.py -- DIGAT dual-graph encoder: pairs ("impressions") scored per second on synthetic MIND-shaped data.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run, one rank per GPU)
  python bench.py --impl reference ...                     (the reference's CPU algorithm = oracle port, host cores)

A step = one pass of the hot path over one batch of `--batch` (user behaviour, candidate news) pairs of the
synthetic MIND-small-dev-sized corpus (BASELINE.json configs[1]).
  value : pairs/s with the corpus resident in HBM (pair = (behaviour index, news id); gathers + encoder + logits)
  e2e   : pairs/s through scoring.score_host_batches with HOST (pinned) per-pair tensors, H2D + D2H inside the timing
          (both drivers stage batch k+1 on a side stream while batch k is encoded)
One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D = 400
WORKLOADS = {
    # name: (SAG_neighbors, SAG_hops, graph_depth, n_news, n_behaviors, mean_candidates)
    'mind_small_dev_n3_L3': (3, 2, 3, 65238, 73152, 37.5),       # BASELINE.json configs[1]
    'mind_small_dev_n5_L3': (5, 2, 3, 65238, 73152, 37.5),       # the reference's argparse default (config.py:53)
    'wide_n8_L7': (8, 2, 7, 65238, 73152, 37.5),                 # BASELINE.json configs[3]
}


def algorithmic_bytes_per_pair(n_n, n_u, L, H=50, C=18):
    """B_alg of SURVEY.md section 8(d): layer-granular compulsory HBM traffic of the design."""
    S, T = n_n * n_n + n_u * n_u, n_n + n_u
    b_in = 4 * D * (n_n + H + 1) + S + n_n + (C + 1) + 8 * H + 16
    return L * (4 * D * (9 * T + n_n + H) + S) + b_in


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons while the timed region runs: through NVML in-process (pynvml, two cheap
    queries every 25 ms) when available, else with an `nvidia-smi -lms 100` child process (the recipe's clocks line)."""
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.proc = gpu_index, [], None
        self.mode = os.environ.get('DIGAT_BENCH_SAMPLER', 'nvml')
        self._stop_flag = False

    def _run_nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get('CUDA_VISIBLE_DEVICES')
        idx = int(vis.split(',')[self.gpu]) if vis and all(x.strip().isdigit() for x in vis.split(',')) else self.gpu
        h = nv.nvmlDeviceGetHandleByIndex(idx)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        bits = [(getattr(nv, 'nvmlClocksEventReasonHwSlowdown', 0x8), 3), (getattr(nv, 'nvmlClocksEventReasonHwThermalSlowdown', 0x40), 4),
                (getattr(nv, 'nvmlClocksEventReasonSwThermalSlowdown', 0x20), 5), (getattr(nv, 'nvmlClocksEventReasonSwPowerCap', 0x4), 6)]
        while not self._stop_flag:
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            row = [str(sm), str(mx), '', 'Not Active', 'Not Active', 'Not Active', 'Not Active']
            for bit, col in bits:
                if rs & bit:
                    row[col] = 'Active'
            self.rows.append(row)
            time.sleep(0.025)

    def run(self):
        if self.mode == 'none':
            return
        if self.mode == 'nvml':
            try:
                self._run_nvml()
                return
            except Exception:
                if self.rows:
                    return
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.FIELDS,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(',')])
        except Exception:
            pass

    def stop(self):
        self._stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i] == 'Active' for r in self.rows)]
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm)}


def build_workload(name, seed=0):
    from digat_b200 import synth
    N, hops, L, n_news, n_beh, cand = WORKLOADS[name]
    cfg = synth.make_config(SAG_neighbors=N, SAG_hops=hops, graph_depth=L)
    sd = synth.make_state_dict(cfg, D=D, seed=seed)
    corpus = synth.make_corpus(cfg, D=D, n_news=n_news, n_behaviors=n_beh, mean_candidates=cand, seed=seed)
    return cfg, sd, corpus


def oracle_pairs_per_second(cfg, sd, corpus, rows, budget_s, steps=None, warmup=2):
    """Times the CPU oracle (port of the reference algorithm, oracle/digat_oracle.py) on `rows` pairs per call."""
    from digat_b200 import synth
    from oracle import digat_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    P = O.cast_params(sd)
    order = ('news_graph_embeddings', 'news_graph', 'news_graph_mask', 'user_news_embedding', 'user_graph',
             'user_category_mask', 'user_category_indices')
    batch = synth.make_batch(corpus, np.arange(rows))
    args = [batch[k] for k in order]
    with torch.no_grad():
        c_n0 = O.news_graph_context(P, batch['news_graph_embeddings'], batch['news_graph_mask'])
        for _ in range(warmup):
            O.logits(*O.inference(P, *args, c_n0))
        n, t0 = 0, time.perf_counter()
        while True:
            O.logits(*O.inference(P, *args, c_n0))
            n += 1
            el = time.perf_counter() - t0
            if (steps is not None and n >= steps) or (steps is None and el >= budget_s):
                break
    return n * rows / el, el / n, cores, n


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cfg, sd, corpus = build_workload(args.workload)
    rows = 64
    pps, sec, cores, n = oracle_pairs_per_second(cfg, sd, corpus, rows, None, steps=args.steps, warmup=args.warmup)
    sample = '%d calls of DIGAT.inference on %d pairs each (oracle port of reference graphEncoders.py:189-198)' % (n, rows)
    print(json.dumps({
        'impl': 'reference', 'metric': 'impressions_scored_per_sec', 'value': pps, 'unit': 'pairs/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': sec * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': args.workload, 'rows_per_step': rows, 'device': 'host cpu'},
        'cpu_baseline': {'value': pps, 'unit': 'pairs/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': pps, 'unit': 'pairs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--batch', type=int, default=4096, help='pairs per step per GPU')
    ap.add_argument('--workload', default='mind_small_dev_n3_L3', choices=list(WORKLOADS))
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--cpu-budget', type=float, default=12.0, help='seconds of CPU-oracle timing for cpu_baseline')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--eager-train', action='store_true',
                    help='train mode: run the step eagerly instead of replaying it as one CUDA graph')
    ap.add_argument('--mode', default='score', choices=['score', 'train'],
                    help="'score' = the headline metric; 'train' = fwd+bwd(+DDP all-reduce)+Adam step, samples/s")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup

    if args.impl == 'reference':
        run_reference_arm(args)
        return
    if args.mode == 'train':
        run_train(args)
        return

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)

    from digat_b200 import _lib, scoring
    from digat_b200.graphEncoders import DIGAT
    _lib.require_device(local_rank)           # fails loudly without the sm_100a library / device

    cfg, sd, corpus = build_workload(args.workload)
    enc = DIGAT(cfg, D)
    enc.load_state_dict(sd)
    enc = enc.to(dev).eval()
    scorer = scoring.Scorer(enc, corpus, dev)
    scorer.cache_news_context()

    n_pairs = corpus.pair_behavior.shape[0]
    lo, hi = scoring.shard_range(n_pairs, rank, world)           # contiguous shard of the ordered pair list
    total_steps = args.warmup + args.steps
    assert (hi - lo) >= total_steps * args.batch, 'shard too small for steps*batch'
    pair_beh = torch.from_numpy(corpus.pair_behavior[lo:hi]).to(dev)
    pair_news = torch.from_numpy(corpus.pair_news[lo:hi]).to(dev)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    # ---------------------------------------------------------------- resident path (value)
    def step_resident(s):
        a = s * args.batch
        return scorer.score_resident(pair_beh[a:a + args.batch], pair_news[a:a + args.batch])

    def index_batches(s0, s1):
        return ((pair_beh[s * args.batch:(s + 1) * args.batch], pair_news[s * args.batch:(s + 1) * args.batch])
                for s in range(s0, s1))

    # The clock sampler (an nvidia-smi process) is started BEFORE the warm-up and must have delivered its first sample
    # before the timed region begins: its start-up (NVML initialisation) stalls kernel launches for tens of milliseconds,
    # which used to land inside the first timed steps every few runs (seen as 2-3x slower `value` at unchanged `e2e`).
    sampler = ClockSampler(local_rank)
    sampler.start()
    scoring.score_resident_batches(scorer, index_batches(0, args.warmup))
    t_wait = time.time()
    while not sampler.rows and sampler.mode != 'none' and time.time() - t_wait < 10.0:
        time.sleep(0.05)
    # A full (generation-2) garbage collection of a process that has imported torch walks ~10^6 objects: 50-150 ms during
    # which nothing is enqueued.  Every few runs one landed inside the timed steps (one step of 45-130 ms).  Collect now and
    # freeze the survivors so that collections inside the timed region only look at the objects created there.
    import gc
    gc.collect()
    gc.freeze()
    barrier()
    _lib.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    # the public pipelined driver: the flag kernels of batch k+1 are enqueued ahead of the encoder pass of batch k and
    # the host waits for their counts (the only synchronisation of a step) while that pass runs
    step_events = []
    scoring.score_resident_batches(scorer, index_batches(args.warmup, total_steps), events=step_events)
    e1.record()
    barrier()
    launches = _lib.launch_count()
    ms_resident = max_over_ranks(e0.elapsed_time(e1))
    marks = [e0] + step_events
    step_ms = sorted(marks[k].elapsed_time(marks[k + 1]) for k in range(len(marks) - 1))
    scorer.check_index_errors()

    # ---------------------------------------------------------------- end-to-end path (host buffers)
    host = [scoring.host_batch(corpus, np.arange(lo + s * args.batch, lo + (s + 1) * args.batch), pin=True)
            for s in range(total_steps)]
    h2d = int(sum(t.numel() * t.element_size() for t in host[0]))
    res = torch.empty((total_steps, args.batch), dtype=torch.float32).pin_memory()
    scoring.score_host_batches(scorer, host[:args.warmup], res[:args.warmup])
    barrier()
    e0.record()
    # the public pipelined call: every timed step's H2D copy and D2H read is issued inside the timed region
    # (batch k+1's copies overlap batch k's kernels on a side stream)
    scoring.score_host_batches(scorer, host[args.warmup:total_steps], res[args.warmup:total_steps])
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop()

    # ---------------------------------------------------------------- per-kernel timing -> roofline (untimed pass)
    prof = _lib.start_profile()
    for s in range(args.warmup, min(total_steps, args.warmup + 3)):
        step_resident(s)
    torch.cuda.synchronize()
    kernels = _lib.stop_profile()

    pairs_total = world * args.steps * args.batch
    value = pairs_total / (ms_resident * 1e-3)
    e2e = pairs_total / (ms_e2e * 1e-3)

    if rank == 0:
        peaks = {}
        pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
        if os.path.isfile(pk):
            peaks = json.load(open(pk))
        hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
        tensor_peak = float(peaks.get('bf16_tflops_sustained', peaks.get('bf16_tflops', 1590.0)))
        peak_src = 'measured (MEASURED_PEAKS.json)' if peaks else 'fallback (B200_PROFILING.md)'
        n_n, n_u, L = cfg.news_graph_size, cfg.max_history_num + cfg.category_num, cfg.graph_depth
        b_alg = algorithmic_bytes_per_pair(n_n, n_u, L)
        table = summarize_kernels(kernels, hbm_peak, tensor_peak)
        top = table[0] if table else None
        roofline = None
        if top:
            traffic, traffic_note = None, None
            tj = os.path.join(ROOT, 'profiles', 'r1_traffic.json')      # DRAM bytes per launch from the ncu --set full captures
            if os.path.isfile(tj):
                ent = json.load(open(tj)).get(top['kernel'])
                if ent:
                    traffic, traffic_note = ent['bytes'], ent['launch'] + ' / ' + ent['capture']
            roofline = {'kernel': top['kernel'], 'bound': top['bound'], 'achieved': top['achieved'], 'peak': top['peak'],
                        'unit': top['unit'], 'frac': top['achieved'] / top['peak'], 'traffic': traffic,
                        'traffic_note': traffic_note,
                        'share_of_step': top['share'], 'peak_source': peak_src, 'avg_launch_ms': top['avg_ms']}
        line = {
            'metric': 'impressions_scored_per_sec', 'value': value, 'unit': 'pairs/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_resident / args.steps,
            'step_ms': {'min': step_ms[0], 'median': step_ms[len(step_ms) // 2], 'max': step_ms[-1]},
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': args.workload, 'pairs_per_step_per_gpu': args.batch, 'SAG_neighbors': cfg.SAG_neighbors,
                       'SAG_hops': cfg.SAG_hops, 'news_graph_size': n_n, 'user_graph_size': n_u, 'graph_depth': L,
                       'D': D, 'n_news': int(corpus.news_embeddings.shape[0]), 'n_pairs': int(n_pairs),
                       'parallelism': 'pair-sharded x%d, no communication' % world,
                       'l2_policy': 'inputs larger than L2 (per-step intermediates ~%.1f GB)' %
                                    (args.batch * n_u * 3 * D * 4 / 1e9)},
            'e2e': {'value': e2e, 'unit': 'pairs/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': args.batch * 4,
                    'ms_per_step': ms_e2e / args.steps},
            'gpu_launches': launches,
            'clocks': clocks,
            'roofline': roofline,
            'roofline_path': {'bound': 'hbm', 'bytes_per_pair_alg': b_alg, 'achieved': value * b_alg / 1e9,
                              'peak': hbm_peak, 'unit': 'GB/s', 'frac': value * b_alg / 1e9 / hbm_peak,
                              'note': 'SURVEY 8(d) B_alg; the path is tensor/ALU-bound, see DESIGN.md'},
            'kernels': table[:8],
        }
        if not args.no_cpu_baseline and world == 1:
            pps, sec, cores, n = oracle_pairs_per_second(cfg, sd, corpus, 64, args.cpu_budget)
            line['cpu_baseline'] = {'value': pps, 'unit': 'pairs/s', 'cores': cores, 'kind': 'port',
                                    'sample': '%d calls x 64 pairs of the same workload (%.1f s of CPU work), '
                                              'oracle/digat_oracle.py inference, torch CPU fp32' % (n, n * sec)}
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


def run_train(args):
    """BASELINE.json configs[2]: DIGAT training fwd+bwd on synthetic MIND-shaped batches, per-GPU batch = 64 behaviours
    x (1 + 4 negatives) = 320 rows (reference config.py:31,34), dropout 0.2, DDP gradient all-reduce over NCCL when
    launched with N>1 ranks, clip-norm 1 + Adam step as reference trainer.py:98-105.  News-encoder excluded."""
    import torch.nn.functional as F
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        torch.distributed.init_process_group('nccl', device_id=dev)
    from digat_b200 import _lib, synth
    from digat_b200.graphEncoders import DIGAT
    from digat_b200.model import Model
    _lib.require_device(local_rank)
    N, hops, L = WORKLOADS[args.workload][:3]
    cfg = synth.make_config(SAG_neighbors=N, SAG_hops=hops, graph_depth=L, dropout_rate=0.2)
    sd = synth.make_state_dict(cfg, D=D, seed=0)
    corpus = synth.make_corpus(cfg, D=D, n_news=20000, n_behaviors=4096, mean_candidates=8.0, seed=rank)
    model = Model(cfg, D)
    model.graph_encoder.load_state_dict(sd)
    model = model.to(dev).train()
    net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank]) if world > 1 else model
    # whole-step CUDA graph (digat_b200/training.py) on one GPU; the DDP step runs eagerly (capturing DDP's reducer +
    # NCCL all-reduce failed in this PyTorch build: capture_end reported an invalidated capture from the backward)
    graphed = world == 1 and not args.eager_train
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, capturable=graphed)
    bs, news_num = 64, 5
    rng = np.random.Generator(np.random.PCG64(rank))
    emb = torch.from_numpy(corpus.news_embeddings).to(dev)
    node = torch.from_numpy(corpus.news_node_ID.astype(np.int64)).to(dev)
    ng, nm = torch.from_numpy(corpus.news_graph).to(dev), torch.from_numpy(corpus.news_graph_mask).to(dev)
    hist = torch.from_numpy(corpus.history.astype(np.int64)).to(dev)
    ug, cm, ci = (torch.from_numpy(x).to(dev) for x in (corpus.user_graph, corpus.user_category_mask, corpus.user_category_indices))

    def make_step_inputs():
        beh = torch.from_numpy(rng.integers(0, hist.shape[0], size=bs)).to(dev)
        cand = torch.from_numpy(rng.integers(1, emb.shape[0], size=(bs, news_num))).to(dev)
        return (emb[hist[beh]], ug[beh], cm[beh], ci[beh], emb[node[cand]], ng[cand], nm[cand])

    def step(inp):
        logits = net.forward_embeddings(*inp) if world == 1 else net.module.forward_embeddings(*inp)
        loss = (-F.log_softmax(logits, dim=1).select(1, 0)).mean()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        return loss

    if world > 1:
        # DDP hooks fire on the wrapped module's forward: route forward_embeddings through it
        class _Fwd(torch.nn.Module):
            def __init__(self, m):
                super().__init__()
                self.m = m

            def forward(self, *a):
                return self.m.forward_embeddings(*a)
        fwd = torch.nn.parallel.DistributedDataParallel(_Fwd(model), device_ids=[local_rank])

        def step(inp):  # noqa: F811
            logits = fwd(*inp)
            loss = (-F.log_softmax(logits, dim=1).select(1, 0)).mean()
            opt.zero_grad(set_to_none=True)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
            opt.step()
            return loss

    inputs = [make_step_inputs() for _ in range(args.warmup + args.steps)]
    if graphed:
        from digat_b200.training import GraphedTrainStep
        eager_step = step
        gstep = GraphedTrainStep(lambda *a: eager_step(a), inputs[0])
        step = lambda inp: gstep(*inp)                      # noqa: E731
    for s_ in range(args.warmup):
        step(inputs[s_])
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    _lib.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s_ in range(args.warmup, args.warmup + args.steps):
        loss = step(inputs[s_])
    e1.record()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        print(json.dumps({
            'metric': 'train_samples_per_sec', 'value': world * args.steps * bs / (ms * 1e-3), 'unit': 'samples/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': args.workload + ':train', 'behaviours_per_gpu': bs, 'candidates': news_num,
                       'rows_per_gpu': bs * news_num, 'dropout': 0.2, 'optimizer': 'Adam + clip_grad_norm 1',
                       'execution': 'whole step replayed as one CUDA graph' if graphed else 'eager',
                       'parallelism': 'DDP x%d, NCCL gradient all-reduce' % world},
            'gpu_launches': _lib.launch_count(), 'final_loss': float(loss)}))
    if world > 1:
        torch.distributed.destroy_process_group()


def summarize_kernels(records, hbm_peak, tensor_peak):
    """records: [(name, args, ms)] from _lib.stop_profile -> per-kernel-class share and roofline numbers."""
    agg = {}
    for name, a, ms in records:
        if name == 'digat_graph_layer_fwd':
            B, n, Dd = a[6], a[7], a[8]
            key = '%s[n=%d]' % (name, n)
            work, bound = B * (5 * n * Dd * 4 + n * n + Dd * 4), 'hbm'
        elif name == 'digat_linear_f32' or name == 'digat_linear_tf32x3':
            M, N, K = (a[7], a[8], a[9]) if name == 'digat_linear_f32' else (a[8], a[9], a[10])
            key = '%s[N=%d,K=%d,%s]' % (name, N, K, 'M>2048' if M > 2048 else 'M<=2048')
            work, bound = 2.0 * M * N * K, 'tensor'
        elif name == 'digat_attention_pool_fwd':
            B, m, Dd = a[12], a[13], a[14]
            key = '%s[m=%d]' % (name, m)
            work, bound = B * (m * Dd * 4 * (2 if a[3] else 1) + 2 * Dd * 4), 'hbm'
        elif name == 'digat_topic_segment_fwd':
            B, H, S, Dd = a[12], a[13], a[14], a[15]
            key, work, bound = name, B * (H * Dd * 4 + S * Dd * 4 + Dd * 4 + H * 8), 'hbm'
        elif name in ('digat_gather_sag_i32',):
            key, work, bound = name, a[6] * a[3] * a[7] * 4 * 2, 'hbm'
        elif name == 'digat_gather_rows_i32':
            key, work, bound = name, a[5] * a[6] * 4 * 2, 'hbm'
        elif name == 'digat_build_user_nodes':
            key, work, bound = name, a[6] * (a[7] + a[8]) * a[9] * 4 * 2, 'hbm'
        else:
            key, work, bound = name, 0.0, 'hbm'
        e = agg.setdefault(key, {'kernel': key, 'bound': bound, 'ms': 0.0, 'work': 0.0, 'launches': 0})
        e['ms'] += ms
        e['work'] += work
        e['launches'] += 1
    total = sum(e['ms'] for e in agg.values()) or 1.0
    out = []
    for e in agg.values():
        if e['bound'] == 'hbm':
            ach, peak, unit = e['work'] / (e['ms'] * 1e-3) / 1e9, hbm_peak, 'GB/s'
        else:
            ach, peak, unit = e['work'] / (e['ms'] * 1e-3) / 1e12, tensor_peak, 'TFLOP/s'
        out.append({'kernel': e['kernel'], 'bound': e['bound'], 'share': e['ms'] / total, 'avg_ms': e['ms'] / e['launches'],
                    'launches': e['launches'], 'achieved': ach, 'peak': peak, 'unit': unit, 'frac': ach / peak})
    out.sort(key=lambda r: -r['share'])
    return out


if __name__ == '__main__':
    main()
