#!/usr/bin/env python
"""VB-EM frames/s on the HMM-GMM hot path (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg3|cfg2|cfg2ali] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Default workload = BASELINE configs[2], the north-star target: HMM-GMM 1000 states x 8 diag-Gauss, 1250 utterances x
1000 frames per GPU (at N = 8 that IS the 10 000-utterance configuration).  BASELINE configs[1] (100 states x 1
Gaussian, 4096 utterances per GPU) is measured in the same run and reported under `secondary`.

A "step" is one full VB-EM iteration over the rank's resident utterances: emission weights + KL, per-frame llh (KA),
forward-backward (KB), statistics (KC), one all-reduce of the flat statistics buffer, M-step.  `value` = frames of all
ranks / (max over ranks of the CUDA-event time per step), inputs resident in HBM.  `e2e` = the same through the public
engine API with the features in pinned HOST memory (H2D copy of every frame and D2H read of the ELBO inside the timed
region).  `elbo_check` = ELBO of the engine against the reference itself run in float64 on the CPU (baseline/_ref;
the numpy fp64 oracle when the reference is not installed) on a small subset, same initial model, in this run.
`--impl reference` times the unmodified reference (beer.evidence_lower_bound + backward + optimizer step) on all host
cores instead.
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

METRIC = 'VB-EM frames/sec on HMM-GMM (40-d fbank)'
UNIT = 'frames/s'


def workload_name(cfg, c):
    if c.get('gmm'):
        return (f"{cfg}: {c['n_comp']}-component diagonal GMM (the E-step + conjugate M-step inside GSM-GMM), {c['dim']}-d "
                f"synthetic fbank, {c['n_utts']} utterances x {c['n_frames']} frames per GPU, no HMM")
    K = c['n_units'] * c['n_states']
    return (f"{cfg}: HMM-GMM {K} states x {c['n_comp']} diag-Gauss, {c['dim']}-d synthetic fbank, "
            f"{c['n_utts']} utterances x {c['n_frames']} frames per GPU, phone-loop graph "
            f"({c['n_units']} units x {c['n_states']} states)"
            + (', every utterance aligned to its own left-to-right chain' if c.get('aligned') else '')
            + (', Viterbi training' if c.get('viterbi') else ''))


def measured_peaks():
    """(HBM GB/s, source): the driver-written MEASURED_PEAKS.json when present, else the fallback of
    B200_PROFILING.md."""
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as f:
            p = json.load(f)
        if isinstance(p.get('hbm_gbs'), (int, float)):
            return float(p['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except (OSError, ValueError):
        pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '50', '-i', str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


# ---------------------------------------------------------------------------------------
# CPU side: the reference itself (baseline/_ref, else /root/reference), else the numpy port
# ---------------------------------------------------------------------------------------

def host_utterances(c, n_utts, seed):
    """`n_utts` synthetic utterances of configuration `c` sampled on the host (numpy, fp32)."""
    from oracle import beer_oracle as O
    if c.get('gmm'):
        from beer_b200.synthetic import sample_gmm_frames
        X = sample_gmm_frames(n_utts * c['n_frames'], c['dim'], seed).numpy()
        return [X[i * c['n_frames']:(i + 1) * c['n_frames']] for i in range(n_utts)], None
    rng = np.random.default_rng(seed)
    graph, _, _ = O.phone_loop_graph(c['n_units'], c['n_states'])
    K = c['n_units'] * c['n_states']
    means = 2.0 * rng.standard_normal((K, c['dim']))
    return O.sample_utterances(rng, graph, means, n_utts, c['n_frames']), graph


def _port_worker(args):
    """numpy port of the reference E-step + accumulate over a shard (fallback when the reference is absent)."""
    c, utts, dtype = args
    from oracle import beer_oracle as O
    if c.get('gmm'):
        M, D = c['n_comp'], c['dim']
        rng = np.random.default_rng(2)
        post = tuple(a.astype(dtype) for a in (rng.standard_normal((M, D)), np.ones((M, 1)), np.ones((M, 1)), np.ones((M, D))))
        dpost = (np.ones(M) / M).astype(dtype)
        t0 = time.perf_counter()
        frames = 0
        with np.errstate(all='ignore'):
            for X in utts:
                O.gmm_estep(X.astype(dtype), post, dpost)
                frames += len(X)
        return dict(frames=frames, seconds=time.perf_counter() - t0)
    graph, _, _ = O.phone_loop_graph(c['n_units'], c['n_states'])
    K = c['n_units'] * c['n_states']
    M, D = K * c['n_comp'], c['dim']
    rng = np.random.default_rng(2)
    post = (rng.standard_normal((M, D)), np.ones((M, 1)), np.ones((M, 1)), np.ones((M, D)))
    dpost = np.ones((K, c['n_comp'])) / c['n_comp'] if c['n_comp'] > 1 else None
    graph = tuple(np.asarray(a, dtype=dtype) for a in graph[:3]) + (graph[3],)
    post = tuple(a.astype(dtype) for a in post)
    t0 = time.perf_counter()
    frames = 0
    with np.errstate(all='ignore'):
        for X in utts:
            O.hmm_estep(X.astype(dtype), post, None if dpost is None else dpost.astype(dtype), graph)
            frames += len(X)
    return dict(frames=frames, seconds=time.perf_counter() - t0)


class CpuArm:
    """A pool of single-threaded worker processes, one per host core: the reference's own parallel style
    (recipes/zrc2019/steps/aud_gnu_parallel.sh:73-86, one `beer hmm accumulate` job per shard, then `update`)."""

    def __init__(self, c, n_workers=None):
        import multiprocessing as mp
        from baseline import reference_arm as R
        self.c, self.R = c, R
        self.kind = 'reference' if R.find_reference() is not None else 'port'
        self.n_workers = n_workers or os.cpu_count() or 1
        os.environ.setdefault('OMP_NUM_THREADS', '1')
        os.environ.setdefault('MKL_NUM_THREADS', '1')
        self.pool = mp.get_context('spawn').Pool(self.n_workers)

    def close(self):
        self.pool.close()
        self.pool.join()

    def throughput(self, shards):
        """One accumulate + update pass over `shards` (one per worker, all at once): frames / slowest worker."""
        if self.kind == 'reference':
            res = self.pool.map(self.R._worker, [(self.c, s, 1, False, False, None) for s in shards])
        else:
            res = self.pool.map(_port_worker, [(self.c, s, np.float32) for s in shards])
        wall = max(r['seconds'] for r in res)
        frames = sum(r['frames'] for r in res)
        return frames / wall, frames, wall

    def describe(self, upw, T):
        what = ('unmodified reference (beer.evidence_lower_bound per utterance + backward + VBConjugateOptimizer.step, '
                'torch CPU fp32)' if self.kind == 'reference' else
                'numpy fp32 port of the reference E-step + accumulate (reference not installed)')
        return f'{self.n_workers} single-threaded worker processes x {upw} utterance(s) x {T} frames: {what}'

    def elbo_fp64(self, utts, datasize):
        """(sum of the per-utterance ELBOs in float64, initial model arrays)."""
        if self.kind == 'reference':
            res = self.pool.map(self.R._worker, [(self.c, [u], 1, True, i == 0, datasize, False) for i, u in enumerate(utts)])
            # every worker's `elbo` = empty accumulator + its utterance: plain sum (objectives.py:78-90)
            return sum(r['per_utt'][0] for r in res), res[0]['model'], 'reference (float64)'
        from oracle import beer_oracle as O
        c = self.c
        if c.get('gmm'):
            M, D = c['n_comp'], c['dim']
            rng = np.random.default_rng(2)
            prior = (np.zeros((M, D)), np.ones((M, 1)), np.ones((M, 1)), np.ones((M, D)))
            post = (rng.standard_normal((M, D)).astype(np.float32).astype(np.float64),) + prior[1:]
            dp = np.ones(M) / M
            kl = O.normalgamma_kl(post, prior).sum() + O.dirichlet_kl(dp, dp).sum()
            elbo = sum(O.elbo_value(O.gmm_estep(u.astype(np.float64), post, dp)['exp_llh'], kl, datasize) for u in utts)
            return elbo, dict(ng_prior=prior, ng_post=post, dir_prior=dp, dir_post=dp), 'numpy oracle (float64)'
        K = c['n_units'] * c['n_states']
        M, D, C = K * c['n_comp'], c['dim'], c['n_comp']
        rng = np.random.default_rng(2)
        prior = (np.zeros((M, D)), np.ones((M, 1)), np.ones((M, 1)), np.ones((M, D)))
        post = (rng.standard_normal((M, D)).astype(np.float32).astype(np.float64),) + prior[1:]
        model = dict(ng_prior=prior, ng_post=post)
        if C > 1:
            model['dir_prior'] = model['dir_post'] = np.ones((K, C)) / C
        graph, _, _ = O.phone_loop_graph(c['n_units'], c['n_states'])
        elbo, _, _, _ = O.vb_iteration_hmm([u.astype(np.float64) for u in utts], prior, post, model.get('dir_prior'),
                                           model.get('dir_post'), graph, datasize=datasize)
        return elbo, model, 'numpy oracle (float64)'


def cpu_sizes(c):
    """(utterances per worker, frames kept of each utterance) for ~5-10 s of CPU work per pass: the cost of the
    reference is linear in the frames, so the large configuration is sampled with the first 250 frames of one
    utterance per worker."""
    if c.get('gmm'):
        return 64, c['n_frames']
    big = c['n_comp'] * c['n_units'] * c['n_states'] > 2000
    return (1, min(250, c['n_frames'])) if big else (16, c['n_frames'])


def cpu_shards(c, n_workers, seed=1000):
    upw, keep = cpu_sizes(c)
    utts, _ = host_utterances(c, n_workers * upw, seed=seed)
    utts = [u[:keep] for u in utts]
    return [utts[i * upw:(i + 1) * upw] for i in range(n_workers)], upw, keep


def run_reference(args, name, c):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    arm = CpuArm(c)
    shards, upw, keep = cpu_shards(c, arm.n_workers)
    for _ in range(args.warmup):
        arm.throughput([[s[0][:50]] for s in shards])
    tot_frames, tot_time = 0, 0.0
    for _ in range(args.steps):
        _, frames, wall = arm.throughput(shards)
        tot_frames += frames
        tot_time += wall
    arm.close()
    value = tot_frames / tot_time
    sample = arm.describe(upw, keep) + ' per step'
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * tot_time / max(args.steps, 1),
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': workload_name(name, c), 'sample': sample},
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': arm.n_workers, 'kind': arm.kind, 'sample': sample},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------

class Ctx:
    pass


def make_engine_from_arrays(ctx, c, model, utts, plan, datasize):
    import torch
    from beer_b200.engine import EmissionParams, VBEngine, WeightGroup
    dev = ctx.dev
    K, C = (1, c['n_comp']) if c.get('gmm') else (c['n_units'] * c['n_states'], c['n_comp'])

    def ng(t):
        m, k, a, b = t
        f = lambda x: torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float32, device=dev).contiguous()
        return f(m), f(np.reshape(k, -1)), f(np.reshape(a, -1)), f(b)

    prior, post = ng(model['ng_prior']), ng(model['ng_post'])
    groups, comp_off = (), None
    if C > 1:
        dp = torch.as_tensor(model['dir_prior'], dtype=torch.float32, device=dev).reshape(K, C).contiguous()
        dq = torch.as_tensor(model['dir_post'], dtype=torch.float32, device=dev).reshape(K, C).contiguous()
        groups = (WeightGroup(0, K, C, dp, dq),)
        comp_off = np.arange(K + 1) * C
    em = EmissionParams(prior, post, comp_off=comp_off, weight_groups=groups)
    return VBEngine(em, plan, utts, datasize=datasize, distributed=False)


def elbo_check(ctx, name, c, arm, n_check):
    """Engine (fp32, GPU) against the reference in float64 (CPU) on `n_check` utterances, same initial model:
    the summed ELBO of accumulate + update (objectives.py:180-184 added up as in accumulate.py:39-59)."""
    import torch
    from beer_b200 import ops, synthetic
    from beer_b200.engine import Utterances
    utts, _ = host_utterances(c, n_check, seed=4242)
    utts = [u[:cpu_sizes(c)[1]] for u in utts]
    N = float(sum(len(u) for u in utts))
    t0 = time.perf_counter()
    want, model, kind = arm.elbo_fp64(utts, N)
    plan = None
    if not c.get('gmm'):
        graph, _, _ = synthetic.phone_loop_graph(c['n_units'], c['n_states'])
        K = c['n_units'] * c['n_states']
        plan = ops.GraphPlan(graph.init_log_probs.numpy(), graph.final_log_probs.numpy(),
                             graph.trans_log_probs.numpy(), graph.pdf_id_mapping, n_pdfs=K)
    X = torch.as_tensor(np.concatenate(utts), dtype=torch.float32, device=ctx.dev)
    eng = make_engine_from_arrays(ctx, c, model, Utterances(X, [len(u) for u in utts]), plan, N)
    got = float(eng.step().item())
    return {'rel_err': abs(got - want) / abs(want), 'n_utts': n_check, 'frames': int(N), 'engine_elbo': got,
            'cpu_elbo': want, 'against': kind, 'seconds': time.perf_counter() - t0}


def run_config(ctx, args, name, c, steps, warmup, primary):
    """Time one configuration; returns the dict of its measurements (rank 0: everything, others: None)."""
    import torch
    import torch.distributed as dist
    from beer_b200 import ops, synthetic
    from beer_b200.engine import EmissionParams, Utterances, VBEngine, WeightGroup
    dev, world, rank = ctx.dev, ctx.world, ctx.rank

    gmm = bool(c.get('gmm'))
    K = 1 if gmm else c['n_units'] * c['n_states']
    C = c['n_comp']
    M, D, T, U = K * C, c['dim'], c['n_frames'], c['n_utts']
    plan = None
    if gmm:
        X = synthetic.sample_gmm_frames(U * T, D, seed=100 + rank, device=dev)
    else:
        graph, _, _ = synthetic.phone_loop_graph(c['n_units'], c['n_states'])
        plan = ops.GraphPlan(graph.init_log_probs.numpy(), graph.final_log_probs.numpy(),
                             graph.trans_log_probs.numpy(), graph.pdf_id_mapping, n_pdfs=K)
        gen = torch.Generator().manual_seed(7)
        means = 2.0 * torch.randn(K, D, generator=gen)
    if gmm:
        pass
    elif c.get('aligned'):
        X, paths = synthetic.sample_utterances(graph, means, U, T, seed=100 + rank, device=dev, return_paths=True)
        plan = ops.ChainBatch.from_arrays(*synthetic.alignment_chains(paths, c['n_states']), device=dev)
        del paths
    else:
        X = synthetic.sample_utterances(graph, means, U, T, seed=100 + rank, device=dev)
    utts = Utterances(X, [T] * U)

    def make_engine(utts=utts, chunk_frames=args.chunk_frames, use_graph=False, sparse_stats=None):
        prior, post = synthetic.initial_normal_gamma(M, D, seed=2, device=dev)
        groups, comp_off = (), None
        if C > 1:
            conc = torch.full((K, C), 1.0 / C, device=dev)
            groups = (WeightGroup(0, K, C, conc.clone(), conc.clone()),)
            comp_off = np.arange(K + 1) * C
        em = EmissionParams(prior, post, comp_off=comp_off, weight_groups=groups)
        return VBEngine(em, plan, utts, datasize=float(world * U * T), chunk_frames=chunk_frames,
                        distributed=world > 1, use_graph=use_graph, viterbi=bool(c.get('viterbi')),
                        sparse_stats=sparse_stats)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    eng = make_engine(use_graph=not args.no_graph)
    elbos = []
    for _ in range(warmup):
        elbos.append(eng.step().clone())
    barrier()
    eng.gpu_launches = 0
    sampler = ClockSampler(ctx.local_rank)
    if rank == 0 and primary:
        sampler.start()
    t_wall = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.nvtx.range_push('bench_timed')     # ncu --nvtx --nvtx-include "bench_timed/" selects these launches
    ev0.record()
    for _ in range(steps):
        elbos.append(eng.step().clone())     # the graph's ELBO buffer is overwritten by the next step
    ev1.record()
    torch.cuda.nvtx.range_pop()
    barrier()
    wall = time.perf_counter() - t_wall
    clocks = sampler.stop() if (rank == 0 and primary) else None
    ms = ev0.elapsed_time(ev1)
    launches = eng.gpu_launches
    # per-stage kernel durations: the same iteration launched eagerly with CUDA events around the stages
    eng.profile = {}
    for _ in range(3):
        eng.step()
    torch.cuda.synchronize()
    stage_ms = {k: float(np.mean([a.elapsed_time(b) for a, b in v])) for k, v in eng.profile.items()}
    eng.profile = None
    active = getattr(eng, 'active_fraction', None)
    active = float(active.item()) if active is not None else None
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if primary and ms < 400.0:
        # the timed region is shorter than a few nvidia-smi sampling periods: sample the same step loop,
        # untimed, for ~0.6 s right behind it (same number of extra steps on every rank)
        n_probe = int(600.0 / max(ms / steps, 1e-3)) + 1
        probe = ClockSampler(ctx.local_rank)
        if rank == 0:
            probe.start()
        for _ in range(n_probe):
            eng.step()
        barrier()
        if rank == 0:
            clocks = probe.stop()
            clocks['window'] = (f'{n_probe} more steps of the same loop right after the timed region '
                                '(timed region too short to sample)')
    frames_per_step = world * U * T
    value = frames_per_step * steps / (ms * 1e-3)
    elbo_pf = [float(eng.elbo_per_frame(e).item()) for e in elbos]
    tensor_kind = getattr(eng, 'tensor_kind', 'tf32')

    # ---- the same steps with the statistics kernel working on EVERY (frame tile, Gaussian tile) pair ----
    # (mixtures skip the pairs whose weights are exactly zero in the kernel's fp16 operands -- the same non-zero products, summed in another grouping,
    # beer_mix16_accumulate_blocks --, which makes the step time depend on how peaked the posteriors are: the dense
    # figure is the data-independent one)
    dense_ms = None
    if active is not None:
        eng_d = make_engine(use_graph=not args.no_graph, sparse_stats=False)
        for _ in range(warmup):
            eng_d.step()
        barrier()
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record()
        for _ in range(steps):
            eng_d.step()
        d1.record()
        barrier()
        td = torch.tensor([d0.elapsed_time(d1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(td, op=dist.ReduceOp.MAX)
        dense_ms = float(td.item()) / steps
        del eng_d

    # ---- end to end: features in pinned host memory, H2D every step, ELBO read back --------
    host_X = torch.empty(X.shape, dtype=torch.float32, pin_memory=True)
    host_X.copy_(X)
    del eng
    # features stay in pinned host memory: the engine streams them in chunks of whole utterances (the H2D
    # copy of chunk i+1 under the kernels of chunk i) and the ELBO is read back to the host every step
    eng2 = make_engine(Utterances(host_X, [T] * U), chunk_frames=args.e2e_chunk_frames or max(T, min(U * T, 4_200_000)))
    n_e2e = max(1, min(steps, 5))

    def e2e_step():
        return float(eng2.step().item())

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = frames_per_step * n_e2e / float(te.item())
    h2d_rate = X.numel() * 4 * n_e2e / e2e_s / 1e9       # this rank's host -> device rate while streaming
    rates = torch.tensor([h2d_rate], device=dev, dtype=torch.float64)
    all_rates = [rates.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(all_rates, rates)
    del eng2, host_X
    if rank != 0:
        return None

    peak, peak_src = measured_peaks()
    Q = 2 * D + 2
    nf = U * T
    # algorithmic work per frame of every stage: bytes B_alg = 8D + 16K split over KA (read X, write
    # llh), KB (read llh once more, write + read alpha) and KC (read X); flops F_alg = 4 Q M split over KA and KC
    alg_bytes = {'KA_emission_llh': 4 * D + 4 * K, 'KB_forward_backward': 12 * K, 'KC_accumulate': 4 * D}
    if gmm:      # B_alg = 8 D for the GMM-only configurations (the per-frame llh stays on chip in the count)
        alg_bytes = {'KA_emission_llh': 4 * D, 'KB_forward_backward': 0, 'KC_accumulate': 4 * D}
    alg_flops = {'KA_emission_llh': 2 * Q * M, 'KC_accumulate': 2 * Q * M}
    traffic_pf = {}
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath):        # dram bytes per frame per launch from the committed ncu capture
        with open(tpath) as f:
            traffic_pf = json.load(f).get(name, {})
    tc_peak = ctx.tc_peak[tensor_kind] / 3.0      # 3-pass split (hi.hi + lo.hi + hi.lo) of the measured kind peak
    fps = nf / (ms / steps * 1e-3)                # per GPU
    Pn = c['n_units']
    nnz = 0 if gmm else 2 * K - Pn + Pn * Pn
    mhz = (clocks or {}).get('sm_mhz') or 1965.0
    fractions = {
        'hbm': fps * (8 * D + (0 if gmm else 16 * K)) / 1e9 / peak,
        'tensor_3pass': fps * 4 * Q * M / 1e12 / tc_peak,
        'scan_mufu': fps * 2 * nnz / (148 * 16 * mhz * 1e6),
        'tensor_peak_tflops': tc_peak,
        'tensor_peak_source': f'beer_probe_mma kind::{tensor_kind} measured in this run '
                              f'({ctx.tc_peak[tensor_kind]:.0f} TFLOP/s dense) / 3 passes',
        'alg_bytes_per_frame': 8 * D + (0 if gmm else 16 * K), 'alg_flops_per_frame': 4 * Q * M}
    # the binding roofline of the configuration: tensor pipe when the 3-pass contraction needs more time
    # than the algorithmic bytes at HBM speed
    tensor_bound = (4 * Q * M / 1e12 / tc_peak) > ((8 * D + (0 if gmm else 16 * K)) / 1e9 / peak)
    dom = max((k for k in stage_ms if k in alg_bytes), key=lambda k: stage_ms[k], default=None)
    roofline = None
    if dom is not None:
        sec = stage_ms[dom] * 1e-3
        if tensor_bound and dom in alg_flops:
            achieved = alg_flops[dom] * nf / sec / 1e12
            roofline = {'bound': 'tensor', 'kernel': dom, 'achieved': achieved, 'peak': tc_peak, 'unit': 'TFLOP/s',
                        'frac': achieved / tc_peak, 'peak_source': fractions['tensor_peak_source'],
                        'alg_flops_per_frame': alg_flops[dom]}
        else:
            achieved = alg_bytes[dom] * nf / sec / 1e9
            roofline = {'bound': 'hbm', 'kernel': dom, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                        'frac': achieved / peak, 'peak_source': peak_src, 'alg_bytes_per_frame': alg_bytes[dom]}
        roofline.update({
            'traffic': (traffic_pf[dom] * nf if dom in traffic_pf else None),
            'traffic_unit': 'bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)',
            'launch_ms': stage_ms[dom], 'stage_ms': stage_ms,
            'step_frac': fractions['tensor_3pass'] if tensor_bound else fractions['hbm'],
            'step_bound': 'tensor' if tensor_bound else 'hbm', 'step_fractions': fractions})
    return {'workload': workload_name(name, c), 'value': value, 'ms_per_step': ms / steps, 'steps': steps,
            'warmup': warmup, 'clocks': clocks, 'wall_s_timed_region': wall,
            'l2': f'inputs larger than L2 ({X.numel() * 4 / 2**20:.0f} MiB of features per GPU, no flush needed)',
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(world * X.numel() * 4),
                    'd2h_bytes_per_step': 8 * world, 'steps': n_e2e,
                    'h2d_gbs_per_rank': [round(float(r.item()), 2) for r in all_rates]},
            'gpu_launches': launches, 'roofline': roofline, 'tensor_kind': tensor_kind,
            'sparse_statistics': (None if active is None else {
                'active_pair_fraction': active,
                'what': 'fraction of (64-frame tile, 128-Gaussian tile) pairs the statistics kernel works on in the last '
                        'timed iterations; the others carry weights that are exactly zero in its fp16 operands '
                        '(the same non-zero products as the dense kernel, grouped differently into its fp32 partial sums)',
                'dense_ms_per_step': dense_ms,
                'dense_value': (frames_per_step / (dense_ms * 1e-3)) if dense_ms else None}),
            'elbo_per_frame': {'first': elbo_pf[0], 'last': elbo_pf[-1]}}


def run_gpu(args, configs):
    import torch
    import torch.distributed as dist
    from beer_b200 import ops
    from beer_b200.synthetic import CONFIGS

    ctx = Ctx()
    ctx.world = int(os.environ.get('WORLD_SIZE', '1'))
    ctx.rank = int(os.environ.get('RANK', '0'))
    ctx.local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if ctx.world != args.gpus:
        raise SystemExit(f'--gpus {args.gpus} needs {args.gpus} ranks (torchrun), got WORLD_SIZE={ctx.world}')
    torch.cuda.set_device(ctx.local_rank)
    ctx.dev = torch.device('cuda', ctx.local_rank)
    if ctx.world > 1:
        from beer_b200.engine import bind_to_gpu_cpus
        ctx.cpus = bind_to_gpu_cpus(ctx.local_rank)      # pinned feature buffers next to this rank's GPU (e2e leg)
        # NCCL announces its version on stdout at the first collective: keep stdout to the one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=ctx.dev)
            dist.all_reduce(torch.zeros(1, device=ctx.dev))
            torch.cuda.synchronize()
        finally:
            os.dup2(saved, 1)
            os.close(saved)
    ops.require_cuda()
    # tensor-pipe peaks of the MMA kinds the kernels use, measured on this GPU before the timed work
    ctx.tc_peak = {k: ops.probe_mma_tflops(k) for k in ('tf32', 'f16')}

    results = []
    for i, name in enumerate(configs):
        c = dict(CONFIGS[name])
        if args.n_utts:
            c['n_utts'] = args.n_utts
        if args.viterbi:
            c['viterbi'] = True
        steps, warmup = (args.steps, args.warmup) if i == 0 else (max(args.steps, 10), max(args.warmup, 3))
        results.append((name, c, run_config(ctx, args, name, c, steps, warmup, primary=(i == 0))))
        torch.cuda.empty_cache()
    if ctx.world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    if ctx.rank == 0:
        # CPU side, after the last collective: ELBO check of every configuration against the reference in float64
        # and (one GPU only) the reference timed on the host cores
        for name, c, r in results:
            arm = None
            if not args.no_elbo_check or (ctx.world == 1 and not args.no_cpu_baseline):
                arm = CpuArm(c)
            if not args.no_elbo_check:
                n_check = 2 if c['n_comp'] * max(c['n_units'] * c['n_states'], 1) > 2000 and not c.get('gmm') else 8
                r['elbo_check'] = elbo_check(ctx, name, c, arm, min(n_check, arm.n_workers))
            r['cpu_baseline'] = None
            if ctx.world == 1 and not args.no_cpu_baseline:
                shards, upw, keep = cpu_shards(c, arm.n_workers)
                arm.throughput([[s[0][:50]] for s in shards])  # warm the workers (imports, first-call set-up)
                v, frames, took = arm.throughput(shards)
                r['cpu_baseline'] = {'value': v, 'unit': UNIT, 'cores': arm.n_workers, 'kind': arm.kind,
                                     'sample': arm.describe(upw, keep) + f' ({took:.1f} s)'}
            if arm is not None:
                arm.close()
        name, c, r = results[0]
        line = {'metric': METRIC, 'value': r['value'], 'unit': UNIT, 'n_gpus': ctx.world, 'steps': r['steps'],
                'warmup': r['warmup'], 'ms_per_step': r['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                'config': {'workload': r['workload'], 'l2': r['l2'], 'chunk_frames': args.chunk_frames,
                           'tensor_kind': r['tensor_kind'],
                           'statistics': ('the mixture statistics kernel works only on the (frame tile, Gaussian tile) pairs '
                                          'whose weights are not exactly zero in its fp16 operands; `sparse_statistics` has '
                                          'the marked fraction and the time of the same steps with every pair worked on'
                                          if r.get('sparse_statistics') else 'dense'),
                           'parallelism': f'dp{ctx.world} (utterances sharded, one all-reduce of the statistics per step)'},
                'clocks': r['clocks'], 'wall_s_timed_region': r['wall_s_timed_region'], 'e2e': r['e2e'],
                'gpu_launches': r['gpu_launches'], 'roofline': r['roofline'], 'cpu_baseline': r.get('cpu_baseline'),
                'elbo_check': r.get('elbo_check'), 'elbo_per_frame': r['elbo_per_frame'],
                'sparse_statistics': r.get('sparse_statistics'),
                'tensor_peaks_tflops': {k: round(v, 1) for k, v in ctx.tc_peak.items()}}
        if len(results) > 1:
            line['secondary'] = {n: {k: v for k, v in rr.items() if k not in ('clocks', 'wall_s_timed_region')}
                                 for n, _, rr in results[1:]}
        print(json.dumps(line), flush=True)
    if ctx.world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--config', default=None, help='cfg3 (default, + cfg2 as `secondary`), cfg2, cfg2ali, cfg5 (GMM only)')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--chunk-frames', type=int, default=None)
    ap.add_argument('--e2e-chunk-frames', type=int, default=None)
    ap.add_argument('--n-utts', type=int, default=None, help='override utterances per GPU (debug)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-elbo-check', action='store_true')
    ap.add_argument('--no-secondary', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='launch every kernel eagerly (no CUDA graph)')
    ap.add_argument('--viterbi', action='store_true', help='Viterbi training (one-hot posteriors of the best path) '
                    'instead of forward-backward; a secondary workload, not the BASELINE metric')
    args = ap.parse_args()
    from beer_b200.synthetic import CONFIGS
    if args.config is None:
        configs = ['cfg3'] + ([] if args.no_secondary else ['cfg2'])
    else:
        configs = [args.config]
    if args.impl == 'reference':
        c = dict(CONFIGS[configs[0]])
        if args.n_utts:
            c['n_utts'] = args.n_utts
        run_reference(args, configs[0], c)
    else:
        run_gpu(args, configs)


if __name__ == '__main__':
    main()
