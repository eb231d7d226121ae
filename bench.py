#!/usr/bin/env python
"""VB-EM frames/s on the HMM-GMM hot path (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg2|cfg3] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full VB-EM iteration over the rank's resident utterances: emission weights +
KL, per-frame llh (KA), forward-backward (KB), statistics (KC), one all-reduce of the flat
statistics buffer, M-step.  `value` = frames of all ranks / (max over ranks of the CUDA-event
time per step), inputs resident in HBM.  `e2e` = the same through the public API with the
features in pinned HOST memory (H2D copy of every frame and D2H read of the ELBO inside the
timed region).  `--impl reference` times the CPU restatement of the reference algorithm
(oracle/beer_oracle.py, numpy) on all host cores instead.
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
    K = c['n_units'] * c['n_states']
    return (f"{cfg}: HMM-GMM {K} states x {c['n_comp']} diag-Gauss, {c['dim']}-d synthetic fbank, "
            f"{c['n_utts']} utterances x {c['n_frames']} frames per GPU, phone-loop graph "
            f"({c['n_units']} units x {c['n_states']} states)"
            + (', every utterance aligned to its own left-to-right chain' if c.get('aligned') else '')
            + (', Viterbi training' if c.get('viterbi') else ''))


def measured_peaks():
    """HBM peak in GB/s: the driver-written MEASURED_PEAKS.json when present (key `hbm_gbs`; any numeric key
    naming hbm is accepted), else the fallback of B200_PROFILING.md."""
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        try:
            with open(path) as f:
                p = json.load(f)

            def find(d):
                if isinstance(d, dict):
                    if isinstance(d.get('hbm_gbs'), (int, float)):
                        return float(d['hbm_gbs'])
                    for k, v in d.items():
                        if 'hbm' in str(k).lower() and isinstance(v, (int, float)):
                            return float(v) * (1000.0 if float(v) < 50 else 1.0)     # TB/s -> GB/s
                    for v in d.values():
                        r = find(v)
                        if r:
                            return r
                return None
            v = find(p)
            if v:
                return v, 'measured'
        except (OSError, ValueError):
            pass
    return 6650.0, 'fallback'


def measured_bf16_tflops():
    """Dense bf16 TFLOP/s (sustained figure: the contraction runs inside a long step) from MEASURED_PEAKS.json, else
    the 1590 fallback of B200_PROFILING.md."""
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as f:
            p = json.load(f)
        for k in ('bf16_tflops_sustained', 'bf16_tflops'):
            if isinstance(p.get(k), (int, float)):
                return float(p[k]), 'measured'
    except (OSError, ValueError):
        pass
    return 1590.0, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
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
# CPU baseline: the oracle port of the reference algorithm, one worker per host core
# ---------------------------------------------------------------------------------------

def _cpu_worker(args):
    cfg, seed, n_utts = args
    from oracle import beer_oracle as O
    c = cfg
    rng = np.random.default_rng(seed)
    graph, _, _ = O.phone_loop_graph(c['n_units'], c['n_states'])
    K = c['n_units'] * c['n_states']
    M = K * c['n_comp']
    D = c['dim']
    means = 2.0 * rng.standard_normal((K, D))
    utts = O.sample_utterances(rng, graph, means, n_utts, c['n_frames'])
    prior = (np.zeros((M, D)), np.ones((M, 1)), np.ones((M, 1)), np.ones((M, D)))
    post = (rng.standard_normal((M, D)), np.ones((M, 1)), np.ones((M, 1)), np.ones((M, D)))
    dprior = dpost = None
    if c['n_comp'] > 1:
        dprior = np.ones((K, c['n_comp'])) / c['n_comp']
        dpost = dprior.copy()
    utts = [u.astype(np.float32) for u in utts]       # the reference default dtype is fp32
    graph32 = tuple(np.asarray(a, dtype=np.float32) for a in graph[:3]) + (graph[3],)
    post32 = tuple(a.astype(np.float32) for a in post)
    t0 = time.perf_counter()
    frames = 0
    with np.errstate(all='ignore'):
        for X in utts:
            O.hmm_estep(X, post32, None if dpost is None else dpost.astype(np.float32), graph32)
            frames += len(X)
    return frames, time.perf_counter() - t0


def cpu_reference_throughput(c, utts_per_worker, n_workers=None):
    """frames/s of the CPU port over `n_workers` processes (one per host core), the reference's
    own parallel style (recipes/zrc2019/steps/aud_gnu_parallel.sh:73-86)."""
    import multiprocessing as mp
    n_workers = n_workers or os.cpu_count() or 1
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    ctx = mp.get_context('spawn')
    t0 = time.perf_counter()
    with ctx.Pool(n_workers) as pool:
        res = pool.map(_cpu_worker, [(c, 1000 + i, utts_per_worker) for i in range(n_workers)])
    wall = max(r[1] for r in res)        # workers run concurrently; the slowest one ends the job
    frames = sum(r[0] for r in res)
    return frames / wall, n_workers, frames, time.perf_counter() - t0


def run_reference(args, c):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    upw = 48 if c['n_comp'] == 1 else 1          # ~5-10 s of CPU work per step on 16 cores
    vals = []
    for _ in range(args.warmup):
        cpu_reference_throughput(c, 1)
    for _ in range(args.steps):
        v, cores, frames, _ = cpu_reference_throughput(c, upw)
        vals.append((v, frames))
    tot_frames = sum(f for _, f in vals)
    tot_time = sum(f / v for v, f in vals)
    value = tot_frames / tot_time
    sample = f'{cores} worker processes x {upw} utterance(s) x {c["n_frames"]} frames per step, E-step + accumulate'
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * tot_time / max(args.steps, 1),
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': workload_name(args.config, c), 'sample': sample},
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------

def run_gpu(args, c):
    import torch
    import torch.distributed as dist
    from beer_b200 import ops, synthetic
    from beer_b200.engine import EmissionParams, Utterances, VBEngine, WeightGroup

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus:
        raise SystemExit(f'--gpus {args.gpus} needs {args.gpus} ranks (torchrun), got WORLD_SIZE={world}')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        from beer_b200.engine import bind_to_gpu_cpus
        bind_to_gpu_cpus(local_rank)      # pinned feature buffers next to this rank's GPU (e2e leg)
    if world > 1:
        # NCCL announces its version on stdout at the first collective: keep stdout to the one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            os.dup2(saved, 1)
            os.close(saved)
    ops.require_cuda()

    K = c['n_units'] * c['n_states']
    C = c['n_comp']
    M, D, T, U = K * C, c['dim'], c['n_frames'], c['n_utts']
    graph, _, _ = synthetic.phone_loop_graph(c['n_units'], c['n_states'])
    plan = ops.GraphPlan(graph.init_log_probs.numpy(), graph.final_log_probs.numpy(),
                         graph.trans_log_probs.numpy(), graph.pdf_id_mapping, n_pdfs=K)
    gen = torch.Generator().manual_seed(7)
    means = 2.0 * torch.randn(K, D, generator=gen)
    if c.get('aligned'):
        X, paths = synthetic.sample_utterances(graph, means, U, T, seed=100 + rank, device=dev, return_paths=True)
        plan = ops.ChainBatch.from_arrays(*synthetic.alignment_chains(paths, c['n_states']), device=dev)
        del paths
    else:
        X = synthetic.sample_utterances(graph, means, U, T, seed=100 + rank, device=dev)
    utts = Utterances(X, [T] * U)

    def make_engine(utts=utts, chunk_frames=args.chunk_frames, use_graph=False):
        prior, post = synthetic.initial_normal_gamma(M, D, seed=2, device=dev)
        groups, comp_off = (), None
        if C > 1:
            conc = torch.full((K, C), 1.0 / C, device=dev)
            groups = (WeightGroup(0, K, C, conc.clone(), conc.clone()),)
            comp_off = np.arange(K + 1) * C
        em = EmissionParams(prior, post, comp_off=comp_off, weight_groups=groups)
        return VBEngine(em, plan, utts, datasize=float(world * U * T), chunk_frames=chunk_frames,
                        distributed=world > 1, use_graph=use_graph, viterbi=args.viterbi)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    eng = make_engine(use_graph=not args.no_graph)
    elbos = []
    for _ in range(args.warmup):
        elbos.append(eng.step().clone())
    barrier()
    eng.gpu_launches = 0
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_wall = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.nvtx.range_push('bench_timed')     # ncu --nvtx --nvtx-include "bench_timed/" selects these launches
    ev0.record()
    for _ in range(args.steps):
        elbos.append(eng.step().clone())     # the graph's ELBO buffer is overwritten by the next step
    ev1.record()
    torch.cuda.nvtx.range_pop()
    barrier()
    wall = time.perf_counter() - t_wall
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    launches = eng.gpu_launches
    # per-stage kernel durations: the same iteration launched eagerly with CUDA events around the stages
    eng.profile = {}
    for _ in range(3):
        eng.step()
    torch.cuda.synchronize()
    stage_ms = {k: float(np.mean([a.elapsed_time(b) for a, b in v])) for k, v in eng.profile.items()}
    eng.profile = None
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if ms < 400.0:
        # the timed region is shorter than a few nvidia-smi sampling periods: sample the same step loop,
        # untimed, for ~0.6 s right behind it (same number of extra steps on every rank)
        n_probe = int(600.0 / max(ms / args.steps, 1e-3)) + 1
        probe = ClockSampler(local_rank)
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
    value = frames_per_step * args.steps / (ms * 1e-3)
    elbo_pf = [float(eng.elbo_per_frame(e).item()) for e in elbos]
    launches_step = launches // max(args.steps, 1)

    # ---- end to end: features in pinned host memory, H2D every step, ELBO read back --------
    host_X = torch.empty(X.shape, dtype=torch.float32, pin_memory=True)
    host_X.copy_(X)
    del eng
    # features stay in pinned host memory: the engine streams them in chunks of whole utterances (the H2D
    # copy of chunk i+1 under the kernels of chunk i) and the ELBO is read back to the host every step
    eng2 = make_engine(Utterances(host_X, [T] * U), chunk_frames=args.e2e_chunk_frames or max(T, U * T // 8))
    n_e2e = max(1, min(args.steps, 5))

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

    if rank == 0:
        peak, peak_src = measured_peaks()
        # dominant kernel and its algorithmic bytes per frame (DESIGN.md "Kernels and rooflines")
        # B_alg = 8D + 16K per frame = KA (read X, write llh) + KB (read llh once more, write and
        # read alpha) + KC (read X); posteriors and responsibilities count as on-chip in the algorithmic figure
        alg = {'KA_emission_llh': 4 * D + 4 * K, 'KB_forward_backward': 12 * K, 'KC_accumulate': 4 * D}
        traffic_pf = {}
        tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
        if os.path.exists(tpath):        # dram bytes per frame per launch from the committed ncu capture
            with open(tpath) as f:
                traffic_pf = json.load(f).get(args.config, {})
        dom = max((k for k in stage_ms if k in alg), key=lambda k: stage_ms[k], default=None)
        roofline = None
        if dom is not None:
            achieved = alg[dom] * U * T / (stage_ms[dom] * 1e-3) / 1e9
            roofline = {'bound': 'hbm', 'kernel': dom, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                        'frac': achieved / peak,
                        'traffic': (traffic_pf[dom] * U * T if dom in traffic_pf else None),
                        'traffic_unit': 'bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)',
                        'peak_source': peak_src,
                        'alg_bytes_per_frame': alg[dom], 'launch_ms': stage_ms[dom],
                        'step_alg_bytes_per_frame': 8 * D + 16 * K,
                        'step_frac': (8 * D + 16 * K) * U * T / (ms / args.steps * 1e-3) / 1e9 / peak,
                        'stage_ms': stage_ms}
            # the whole step against its three rooflines, per GPU: HBM (B_alg = 8D + 16K bytes / frame),
            # tensor pipe (F_alg = 4 Q M flop / frame at 3xTF32 = bf16 / 6) and the scan's special-function unit
            # (S_alg = 2 nnz(A) log-add-exp terms / frame, phone loop nnz = 2K - P + P^2, against 148 SMs x 16 MUFU
            # lanes x the measured SM clock)
            fps = U * T / (ms / args.steps * 1e-3)
            bf16, bf16_src = measured_bf16_tflops()
            Pn = c['n_units']
            nnz = 2 * K - Pn + Pn * Pn
            mhz = (clocks or {}).get('sm_mhz') or 1965.0
            roofline['step_fractions'] = {
                'hbm': fps * (8 * D + 16 * K) / 1e9 / peak,
                'tensor_3xtf32': fps * 4 * (2 * D + 2) * M / 1e12 / (bf16 / 6.0),
                'scan_mufu': fps * 2 * nnz / (148 * 16 * mhz * 1e6),
                'tensor_peak_tflops': bf16 / 6.0, 'tensor_peak_source': bf16_src + ' bf16 / 6'}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            upw = 96 if C == 1 else 1        # ~10 s of CPU work
            v, cores, frames, took = cpu_reference_throughput(c, upw)
            cpu = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                   'sample': f'{cores} worker processes x {upw} utterance(s) x {T} frames, numpy port of the '
                             f'reference E-step + accumulate ({took:.1f} s)'}
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                'config': {'workload': workload_name(args.config, c), 'l2': 'inputs larger than L2 '
                           f'({X.numel() * 4 / 2**20:.0f} MiB of features per GPU, no flush needed)',
                           'chunk_frames': args.chunk_frames, 'parallelism': f'dp{world} (utterances sharded, '
                           'one all-reduce of the statistics per step)'},
                'clocks': clocks, 'wall_s_timed_region': wall,
                'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(world * X.numel() * 4),
                        'd2h_bytes_per_step': 8 * world, 'steps': n_e2e},
                'gpu_launches': launches, 'roofline': roofline, 'cpu_baseline': cpu,
                'elbo_per_frame': {'first': elbo_pf[0], 'last': elbo_pf[-1]}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--config', default='cfg2')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--chunk-frames', type=int, default=None)
    ap.add_argument('--e2e-chunk-frames', type=int, default=None)
    ap.add_argument('--n-utts', type=int, default=None, help='override utterances per GPU (debug)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='launch every kernel eagerly (no CUDA graph)')
    ap.add_argument('--viterbi', action='store_true', help='Viterbi training (one-hot posteriors of the best path) '
                    'instead of forward-backward; a secondary workload, not the BASELINE metric')
    args = ap.parse_args()
    from beer_b200.synthetic import CONFIGS
    c = dict(CONFIGS[args.config])
    if args.n_utts:
        c['n_utts'] = args.n_utts
    if args.viterbi:
        c['viterbi'] = True
    if args.impl == 'reference':
        run_reference(args, c)
    else:
        run_gpu(args, c)


if __name__ == '__main__':
    main()
