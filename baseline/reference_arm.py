"""The UNMODIFIED reference (beer-asr/beer) timed and queried through its own public API.

Test / benchmark infrastructure only: `bench.py --impl reference`, the `cpu_baseline` leg and the in-bench ELBO
check import this module; nothing under `beer_b200/` does.  The reference is looked up in `baseline/_ref/`
(installed there by `baseline/install_reference.sh`: git-ignored, travels to the GPU box with the snapshot) and then in
`/root/reference` (the build container).  When neither exists, `find_reference()` returns None and the callers fall back
to the numpy port `oracle/beer_oracle.py`.

What is run is the loop of `beer hmm accumulate` + `beer hmm update` (beer/cli/subcommands/hmm/accumulate.py:37-63,
update.py:39-72):

    optim.init_step()
    elbo = beer.evidence_lower_bound(datasize=N)
    for X_u in shard:  elbo += beer.evidence_lower_bound(model, X_u, inference_graph=graph, datasize=N)
    elbo.backward(); optim.step()
"""
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def find_reference():
    for path in (os.path.join(HERE, '_ref'), '/root/reference'):
        if os.path.isdir(os.path.join(path, 'beer', 'models')):
            return path
    return None


def _import_reference():
    path = find_reference()
    if path is None:
        raise ImportError('the reference is neither in baseline/_ref nor in /root/reference')
    if path not in sys.path:
        sys.path.insert(0, path)
    warnings.filterwarnings('ignore')
    import beer
    if not hasattr(beer, 'evidence_lower_bound') or 'beer_b200' in (getattr(beer, '__file__', '') or ''):
        raise ImportError('`import beer` did not resolve to the reference')
    return beer


def _unit_graph(beer, n_states, first_pdf, self_loop=0.75):
    g = beer.graph.Graph()
    sts = [g.add_state(pdf_id=None)]
    for i in range(n_states):
        sts.append(g.add_state(pdf_id=first_pdf + i))
    sts.append(g.add_state(pdf_id=None))
    g.start_state, g.end_state = sts[0], sts[-1]
    g.add_arc(sts[0], sts[1], 1.0)
    for a in range(1, n_states + 1):
        g.add_arc(sts[a], sts[a], self_loop)
        g.add_arc(sts[a], sts[a + 1], 1 - self_loop)
    return g


def phone_loop(beer, n_units, n_states):
    """The decoding graph of mkphoneloopgraph.py:28-77 + mkdecodegraph.py:50-58, built with the reference's Graph."""
    g = beer.graph.Graph()
    g.start_state = g.add_state()
    g.end_state = g.add_state()
    pivot = g.add_state()
    us = [g.add_state() for _ in range(n_units)]
    g.add_arc(g.start_state, pivot)
    g.add_arc(pivot, g.end_state)
    for s in us:
        g.add_arc(pivot, s)
        g.add_arc(s, pivot)
    g.normalize()
    for i, s in enumerate(us):
        g.replace_state(s, _unit_graph(beer, n_states, i * n_states))
    g.normalize()
    return g.compile()


def build_model(cfg, seed=2, double=False):
    """HMM over NormalSet (C = 1) or MixtureSet(NormalSet) (C > 1) of the bench configuration -- or, for a `gmm`
    configuration, a plain Mixture of n_comp Gaussians -- created by the reference's own constructors (mean 0, cov 1,
    prior_strength 1, noise_std 1: SURVEY 8d).  Returns (beer, model, NormalSet, MixtureSet / Mixture or None, graph)."""
    import torch
    beer = _import_reference()
    torch.manual_seed(seed)
    C, D = cfg['n_comp'], cfg['dim']
    if cfg.get('gmm'):
        ns = beer.NormalSet.create(torch.zeros(D), torch.ones(D), size=C, prior_strength=1., noise_std=1.,
                                   cov_type='diagonal')
        model = beer.Mixture.create(ns, prior_strength=1.)
        if double:
            model = model.double()
        return beer, model, ns, model, None
    K = cfg['n_units'] * cfg['n_states']
    cg = phone_loop(beer, cfg['n_units'], cfg['n_states'])
    ns = beer.NormalSet.create(torch.zeros(D), torch.ones(D), size=K * C, prior_strength=1., noise_std=1.,
                               cov_type='diagonal')
    modelset = ns if C == 1 else beer.MixtureSet.create(K, ns, prior_strength=1.)
    hmm = beer.HMM.create(cg, modelset)
    if double:
        hmm = hmm.double()
    return beer, hmm, ns, (modelset if C > 1 else None), hmm.graph


def model_arrays(ns, ms):
    """Standard parameters of the reference model as numpy fp64 (copy them into the engine)."""
    def std(dist):
        p = dist.params
        return tuple(t.detach().double().numpy().copy() for t in (p.mean, p.scale, p.shape, p.rates))
    out = dict(ng_prior=std(ns.means_precisions.prior), ng_post=std(ns.means_precisions.posterior))
    if ms is not None:
        w = ms.categoricalset.weights if hasattr(ms, 'categoricalset') else ms.categorical.weights     # MixtureSet / Mixture
        out['dir_prior'] = w.prior.params.concentrations.detach().double().numpy().copy()
        out['dir_post'] = w.posterior.params.concentrations.detach().double().numpy().copy()
    return out


def vb_iteration(cfg, utts, threads=1, double=False, update=True, seed=2, want_model=False, datasize=None):
    """One accumulate + update pass of the reference over `utts` (list of [T, D] float arrays).
    Returns dict(frames, seconds (E-step + backward + optimizer step, model construction excluded), elbo, model)."""
    import torch
    torch.set_num_threads(max(1, int(threads)))
    beer, hmm, ns, ms, graph = build_model(cfg, seed=seed, double=double)
    kw = {} if graph is None else {'inference_graph': graph}
    arrays = model_arrays(ns, ms) if want_model else None
    dtype = torch.float64 if double else torch.float32
    data = [torch.from_numpy(np.asarray(u)).to(dtype) for u in utts]
    N = float(sum(len(u) for u in utts)) if datasize is None else float(datasize)
    optim = beer.VBConjugateOptimizer(hmm.mean_field_factorization(), lrate=1.)
    t0 = time.perf_counter()
    optim.init_step()
    elbo = beer.evidence_lower_bound(datasize=N)
    per_utt = []
    for X in data:
        e = beer.evidence_lower_bound(hmm, X, datasize=N, **kw)
        per_utt.append(float(e))
        elbo += e
    if update:
        elbo.backward()
        optim.step()
    seconds = time.perf_counter() - t0
    return dict(frames=int(sum(len(u) for u in utts)), seconds=seconds, elbo=float(elbo), per_utt=per_utt, model=arrays)


def _worker(args):
    cfg, utts, threads, double, want_model, datasize = args[:6]
    update = args[6] if len(args) > 6 else True
    os.environ.setdefault('OMP_NUM_THREADS', str(threads))
    return vb_iteration(cfg, utts, threads=threads, double=double, want_model=want_model, datasize=datasize,
                        update=update)


def parallel_throughput(cfg, shards, double=False, datasize=None):
    """The reference's own parallel style (recipes/zrc2019/steps/aud_gnu_parallel.sh:73-86): one single-threaded
    worker process per shard, all shards at once; the slowest worker ends the job.  -> (frames/s, results)."""
    import multiprocessing as mp
    ctx = mp.get_context('spawn')
    with ctx.Pool(len(shards)) as pool:
        res = pool.map(_worker, [(cfg, s, 1, double, i == 0, datasize) for i, s in enumerate(shards)])
    wall = max(r['seconds'] for r in res)
    frames = sum(r['frames'] for r in res)
    return frames / wall, res


if __name__ == '__main__':
    print(find_reference())
