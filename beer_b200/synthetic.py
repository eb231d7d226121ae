"""Synthetic HMM-GMM workloads named by BASELINE.json: a phone-loop
decoding graph of P left-to-right units and 40-d "fbank" frames sampled from it.

Used by bench.py, __graft_entry__.smoke() and the tests; everything is seeded."""
import numpy as np
import torch

from .graph import Graph

__all__ = ['unit_graph', 'phone_loop_graph', 'sample_utterances', 'sample_gmm_frames', 'alignment_chains',
           'initial_normal_gamma', 'CONFIGS']

# name -> (units, states per unit, Gaussians per state, dim, frames per utterance, utterances per GPU)
CONFIGS = {
    'cfg2': dict(n_units=25, n_states=4, n_comp=1, dim=40, n_frames=1000, n_utts=4096),
    'cfg3': dict(n_units=250, n_states=4, n_comp=8, dim=40, n_frames=1000, n_utts=1250),
    # cfg2 trained with one alignment graph per utterance (`beer hmm accumulate --alis`): the unit sequence each
    # utterance was sampled through, as a left-to-right chain
    'cfg2ali': dict(n_units=25, n_states=4, n_comp=1, dim=40, n_frames=1000, n_utts=4096, aligned=True),
    # BASELINE configs[4]: the E-step + conjugate M-step of the 512-component diagonal GMM inside the GSM-GMM example
    # (no HMM, no forward-backward; the subspace SGD step of the GSM stays in PyTorch)
    'cfg5': dict(gmm=True, n_units=0, n_states=0, n_comp=512, dim=40, n_frames=1000, n_utts=1250),
}


def unit_graph(n_states, first_pdf, self_loop=0.75):
    """Left-to-right unit (recipes/aud/conf/hmm.yml:37-44 topology)."""
    g = Graph()
    states = [g.add_state(pdf_id=None)]
    states += [g.add_state(pdf_id=first_pdf + i) for i in range(n_states)]
    states.append(g.add_state(pdf_id=None))
    g.start_state, g.end_state = states[0], states[-1]
    g.add_arc(states[0], states[1], 1.0)
    for i in range(1, n_states + 1):
        g.add_arc(states[i], states[i], self_loop)
        g.add_arc(states[i], states[i + 1], 1 - self_loop)
    return g


def phone_loop_graph(n_units, n_states=4, self_loop=0.75):
    """Decoding graph start -> pivot -> {units} -> pivot -> end, every unit spliced in with
    replace_state, the construction of beer/cli/subcommands/hmm/mkphoneloopgraph.py:28-77 +
    mkdecodegraph.py:50-58.  Returns (CompiledGraph, start pdf ids, end pdf ids)."""
    g = Graph()
    g.start_state = g.add_state()
    g.end_state = g.add_state()
    pivot = g.add_state()
    placeholders = [g.add_state() for _ in range(n_units)]
    g.add_arc(g.start_state, pivot)
    g.add_arc(pivot, g.end_state)
    for s in placeholders:
        g.add_arc(pivot, s)
        g.add_arc(s, pivot)
    g.normalize()
    starts, ends = [], []
    for u, s in enumerate(placeholders):
        g.replace_state(s, unit_graph(n_states, u * n_states, self_loop))
        starts.append(u * n_states)
        ends.append((u + 1) * n_states - 1)
    g.normalize()
    return g.compile(), starts, ends


def sample_utterances(graph, means, n_utts, n_frames, seed, device='cpu', noise=1.0, return_paths=False):
    """Sample a state path per utterance from `graph` and emit x_t = mu[pdf(s_t)] + noise * eps.
    Everything is drawn on the HOST (numpy, inverse-CDF over the outgoing arcs of the current states, all utterances
    of a frame at once) and moved to `device` in one copy, so that no sampling kernel shows up next to the kernels
    under test.  Returns an [n_utts * n_frames, D] fp32 tensor on `device` (utterances back to back); with
    `return_paths` also the sampled state paths [n_utts, n_frames] (int64, on `device`)."""
    rng = np.random.default_rng(seed)
    init = np.exp(graph.init_log_probs.double().numpy())
    trans = np.exp(graph.trans_log_probs.double().numpy())
    init /= init.sum()
    trans /= trans.sum(axis=1, keepdims=True)
    K = len(init)
    # outgoing arcs of every state, padded to the largest out-degree: [K, deg] next states and cumulative probabilities
    deg = int((trans > 0).sum(axis=1).max())
    order = np.argsort(-trans, axis=1, kind='stable')[:, :deg]
    cum = np.cumsum(np.take_along_axis(trans, order, axis=1), axis=1)
    cum[:, -1] = 1.0
    pdf = np.asarray(graph.pdf_id_mapping, dtype=np.int64)
    means = np.asarray(means.cpu() if torch.is_tensor(means) else means, dtype=np.float32)
    D = means.shape[1]
    paths = np.empty((n_utts, n_frames), dtype=np.int64)
    state = np.minimum(np.searchsorted(np.cumsum(init), rng.random(n_utts)), K - 1)
    for t in range(n_frames):
        if t > 0:
            u = rng.random(n_utts)
            j = (cum[state] < u[:, None]).sum(axis=1)
            state = order[state, np.minimum(j, deg - 1)]
        paths[:, t] = state
    X = means[pdf[paths.reshape(-1)]]
    X += noise * rng.standard_normal(X.shape, dtype=np.float32)
    X = torch.from_numpy(X).to(device)
    return (X, torch.from_numpy(paths).to(device)) if return_paths else X


def sample_gmm_frames(n_frames, dim, seed, n_centres=32, spread=3.0, device='cpu'):
    """Frames of a GMM workload: `n_centres` cluster centres ~ N(0, spread^2 I), unit-variance noise (host RNG, one copy
    to `device`)."""
    rng = np.random.default_rng(seed)
    centres = spread * np.random.default_rng(12345).standard_normal((n_centres, dim)).astype(np.float32)
    X = centres[rng.integers(0, n_centres, n_frames)]
    X += rng.standard_normal(X.shape, dtype=np.float32)
    return torch.from_numpy(X).to(device)


def alignment_chains(paths, n_states, self_loop=0.75):
    """Alignment chains (mkaligraph.py:18-39 after compile) of the unit sequences the state paths of a phone loop
    of `n_states`-state units went through: flat arrays for ops.ChainBatch.from_arrays.  Unit u = states
    u*n_states .. u*n_states+n_states-1 = its pdf ids; every chain state has a self loop `self_loop`."""
    paths = np.asarray(paths.cpu() if torch.is_tensor(paths) else paths)
    offs, pdf = [0], []
    for p in paths:
        units = p // n_states
        # a new unit instance starts where the unit changes or the path re-enters a first state from a last one
        new = np.ones(len(p), dtype=bool)
        new[1:] = (units[1:] != units[:-1]) | ((p[1:] % n_states == 0) & (p[:-1] % n_states == n_states - 1))
        seq = units[new]
        ids = (seq[:, None] * n_states + np.arange(n_states)[None, :]).reshape(-1)
        pdf.append(ids)
        offs.append(offs[-1] + len(ids))
    pdf = np.concatenate(pdf).astype(np.int32)
    log_self = np.full(len(pdf), np.log(self_loop), dtype=np.float32)
    log_next = np.full(len(pdf), np.log(1 - self_loop), dtype=np.float32)
    return np.asarray(offs, dtype=np.int64), pdf, log_self, log_next, np.zeros(len(paths), dtype=np.float32)


def initial_normal_gamma(n_gauss, dim, seed, device='cpu', prior_strength=1.0, noise_std=1.0):
    """Prior / initial posterior of NormalSet.create(mean=0, cov=1, ...) for diagonal
    covariances (beer/models/normalset.py:42-54): returns two tuples
    (mean [M,D], scale [M], shape [M], rates [M,D])."""
    gen = torch.Generator(device='cpu')
    gen.manual_seed(seed)
    mean0 = torch.zeros(n_gauss, dim)
    noise = torch.randn(n_gauss, dim, generator=gen) * noise_std
    scale = torch.full((n_gauss,), float(prior_strength))
    shape = torch.full((n_gauss,), float(prior_strength))
    rates = torch.full((n_gauss, dim), float(prior_strength))
    prior = tuple(t.to(device).contiguous() for t in (mean0, scale, shape, rates))
    post = tuple(t.to(device).contiguous() for t in (mean0 + noise, scale.clone(), shape.clone(), rates.clone()))
    return prior, post
