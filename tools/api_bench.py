#!/usr/bin/env python
"""Cost of the per-utterance MODEL API path (what a drop-in user of `beer.evidence_lower_bound` hits: one call per
utterance, accumulate.py:39-59), next to the same utterances as ONE `Utterances` batch and to the batched engine.

    python tools/api_bench.py > gpurun_out/api_bench.json
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def build(P, S, C, D, dev):
    import beer_b200 as beer
    from beer_b200 import synthetic
    K = P * S
    graph, start_pdf, end_pdf = synthetic.phone_loop_graph(P, S)
    ns = beer.NormalSet.create(torch.zeros(D, device=dev), torch.ones(D, device=dev), size=K * C, prior_strength=1.,
                               noise_std=1., cov_type='diagonal')
    emissions = ns if C == 1 else beer.MixtureSet.create(K, ns, prior_strength=1.)
    return beer.HMM.create(graph, emissions), graph


def run(tag, P, S, C, D, T, n_utts, dev):
    import beer_b200 as beer
    from beer_b200 import synthetic
    model, graph = build(P, S, C, D, dev)
    means = 2.0 * torch.randn(P * S, D, generator=torch.Generator().manual_seed(0))
    X = synthetic.sample_utterances(graph, means, n_utts, T, seed=1, device=dev).reshape(n_utts, T, D)
    N = float(n_utts * T)

    def per_utterance():
        elbo = beer.evidence_lower_bound(datasize=N)
        for u in range(n_utts):
            elbo += beer.evidence_lower_bound(model, X[u], datasize=N, inference_graph=graph)
        return float(elbo)

    def one_batch():
        utts = beer.Utterances(X.reshape(-1, D), [T] * n_utts)
        return float(beer.evidence_lower_bound(model, utts, datasize=N, inference_graph=graph))

    out = {'case': tag, 'utterances': n_utts, 'frames_per_utterance': T}
    for name, fn in (('per_utterance_calls', per_utterance), ('one_batched_call', one_batch)):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            val = fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        out[name] = {'ms_per_utterance': 1e3 * dt / n_utts, 'frames_per_s': N / dt, 'elbo': val}
    return out


if __name__ == '__main__':
    dev = torch.device('cuda', 0)
    res = [run('cfg2 shape: HMM, 100 states x 1 Gaussian, D = 40', 25, 4, 1, 40, 1000, 64, dev),
           run('cfg3 shape: HMM over MixtureSet, 1000 states x 8 Gaussians, D = 40', 250, 4, 8, 40, 1000, 16, dev)]
    print(json.dumps(res, indent=1))
