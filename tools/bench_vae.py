#!/usr/bin/env python
"""BASELINE configs[3] as a measured line: one training step of an HMM-VAE (beer/models/vae.py:63-89, examples/HMM-VAE)
on a batch of utterances -- encoder MLP 40 -> 128, latent 64-d, decoder 64 -> 128 -> 40, prior = HMM of 100 states
(25 units x 4) with diagonal Gaussians over the latent space, one sample per frame.

    python tools/bench_vae.py [--utts 512] [--steps 10] > gpurun_out/bench_vae.json

A step = encoder + reparameterised sample (torch), the prior's E-step on the samples (emission -> forward-backward on
the kernels of the hot path), decoder likelihood (torch), backward -- the gradient of the prior term w.r.t. the samples
is ONE tcgen05 kernel (csrc/emission_bwd.cu) --, Adam step of the networks, statistics + natural-gradient step of the
prior (the two optimisers of the reference's VAE recipe).  The networks are the user's torch modules (library GEMMs):
the number says what the hot-path kernels leave of the step, it is not the BASELINE metric.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class MLP(torch.nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.dim_in, self.dim_out = dim_in, dim_out
        self.net = torch.nn.Sequential(torch.nn.Linear(dim_in, dim_out), torch.nn.Tanh())

    def forward(self, x):
        return self.net(x)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--utts', type=int, default=512)
    ap.add_argument('--frames', type=int, default=1000)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    args = ap.parse_args()
    import beer_b200 as beer
    from beer_b200 import synthetic
    from beer_b200.vae import VAE
    dev = torch.device('cuda', 0)
    torch.manual_seed(0)
    P, S, D, DL = 25, 4, 40, 64
    graph, _, _ = synthetic.phone_loop_graph(P, S)
    ns = beer.NormalSet.create(torch.zeros(DL, device=dev), torch.ones(DL, device=dev), size=P * S, prior_strength=1.,
                               noise_std=1., cov_type='diagonal')
    prior = beer.HMM.create(graph, ns)
    vae = VAE(prior, MLP(D, 128), MLP(DL, 128))
    for name, module in vae.named_children():          # the networks to the GPU (the prior's parameters are there already)
        if name != 'prior':
            module.to(dev)
    means = 2.0 * torch.randn(P * S, D, generator=torch.Generator().manual_seed(0))
    X = synthetic.sample_utterances(graph, means, args.utts, args.frames, seed=1, device=dev)
    utts = beer.Utterances(X, [args.frames] * args.utts)
    N = float(len(utts))
    nets = [p for n, p in vae.named_parameters() if not n.startswith('prior.')]
    adam = torch.optim.Adam(nets, lr=1e-3)
    cjg = beer.VBConjugateOptimizer(vae.mean_field_factorization(), lrate=1.)

    def step():
        adam.zero_grad(set_to_none=True)
        cjg.init_step()
        stats = vae.sufficient_statistics(utts)
        val = vae.expected_log_likelihood(stats, inference_graph=graph)            # [N]: llh - (xent - ent)
        acc = vae.accumulate(stats)
        vae.clear_cache()
        loss = -val.sum() / N
        loss.backward()
        adam.step()
        for param, st in acc.items():                                            # objectives.py:98-106 at datasize = N
            param.store_stats(st)
        cjg.step()
        return loss

    for _ in range(args.warmup):
        first = step()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(args.steps):
        last = step()
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / args.steps
    # share of the hot-path kernels: the same step under the torch profiler's kernel table would be the evidence; here
    # the prior's E-step + gradient kernel are timed alone on the same latent samples
    with torch.no_grad():
        z = torch.randn(int(N), DL, device=dev)
    zs = beer.Utterances(z.requires_grad_(True), [args.frames] * args.utts)
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(args.steps):
        st = prior.sufficient_statistics(zs)
        v = prior.expected_log_likelihood(st, inference_graph=graph)
        prior.accumulate(st)
        prior.clear_cache()
        v.sum().backward()
        zs.X.grad = None
    ev[1].record()
    torch.cuda.synchronize()
    ms_prior = ev[0].elapsed_time(ev[1]) / args.steps
    print(json.dumps({
        'metric': 'HMM-VAE training frames/sec (BASELINE configs[3])', 'value': N / (ms * 1e-3), 'unit': 'frames/s',
        'ms_per_step': ms, 'n_gpus': 1, 'steps': args.steps, 'warmup': args.warmup,
        'config': {'workload': f'cfg4: HMM-VAE, 100-state HMM prior over a 64-d latent, encoder 40 -> 128 -> 64, decoder '
                               f'64 -> 128 -> 40, nsamples = 1, {args.utts} utterances x {args.frames} frames in one batch'},
        'prior_e_step_and_gradient_ms': ms_prior,
        'loss_per_frame': {'after_warmup': float(first), 'last': float(last)}}))


if __name__ == '__main__':
    main()
