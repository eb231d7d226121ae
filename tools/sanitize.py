"""Small end-to-end pass over every kernel family, meant to run under compute-sanitizer:

    compute-sanitizer --tool memcheck python tools/sanitize.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beer_b200 import features, ops, synthetic  # noqa: E402
from beer_b200.engine import EmissionParams, Utterances, VBEngine, WeightGroup  # noqa: E402

dev = torch.device('cuda', 0)


def run(P, S, C, D, lens, tag):
    K, M = P * S, P * S * C
    graph, _, _ = synthetic.phone_loop_graph(P, S)
    plan = ops.GraphPlan(graph.init_log_probs.numpy(), graph.final_log_probs.numpy(), graph.trans_log_probs.numpy(),
                         graph.pdf_id_mapping, n_pdfs=K)
    means = 2.0 * torch.randn(K, D, generator=torch.Generator().manual_seed(0))
    full = synthetic.sample_utterances(graph, means, len(lens), max(lens), seed=1, device=dev).reshape(len(lens), max(lens), D)
    X = torch.cat([full[i, :n] for i, n in enumerate(lens)])
    prior, post = synthetic.initial_normal_gamma(M, D, seed=2, device=dev)
    groups, comp_off = (), None
    if C > 1:
        conc = torch.full((K, C), 1.0 / C, device=dev)
        groups = (WeightGroup(0, K, C, conc.clone(), conc.clone()),)
        comp_off = np.arange(K + 1) * C
    em = EmissionParams(prior, post, comp_off=comp_off, weight_groups=groups)
    eng = VBEngine(em, plan, Utterances(X, lens), datasize=float(sum(lens)), distributed=False)
    vals = [float(eng.step().item()) for _ in range(2)]
    off = torch.as_tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int64, device=dev)
    pdf, _, fref = em.llh(X, *em.refresh(), torch.empty(len(X), em.Kp, device=dev), None if C == 1 else
                          torch.empty(len(X), M, device=dev), torch.empty(len(X), device=dev))
    path = ops.hmm_viterbi(plan, pdf, off)
    counts = torch.zeros(max(plan.n_units, 1), dtype=torch.float64, device=dev)
    if plan.n_units and plan.info['states_per_lane'] <= 16:
        ops.hmm_forward_backward(plan, pdf, fref, off, want_state_post=True, want_frame_llh=True, want_logz=True,
                                 unit_counts=counts)
    torch.cuda.synchronize()
    print(f'{tag}: elbo {vals}, tc={em.use_tc}, path[:5]={path[:5].tolist()}, units={plan.n_units}')


run(25, 4, 1, 40, [200, 37, 1, 129, 64], 'cfg2-like (tcgen05 KA/KC, LR scan)')
run(6, 3, 1, 40, [50, 33], 'S=3 units (LR scalar rows)')
run(16, 4, 8, 40, [130, 77, 40], 'mixtures C=8 (tcgen05, streamed weight image)')
run(130, 4, 1, 40, [60, 35], 'many units (block scan), 5 Gaussian tiles')
run(5, 4, 2, 12, [40, 41], 'SIMT kernels (D=12)')
run(40, 4, 8, 40, [300, 61, 129], 'mixtures C=8, M=1280 (fp16-split kernels, 8 weight chunks)')
run(250, 4, 8, 40, [300, 61, 129, 700], 'BASELINE configs[2] shape: 1000 pdfs x 8 (fp16-split kernels, 8-warp scan)')
run(12, 4, 4, 20, [90, 64, 1], 'mixtures C=4, D=20 (fp16-split kernels, single resident weight chunk)')
os.environ['BEER_B200_NO_MIX16'] = '1'
run(40, 4, 8, 40, [300, 61, 129], 'mixtures C=8, M=1280 (3xTF32 kernels: TMA tensor-map stores in KA, cp.async raw ring in KC)')
del os.environ['BEER_B200_NO_MIX16']


def run_gmm():
    """A plain mixture of 64 Gaussians (VBEngine without a graph: emission -> frame softmax -> statistics)."""
    C, D, N = 64, 40, 700
    X = synthetic.sample_gmm_frames(N, D, seed=3, device=dev)
    prior, post = synthetic.initial_normal_gamma(C, D, seed=2, device=dev)
    conc = torch.full((1, C), 1.0 / C, device=dev)
    em = EmissionParams(prior, post, comp_off=np.array([0, C]), weight_groups=(WeightGroup(0, 1, C, conc.clone(), conc.clone()),))
    eng = VBEngine(em, None, Utterances(X, [N]), datasize=float(N), distributed=False)
    print('gmm: elbo', [float(eng.step().item()) for _ in range(2)])


run_gmm()


def run_chains():
    """Per-utterance alignment chains of three length classes + the transition-posterior kernel."""
    rng = np.random.default_rng(1)
    Kp, D = 60, 40
    shapes = [(7, 40), (130, 150), (260, 300), (1, 3)]
    offs, pdf = [0], []
    for L, _ in shapes:
        pdf.append(rng.integers(0, Kp, L))
        offs.append(offs[-1] + L)
    n = offs[-1]
    chains = ops.ChainBatch.from_arrays(offs, np.concatenate(pdf), np.full(n, np.log(.7)), np.full(n, np.log(.3)),
                                        np.zeros(len(shapes)), dev)
    lens = [T for _, T in shapes]
    prior, post = synthetic.initial_normal_gamma(Kp, D, seed=3, device=dev)
    X = torch.randn(sum(lens), D, generator=torch.Generator().manual_seed(4)).to(dev)
    eng = VBEngine(EmissionParams(prior, post), chains, Utterances(X, lens), datasize=float(sum(lens)), distributed=False)
    vals = [float(eng.step().item()) for _ in range(2)]
    off = torch.as_tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int64, device=dev)
    llh = torch.randn(sum(lens), Kp, generator=torch.Generator().manual_seed(5)).to(dev)
    r = ops.hmm_forward_backward_chains(chains, llh, None, off, want_state_post=True, want_frame_llh=True, want_logz=True)
    graph, starts, ends = synthetic.phone_loop_graph(5, 3)
    K = graph.n_states
    plan = ops.GraphPlan(graph.init_log_probs.numpy(), graph.final_log_probs.numpy(), graph.trans_log_probs.numpy(),
                         graph.pdf_id_mapping, n_pdfs=K)
    off2 = torch.tensor([0, 50, 81], dtype=torch.int64, device=dev)
    llh2 = torch.randn(81, K, generator=torch.Generator().manual_seed(6)).to(dev)
    fb = ops.hmm_forward_backward(plan, llh2, None, off2, want_state_post=True, want_pdf_post=False)
    init = graph.init_log_probs.float().to(dev).contiguous()
    trans = graph.trans_log_probs.float().to(dev).contiguous()
    xi = ops.hmm_transition_posteriors(llh2, fb['state_post'], off2, init, trans)
    blk = ops.hmm_transition_posteriors(llh2, fb['state_post'], off2, init, trans,
                                        rows=torch.as_tensor(ends, dtype=torch.int32, device=dev),
                                        cols=torch.as_tensor(starts, dtype=torch.int32, device=dev))
    torch.cuda.synchronize()
    print(f'chains: elbo {vals}, logz {r["utt_logz"].tolist()}, xi {tuple(xi.shape)} sums to '
          f'{float(xi.sum()):.3f}, block {tuple(blk.shape)}')


run_chains()
sig = (np.random.default_rng(0).standard_normal(16000) * 1000).astype(np.int16)
fb = features.fbank(sig, nfilters=40)
print('fbank', tuple(fb.shape), tuple(features.add_deltas(fb).shape))
print('mspec', tuple(features.short_term_mspec(sig)[0].shape))
torch.cuda.synchronize()
print('sanitize pass done')
