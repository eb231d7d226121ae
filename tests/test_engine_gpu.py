"""GPU parity tests of the batched VB engine against the numpy oracle of the reference's
accumulate/update loop (beer/cli/subcommands/hmm/accumulate.py:37-63, update.py:37-62)."""
import numpy as np
import pytest
import torch

from oracle import beer_oracle as O

pytestmark = pytest.mark.gpu


def _host(t):
    m, k, a, b = (x.double().cpu().numpy() for x in t)
    return m, k[:, None], a[:, None], b


@pytest.mark.parametrize('C,chunk,graph', [(1, None, False), (1, 130, False), (3, None, False), (3, 200, False),
                                           (1, None, True), (3, 130, True)])
def test_vb_iterations_match_oracle(C, chunk, graph):
    from beer_b200 import ops, synthetic
    from beer_b200.engine import EmissionParams, Utterances, VBEngine, WeightGroup
    dev = torch.device('cuda', 0)
    P, S, D = 6, 3, 10
    K, M = P * S, P * S * C
    lens = [80, 41, 120, 64, 7, 99]
    graph, _, _ = synthetic.phone_loop_graph(P, S)
    plan = ops.GraphPlan(graph.init_log_probs.numpy(), graph.final_log_probs.numpy(),
                         graph.trans_log_probs.numpy(), graph.pdf_id_mapping, n_pdfs=K)
    means = 2.0 * torch.randn(K, D, generator=torch.Generator().manual_seed(0))
    full = synthetic.sample_utterances(graph, means, len(lens), max(lens), seed=1, device=dev)
    full = full.reshape(len(lens), max(lens), D)
    utts_dev = [full[i, :n] for i, n in enumerate(lens)]
    X = torch.cat(utts_dev)
    prior, post = synthetic.initial_normal_gamma(M, D, seed=2, device=dev)
    groups, comp_off, dprior, dpost = (), None, None, None
    if C > 1:
        conc = torch.full((K, C), 1.0 / C, device=dev)
        groups = (WeightGroup(0, K, C, conc.clone(), conc.clone()),)
        comp_off = np.arange(K + 1) * C
        dprior, dpost = conc.double().cpu().numpy(), conc.double().cpu().numpy()
    em = EmissionParams(prior, post, comp_off=comp_off, weight_groups=groups)
    N = sum(lens)
    eng = VBEngine(em, plan, Utterances(X, lens), datasize=float(N), chunk_frames=chunk, distributed=False,
                   use_graph=graph)
    ng_prior, ng_post = _host(prior), _host(post)
    og = (graph.init_log_probs.double().numpy(), graph.final_log_probs.double().numpy(),
          graph.trans_log_probs.double().numpy(), graph.pdf_id_mapping)
    utts = [u.double().cpu().numpy() for u in utts_dev]
    for it in range(4):
        want, ng_post, dpost, info = O.vb_iteration_hmm(utts, ng_prior, ng_post, dprior, dpost, og)
        got = float(eng.step().item())
        assert abs(got - want) <= 1e-5 * abs(want), (it, got, want)
        acc = eng.acc.cpu().numpy()
        assert np.abs(acc - info['acc_normal']).max() <= 3e-5 * np.abs(info['acc_normal']).max()
    for g, w in zip(_host(em.post), ng_post):
        np.testing.assert_allclose(g, w, rtol=2e-4, atol=2e-4)
    if C > 1:
        np.testing.assert_allclose(groups[0].post.double().cpu().numpy(), dpost, rtol=2e-4, atol=1e-5)


def test_shards_sum_to_single_process():
    """Statistics are plain sums over utterances: two "ranks" each holding half of the
    utterances produce flat buffers whose sum equals the single-rank buffer (what the NCCL
    all-reduce computes; objectives.py:78-90)."""
    from beer_b200 import ops, synthetic
    from beer_b200.engine import EmissionParams, Utterances, VBEngine
    dev = torch.device('cuda', 0)
    P, S, D, T, U = 5, 4, 12, 70, 8
    K = P * S
    graph, _, _ = synthetic.phone_loop_graph(P, S)
    plan = ops.GraphPlan(graph.init_log_probs.numpy(), graph.final_log_probs.numpy(),
                         graph.trans_log_probs.numpy(), graph.pdf_id_mapping, n_pdfs=K)
    means = 2.0 * torch.randn(K, D, generator=torch.Generator().manual_seed(0))
    X = synthetic.sample_utterances(graph, means, U, T, seed=1, device=dev)

    def engine(Xs, n):
        prior, post = synthetic.initial_normal_gamma(K, D, seed=2, device=dev)
        return VBEngine(EmissionParams(prior, post), plan, Utterances(Xs, [T] * n), datasize=float(U * T),
                        distributed=False)

    whole = engine(X, U)
    whole._gframes = float(U * T)
    whole.e_step()
    halves = [engine(X[:U // 2 * T], U // 2), engine(X[U // 2 * T:], U - U // 2)]
    tot = torch.zeros_like(whole.flat)
    for h in halves:
        h._gframes = float(U * T)
        h.e_step()
        tot += h.flat
    np.testing.assert_allclose(tot.cpu().numpy(), whole.flat.cpu().numpy(), rtol=2e-6, atol=1e-6)


def test_smoke_entry():
    import __graft_entry__
    __graft_entry__.smoke()
