"""Per-utterance alignment graphs (left-to-right chains, mkaligraph.py:18-39) in one launch:
beer_hmm_forward_backward_chains against the numpy oracle's forward-backward on the dense graph of
every utterance, the batched engine against the oracle's accumulate/update loop with one graph per
utterance (accumulate.py:47-57), and the model API with a list of inference graphs."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import beer_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def ali_graph(seq, n_states, self_loop=0.75):
    """Alignment graph of a unit sequence built like mkaligraph.py:18-39 with the oracle's Graph:
    start -> one placeholder per unit -> end, every placeholder replaced by its left-to-right HMM."""
    g = O.OracleGraph()
    g.start_state = g.add_state()
    last, holders = g.start_state, []
    for _ in seq:
        s = g.add_state()
        holders.append(s)
        g.add_arc(last, s)
        last = s
    g.end_state = g.add_state()
    g.add_arc(last, g.end_state)
    for s, unit in zip(holders, seq):
        u = O.OracleGraph()
        sts = [u.add_state(pdf_id=None)] + [u.add_state(pdf_id=unit * n_states + i) for i in range(n_states)]
        sts.append(u.add_state(pdf_id=None))
        u.start_state, u.end_state = sts[0], sts[-1]
        u.add_arc(sts[0], sts[1], 1.0)
        for a in range(1, n_states + 1):
            u.add_arc(sts[a], sts[a], self_loop)
            u.add_arc(sts[a], sts[a + 1], 1 - self_loop)
        g.replace_state(s, u)
    g.normalize()
    return g.compile()


def random_chain(rng, L, Kp):
    """A chain with random weights as (init, final, trans, map)."""
    loop = rng.uniform(0.2, 0.9, L)
    trans = np.full((L, L), -np.inf)
    trans[np.arange(L), np.arange(L)] = np.log(loop)
    trans[np.arange(L - 1), np.arange(1, L)] = np.log(1 - loop[:-1])
    init = np.full(L, -np.inf)
    init[0] = np.log(rng.uniform(0.5, 1.0))
    final = np.full(L, -np.inf)
    final[-1] = np.log(1 - loop[-1])
    return init, final, trans, rng.integers(0, Kp, L)


@pytest.mark.parametrize('scale,aligned', [(1.0, True), (0.6, True), (1.0, False)])
def test_chain_forward_backward_matches_oracle(scale, aligned):
    """`aligned`: the llhs favour one monotone path through every chain (what aligned training sees).  Otherwise they
    are i.i.d. noise of 6 nats per (frame, pdf): the prefix-best and the globally best paths then differ by hundreds of
    nats; log values relative to a per-FRAME maximum carried 2^-23 of that regret (3e-5 on posteriors), the per-lane
    integer offsets of the kernel keep both cases inside 1e-5."""
    from beer_b200 import ops
    ops.require_cuda()
    rng = np.random.default_rng(3)
    Kp = 52
    shapes = [(1, 9), (5, 5), (5, 40), (37, 64), (130, 200), (129, 131), (300, 333), (12, 1000)]   # (L, T), T >= L
    graphs = [random_chain(rng, L, Kp) for L, _ in shapes]
    lens = [T for _, T in shapes]
    N = sum(lens)
    llh = rng.standard_normal((N, Kp)) * 6 - 40
    fref = rng.standard_normal(N) * 3
    off = np.concatenate([[0], np.cumsum(lens)])
    if aligned:
        llh = rng.standard_normal((N, Kp)) * 2 - 40
        for u, ((L, T), gr) in enumerate(zip(shapes, graphs)):
            states = np.minimum(np.arange(T) * L // T, L - 1)
            llh[off[u] + np.arange(T), gr[3][states]] += 25.0
    tol = 1e-5
    chains = ops.ChainBatch(graphs, DEV)
    assert chains.max_len == 300 and chains.row_stride == 512
    r = ops.hmm_forward_backward_chains(chains, torch.as_tensor(llh, dtype=torch.float32, device=DEV),
                                        torch.as_tensor(fref, dtype=torch.float32, device=DEV),
                                        torch.as_tensor(off, dtype=torch.int64, device=DEV), scale=scale,
                                        want_state_post=True, want_frame_llh=True, want_logz=True)
    sp, pp = r['state_post'].double().cpu().numpy(), r['pdf_post'].double().cpu().numpy()
    fl, ue, lz = (r[k].double().cpu().numpy() for k in ('frame_exp_llh', 'utt_exp_llh', 'utt_logz'))
    l32 = llh.astype(np.float32).astype(np.float64)
    f32 = fref.astype(np.float32).astype(np.float64)
    for u, (init, final, trans, pmap) in enumerate(graphs):
        a, b = off[u], off[u + 1]
        pc = scale * l32[a:b][:, pmap]
        gam, _ = O.posteriors(pc, init, final, trans)
        L = len(pmap)
        assert np.abs(sp[a:b, :L] - gam).max() <= tol, u
        assert (sp[a:b, L:] == 0).all()
        want_pp = np.zeros((b - a, Kp))
        for j in range(L):
            want_pp[:, pmap[j]] += scale * gam[:, j]
        assert np.abs(pp[a:b] - want_pp).max() <= tol, u
        exp_llh = (pc * gam).sum(-1) + scale * f32[a:b]
        np.testing.assert_allclose(fl[a:b], exp_llh, rtol=tol, atol=100 * tol)
        np.testing.assert_allclose(ue[u], exp_llh.sum(), rtol=tol)
        la = O.forward(pc, init, trans)
        want_lz = O.logsumexp(la[-1] + final, axis=0) + scale * f32[a:b].sum()
        np.testing.assert_allclose(lz[u], want_lz, rtol=1e-5)


def test_chain_batch_rejects_other_graphs():
    from beer_b200 import ops
    gr, _, _ = O.phone_loop_graph(3, 2)
    with pytest.raises(ValueError):
        ops.ChainBatch([gr], DEV)


def _host(t):
    m, k, a, b = (x.double().cpu().numpy() for x in t)
    return m, k[:, None], a[:, None], b


@pytest.mark.parametrize('C,chunk', [(1, None), (3, 150)])
def test_engine_with_per_utterance_alignments(C, chunk):
    """Aligned training: every utterance has its own alignment graph; 3 VB iterations against the oracle."""
    from beer_b200 import ops, synthetic
    from beer_b200.engine import EmissionParams, Utterances, VBEngine, WeightGroup
    dev = torch.device('cuda', 0)
    P, S, D = 6, 3, 8
    K, M = P * S, P * S * C
    rng = np.random.default_rng(5)
    seqs = [[0, 3, 3, 1], [5], [2, 4, 0, 1, 5, 2, 2], [1, 0], [4, 4, 4, 3]]
    lens = [70, 9, 140, 33, 64]
    graphs = [ali_graph(s, S) for s in seqs]
    means = 2.0 * rng.standard_normal((K, D))
    utts = []
    for g, T in zip(graphs, lens):
        # a monotone pass through the chain, then noise around the state means
        L = len(g[3])
        states = np.minimum(np.arange(T) * L // T, L - 1)
        utts.append(means[np.asarray(g[3])[states]] + rng.standard_normal((T, D)))
    X = torch.as_tensor(np.concatenate(utts), dtype=torch.float32, device=dev)
    utts = [u.astype(np.float32).astype(np.float64) for u in utts]
    prior, post = synthetic.initial_normal_gamma(M, D, seed=2, device=dev)
    groups, comp_off, dprior, dpost = (), None, None, None
    if C > 1:
        conc = torch.full((K, C), 1.0 / C, device=dev)
        groups = (WeightGroup(0, K, C, conc.clone(), conc.clone()),)
        comp_off = np.arange(K + 1) * C
        dprior, dpost = conc.double().cpu().numpy(), conc.double().cpu().numpy()
    em = EmissionParams(prior, post, comp_off=comp_off, weight_groups=groups)
    N = sum(lens)
    eng = VBEngine(em, ops.ChainBatch(graphs, dev), Utterances(X, lens), datasize=float(N), chunk_frames=chunk,
                   distributed=False)
    ng_prior, ng_post = _host(prior), _host(post)
    og = [tuple(np.asarray(a, dtype=np.float64) if i < 3 else a for i, a in enumerate(g)) for g in graphs]
    for it in range(3):
        want, ng_post, dpost, info = O.vb_iteration_hmm(utts, ng_prior, ng_post, dprior, dpost, None, graphs=og)
        got = float(eng.step().item())
        assert abs(got - want) <= 1e-5 * abs(want), (it, got, want)
        acc = eng.acc.cpu().numpy()
        assert np.abs(acc - info['acc_normal']).max() <= 3e-5 * np.abs(info['acc_normal']).max()
    for g, w in zip(_host(em.post), ng_post):
        np.testing.assert_allclose(g, w, rtol=2e-4, atol=2e-4)


def test_model_api_list_of_inference_graphs():
    """HMM.expected_log_likelihood(batch, inference_graph=[...]) == the per-utterance calls of accumulate.py:47-57
    (golden 'phoneloop_mixtureset': alignment graph with repeated pdf ids, acoustic scale 0.7)."""
    import beer_b200 as beer
    from test_api_gpu import _joint_model, compiled, t32
    g = load_golden('phoneloop_mixtureset')
    emissions, (ns1, ns2, ms1, ms2) = _joint_model(beer, g)
    hmm = beer.HMM.create(compiled(beer, g), emissions)
    ag = compiled(beer, g, 'ali_')
    X3 = t32(g['X3'])
    batch = beer.Utterances.from_list([X3, X3[:31].clone(), X3], device=DEV)
    stats = hmm.sufficient_statistics(batch)
    exp_llh = hmm.expected_log_likelihood(stats, inference_graph=[ag, ag, ag], scale=0.7)
    got = exp_llh.double().cpu().numpy()
    n = len(g['X3'])
    np.testing.assert_allclose(got[:n], g['u3_exp_llh'], rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(got[n + 31:], g['u3_exp_llh'], rtol=1e-5, atol=1e-4)
    acc = hmm.accumulate(stats)
    hmm.clear_cache()
    # the same three utterances one at a time through the graph-plan kernels
    want = None
    for X in (X3, X3[:31], X3):
        s = hmm.sufficient_statistics(X)
        hmm.expected_log_likelihood(s, inference_graph=ag, scale=0.7)
        a = hmm.accumulate(s)
        want = a if want is None else {k: want[k] + a[k] for k in a}
        hmm.clear_cache()
    for param in (ns1.means_precisions, ns2.means_precisions, ms1.categoricalset.weights, ms2.categoricalset.weights):
        w = want[param].cpu().numpy()
        assert np.abs(acc[param].cpu().numpy() - w).max() <= 2e-5 * max(np.abs(w).max(), 1.0)


def skip_graph(rng, L, Kp):
    """An alignment graph that is NOT a chain: left-to-right with skip arcs (the topology of the silence unit in
    recipes/timit_v2/conf_61phns/hmm_gmm/hmm.yml:14-37)."""
    trans = np.full((L, L), -np.inf)
    for i in range(L):
        nxt = [j for j in (i, i + 1, i + 2) if j < L]
        w = rng.dirichlet(np.ones(len(nxt) + (1 if i == L - 1 else 0)))
        trans[i, nxt] = np.log(w[:len(nxt)])
    init = np.full(L, -np.inf)
    init[0] = 0.0
    final = np.full(L, -np.inf)
    final[-1] = np.log(0.3)
    return init, final, trans, rng.integers(0, Kp, L)


def test_engine_runs_any_alignment_graph_utterance_by_utterance():
    """Alignment graphs with skip arcs are no ChainBatch: the engine takes one graph plan per utterance (forward-backward
    launched per utterance, emission / statistics kernels batched); two VB iterations against the oracle of
    accumulate.py:47-57 with one inference graph per utterance."""
    from beer_b200 import ops, synthetic
    from beer_b200.engine import EmissionParams, Utterances, VBEngine
    rng = np.random.default_rng(11)
    Kp, D = 12, 8
    lens = [40, 75, 23, 5]
    graphs = [skip_graph(rng, L, Kp) for L in (7, 12, 5, 3)]
    with pytest.raises(ValueError):
        ops.ChainBatch(graphs, DEV)
    X = torch.as_tensor(2.0 * rng.standard_normal((sum(lens), D)), dtype=torch.float32, device=DEV)
    prior, post = synthetic.initial_normal_gamma(Kp, D, seed=2, device=DEV)
    em = EmissionParams(prior, post)
    plans = [ops.GraphPlan(*g, n_pdfs=Kp) for g in graphs]
    eng = VBEngine(em, plans, Utterances(X, lens), datasize=1000.0, scale=0.9, distributed=False, chunk_frames=120)
    host = lambda t: (t[0].double().cpu().numpy(), t[1].double().cpu().numpy()[:, None],
                      t[2].double().cpu().numpy()[:, None], t[3].double().cpu().numpy())
    ng_prior, ng_post = host(prior), host(post)
    off = np.concatenate([[0], np.cumsum(lens)])
    utts = [X[a:b].double().cpu().numpy() for a, b in zip(off[:-1], off[1:])]
    for it in range(2):
        want, ng_post, _, info = O.vb_iteration_hmm(utts, ng_prior, ng_post, None, None, None, datasize=1000.0, scale=0.9,
                                                    graphs=graphs)
        got = float(eng.step().item())
        assert abs(got - want) <= 1e-5 * abs(want), (it, got, want)
        assert np.abs(eng.acc.cpu().numpy() - info['acc_normal']).max() <= 3e-5 * np.abs(info['acc_normal']).max()
