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


@pytest.mark.parametrize('P,viterbi', [(7, False), (7, True), (40, True)])
def test_engine_learns_unit_weights_like_phoneloop_model(P, viterbi):
    """Batched engine with learned unit weights == the PhoneLoop model driven utterance by utterance through
    evidence_lower_bound (itself checked against the reference golden in test_api_gpu.py): same ELBOs, same
    unit-weight posteriors, same rewritten transitions after 3 iterations.  viterbi: the unit counts are the one-hot
    transition posteriors of the best path (hmm.py:49-54, phoneloop.py:83-101); 40 units = the several-units-per-lane
    Viterbi kernel with learned (non-uniform) unit weights from the second iteration on."""
    import beer_b200 as beer
    from beer_b200 import ops, synthetic
    dev = torch.device('cuda', 0)
    S, D = 4, 40
    K = P * S
    lens = [90, 41, 120, 64]
    N = sum(lens)

    def fresh_graph():
        g, starts, ends = synthetic.phone_loop_graph(P, S)
        return g, starts, ends

    means = 2.0 * torch.randn(K, D, generator=torch.Generator().manual_seed(0))
    g0, starts, ends = fresh_graph()
    full = synthetic.sample_utterances(g0, means, len(lens), max(lens), seed=1, device=dev).reshape(len(lens), max(lens), D)
    utts = [full[i, :n].contiguous() for i, n in enumerate(lens)]
    prior, post = synthetic.initial_normal_gamma(K, D, seed=2, device=dev)

    # (a) model API, one utterance at a time
    ga, _, _ = fresh_graph()
    ns = beer.NormalSet.create(torch.zeros(D, device=dev), torch.ones(D, device=dev), size=K, cov_type='diagonal')
    for dist, src in ((ns.means_precisions.prior, prior), (ns.means_precisions.posterior, post)):
        for name, t in zip(('mean', 'scale', 'shape', 'rates'), src):
            getattr(dist.params, name).copy_(t.reshape(getattr(dist.params, name).shape))
    pl = beer.PhoneLoop.create(ga, {i: s for i, s in enumerate(starts)}, {i: e for i, e in enumerate(ends)}, ns)
    optim = beer.VBConjugateOptimizer(pl.conjugate_bayesian_parameters(keepgroups=True), lrate=1.)
    want = []
    for _ in range(3):
        optim.init_step()
        elbo = beer.evidence_lower_bound(datasize=N)
        for X in utts:
            elbo += beer.evidence_lower_bound(pl, X, datasize=N, viterbi=viterbi)
        elbo.backward()
        want.append(float(elbo))
        optim.step()

    # (b) batched engine
    gb, _, _ = fresh_graph()
    conc = torch.full((P,), 1.0 / P, device=dev)
    units = beer.UnitWeights(conc.clone(), conc.clone(), gb, starts, ends)
    units.rewrite_graph()
    em = beer.EmissionParams(tuple(t.clone() for t in prior), tuple(t.clone() for t in post))
    eng = beer.VBEngine(em, gb.plan(n_pdfs=K), beer.Utterances(torch.cat(utts), lens), datasize=float(N),
                        distributed=False, unit_weights=units, viterbi=viterbi)
    got = [float(eng.step().item()) for _ in range(3)]
    np.testing.assert_allclose(got, want, rtol=2e-6)
    np.testing.assert_allclose(units.post.cpu().numpy(), pl.categorical.weights.posterior.params.concentrations.cpu().numpy(),
                               rtol=1e-5)
    ta, tb = ga.trans_log_probs.numpy(), gb.trans_log_probs.numpy()
    fin = np.isfinite(ta)
    np.testing.assert_allclose(tb[fin], ta[fin], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('C,chunk,scale,P,S,D', [(1, None, 1.0, 5, 3, 8), (2, 150, 0.8, 5, 3, 8), (4, 150, 0.8, 12, 4, 20),
                                                 (8, None, 1.0, 40, 4, 40), (8, 200, 1.0, 250, 4, 40)])
def test_viterbi_training_matches_oracle(C, chunk, scale, P, S, D):
    """Viterbi training in the batched engine (hmm.py:42-58 with viterbi=True): one-hot posteriors of the best path,
    three VB iterations against the oracle.  Mixtures of 4 / 8 Gaussians at D = 20 / 40 take the fp16 emission kernel
    (log2 llhs into the Viterbi recursion) and the sparse statistics along the path; 40 and 250 units (BASELINE
    configs[2]) run the several-units-per-lane Viterbi kernel."""
    from beer_b200 import ops, synthetic
    from beer_b200.engine import EmissionParams, Utterances, VBEngine, WeightGroup
    dev = torch.device('cuda', 0)
    K, M = P * S, P * S * C
    lens = [90, 33, 140, 61]
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
                   scale=scale, viterbi=True)
    assert bool(eng._path_mix) == (C >= 4)
    ng_prior, ng_post = _host(prior), _host(post)
    og = (graph.init_log_probs.double().numpy(), graph.final_log_probs.double().numpy(),
          graph.trans_log_probs.double().numpy(), graph.pdf_id_mapping)
    utts = [u.double().cpu().numpy() for u in utts_dev]
    for it in range(3):
        want, ng_post, dpost, info = O.vb_iteration_hmm(utts, ng_prior, ng_post, dprior, dpost, og, scale=scale,
                                                        viterbi=True)
        got = float(eng.step().item())
        assert abs(got - want) <= 1e-5 * abs(want), (it, got, want)
        acc = eng.acc.cpu().numpy()
        assert np.abs(acc - info['acc_normal']).max() <= 3e-5 * np.abs(info['acc_normal']).max()
    for g, w in zip(_host(em.post), ng_post):
        np.testing.assert_allclose(g, w, rtol=2e-4, atol=2e-4)


@pytest.mark.parametrize('chunk,use_graph,mix16', [(None, False, True), (333, False, True), (None, True, True),
                                                   (None, False, False)])
def test_vb_iterations_cfg2_shape(chunk, use_graph, mix16, monkeypatch):
    """The shape of BASELINE configs[1]: 25 units x 4 states = 100 single-Gaussian pdfs, D = 40 (mix16 = True: the
    fp16-split kernels with the statistics / posteriors as tensor-memory operands, csrc/mix16.cu with C = 1; False: the
    3xTF32 kernels of round 1); ragged utterances, three VB iterations against the fp64 oracle."""
    from beer_b200 import ops, synthetic
    from beer_b200.engine import EmissionParams, Utterances, VBEngine
    if not mix16:
        monkeypatch.setenv('BEER_B200_NO_MIX16', '1')
    dev = torch.device('cuda', 0)
    P, S, D = 25, 4, 40
    K = P * S
    lens = [150, 41, 297, 129, 64, 5]
    graph, _, _ = synthetic.phone_loop_graph(P, S)
    plan = ops.GraphPlan(graph.init_log_probs.numpy(), graph.final_log_probs.numpy(),
                         graph.trans_log_probs.numpy(), graph.pdf_id_mapping, n_pdfs=K)
    means = 2.0 * torch.randn(K, D, generator=torch.Generator().manual_seed(0))
    full = synthetic.sample_utterances(graph, means, len(lens), max(lens), seed=1, device=dev)
    full = full.reshape(len(lens), max(lens), D)
    utts_dev = [full[i, :n] for i, n in enumerate(lens)]
    X = torch.cat(utts_dev)
    prior, post = synthetic.initial_normal_gamma(K, D, seed=2, device=dev)
    em = EmissionParams(prior, post)
    N = sum(lens)
    eng = VBEngine(em, plan, Utterances(X, lens), datasize=float(N), chunk_frames=chunk, distributed=False,
                   use_graph=use_graph)
    assert (eng.mix16 is not None) == mix16
    ng_prior, ng_post = _host(prior), _host(post)
    og = (graph.init_log_probs.double().numpy(), graph.final_log_probs.double().numpy(),
          graph.trans_log_probs.double().numpy(), graph.pdf_id_mapping)
    utts = [u.double().cpu().numpy() for u in utts_dev]
    for it in range(3):
        want, ng_post, _, info = O.vb_iteration_hmm(utts, ng_prior, ng_post, None, None, og)
        got = float(eng.step().item())
        assert abs(got - want) <= 1e-5 * abs(want), (it, got, want)
        acc = eng.acc.cpu().numpy()
        assert np.abs(acc - info['acc_normal']).max() <= 3e-5 * np.abs(info['acc_normal']).max()
    for g, w in zip(_host(em.post), ng_post):
        np.testing.assert_allclose(g, w, rtol=2e-4, atol=2e-4)


@pytest.mark.parametrize('chunk,mix16,scale', [(None, False, 1.0), (170, False, 1.0), (None, True, 1.0), (170, True, 1.0),
                                               (None, True, 0.7), (170, False, 0.7)])
def test_vb_iterations_streamed_mixture_kernels(chunk, mix16, scale, monkeypatch):
    """(mix16 = False: the 3xTF32 kernels that materialise the per-Gaussian llhs; True: the fp16-split kernels that keep
    them on chip, csrc/mix16.cu.)  M = 160 Gaussians (20 pdfs x 8) at D = 20: the emission kernel streams its weight image in chunks and stores the
    per-Gaussian llhs with TMA tensor stores, the statistics kernel runs its bulk-staged mixture variant; three VB
    iterations (ragged utterances, chunk boundaries inside frame tiles) against the oracle."""
    from beer_b200 import ops, synthetic
    from beer_b200.engine import EmissionParams, Utterances, VBEngine, WeightGroup
    dev = torch.device('cuda', 0)
    if not mix16:
        monkeypatch.setenv('BEER_B200_NO_MIX16', '1')
    P, S, D, C = 5, 4, 20, 8
    K, M = P * S, P * S * C
    lens = [150, 41, 97, 129, 64]
    graph, _, _ = synthetic.phone_loop_graph(P, S)
    plan = ops.GraphPlan(graph.init_log_probs.numpy(), graph.final_log_probs.numpy(),
                         graph.trans_log_probs.numpy(), graph.pdf_id_mapping, n_pdfs=K)
    means = 2.0 * torch.randn(K, D, generator=torch.Generator().manual_seed(0))
    full = synthetic.sample_utterances(graph, means, len(lens), max(lens), seed=1, device=dev)
    full = full.reshape(len(lens), max(lens), D)
    utts_dev = [full[i, :n] for i, n in enumerate(lens)]
    X = torch.cat(utts_dev)
    prior, post = synthetic.initial_normal_gamma(M, D, seed=2, device=dev)
    conc = torch.full((K, C), 1.0 / C, device=dev)
    groups = (WeightGroup(0, K, C, conc.clone(), conc.clone()),)
    em = EmissionParams(prior, post, comp_off=np.arange(K + 1) * C, weight_groups=groups)
    assert em.use_tc and ops.accumulate_tc_supported(M, D) and em.use16 == mix16
    dprior, dpost = conc.double().cpu().numpy(), conc.double().cpu().numpy()
    N = sum(lens)
    eng = VBEngine(em, plan, Utterances(X, lens), datasize=float(N), chunk_frames=chunk, distributed=False, scale=scale)
    assert (eng.mix16 is not None) == mix16
    ng_prior, ng_post = _host(prior), _host(post)
    og = (graph.init_log_probs.double().numpy(), graph.final_log_probs.double().numpy(),
          graph.trans_log_probs.double().numpy(), graph.pdf_id_mapping)
    utts = [u.double().cpu().numpy() for u in utts_dev]
    for it in range(3):
        want, ng_post, dpost, info = O.vb_iteration_hmm(utts, ng_prior, ng_post, dprior, dpost, og, scale=scale)
        got = float(eng.step().item())
        assert abs(got - want) <= 1e-5 * abs(want), (it, got, want)
        acc = eng.acc.cpu().numpy()
        assert np.abs(acc - info['acc_normal']).max() <= 3e-5 * np.abs(info['acc_normal']).max()
    for g, w in zip(_host(em.post), ng_post):
        np.testing.assert_allclose(g, w, rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(groups[0].post.double().cpu().numpy(), dpost, rtol=2e-4, atol=1e-5)


@pytest.mark.parametrize('use_graph,mix16', [(False, True), (True, True), (False, False)])
def test_vb_iterations_cfg3_shape(use_graph, mix16, monkeypatch):
    """The exact shape of BASELINE configs[2] (the north-star target): 250 units x 4 states = 1000 states, 8
    Gaussians per state (M = 8000: 63 Gaussian tiles of the statistics kernel, 125 weight chunks of the emission
    kernel), D = 40, the eight-warps-per-utterance left-to-right scan; three short ragged utterances, two VB
    iterations against the fp64 oracle of accumulate + update (accumulate.py:37-63, update.py:39-62)."""
    from beer_b200 import ops, synthetic
    from beer_b200.engine import EmissionParams, Utterances, VBEngine, WeightGroup
    dev = torch.device('cuda', 0)
    if not mix16:
        monkeypatch.setenv('BEER_B200_NO_MIX16', '1')
    P, S, D, C = 250, 4, 40, 8
    K, M = P * S, P * S * C
    lens = [70, 33, 129]
    graph, _, _ = synthetic.phone_loop_graph(P, S)
    plan = ops.GraphPlan(graph.init_log_probs.numpy(), graph.final_log_probs.numpy(),
                         graph.trans_log_probs.numpy(), graph.pdf_id_mapping, n_pdfs=K)
    means = 2.0 * torch.randn(K, D, generator=torch.Generator().manual_seed(0))
    full = synthetic.sample_utterances(graph, means, len(lens), max(lens), seed=1, device=dev)
    full = full.reshape(len(lens), max(lens), D)
    utts_dev = [full[i, :n] for i, n in enumerate(lens)]
    X = torch.cat(utts_dev)
    prior, post = synthetic.initial_normal_gamma(M, D, seed=2, device=dev)
    conc = torch.full((K, C), 1.0 / C, device=dev)
    groups = (WeightGroup(0, K, C, conc.clone(), conc.clone()),)
    em = EmissionParams(prior, post, comp_off=np.arange(K + 1) * C, weight_groups=groups)
    assert em.use_tc and ops.accumulate_tc_supported(M, D)
    dprior, dpost = conc.double().cpu().numpy(), conc.double().cpu().numpy()
    N = sum(lens)
    eng = VBEngine(em, plan, Utterances(X, lens), datasize=float(N), distributed=False, use_graph=use_graph)
    ng_prior, ng_post = _host(prior), _host(post)
    og = (graph.init_log_probs.double().numpy(), graph.final_log_probs.double().numpy(),
          graph.trans_log_probs.double().numpy(), graph.pdf_id_mapping)
    utts = [u.double().cpu().numpy() for u in utts_dev]
    for it in range(3 if use_graph else 2):
        want, ng_post, dpost, info = O.vb_iteration_hmm(utts, ng_prior, ng_post, dprior, dpost, og)
        got = float(eng.step().item())
        assert abs(got - want) <= 1e-5 * abs(want), (it, got, want)
        acc = eng.acc.cpu().numpy()
        assert np.abs(acc - info['acc_normal']).max() <= 3e-5 * np.abs(info['acc_normal']).max()
    # 232 frames over 8000 Gaussians: most Gaussians own a fraction of a frame, so a statistic that is right to 3e-5
    # of the LARGEST entry moves the mean of such a Gaussian by up to ~1e-3 (mean = sum w x / (kappa0 + sum w))
    # (the rates difference the two second moments: b = b0 + (sum w x^2 + kappa0 m0^2 - kappa m^2) / 2)
    for g, w in zip(_host(em.post), ng_post):
        np.testing.assert_allclose(g, w, rtol=2e-4, atol=3e-3)
    np.testing.assert_allclose(groups[0].post.double().cpu().numpy(), dpost, rtol=2e-4, atol=3e-4)


@pytest.mark.parametrize('chunk,use_graph', [(None, False), (500, False), (None, True)])
def test_vb_iterations_gmm_only(chunk, use_graph):
    """BASELINE configs[4]: the batched engine without an HMM (plan=None) on a 512-component diagonal GMM, D = 40 --
    emission kernel over pseudo-pdfs of 8 components, softmax over the frame, statistics with the responsibilities
    recomputed on chip -- three VB iterations over ragged utterances against the oracle of Mixture (mixture.py:70-102)
    summed as `beer hmm accumulate` + `update` do."""
    from beer_b200 import synthetic
    from beer_b200.engine import EmissionParams, Utterances, VBEngine, WeightGroup
    dev = torch.device('cuda', 0)
    C, D = 512, 40
    lens = [300, 41, 777, 129, 64]
    N = sum(lens)
    X = synthetic.sample_gmm_frames(N, D, seed=3, device=dev)
    prior, post = synthetic.initial_normal_gamma(C, D, seed=2, device=dev)
    conc = torch.full((1, C), 1.0 / C, device=dev)
    groups = (WeightGroup(0, 1, C, conc.clone(), conc.clone()),)
    em = EmissionParams(prior, post, comp_off=np.array([0, C]), weight_groups=groups)
    assert em.gmm_C == 8
    eng = VBEngine(em, None, Utterances(X, lens), datasize=float(N), chunk_frames=chunk, distributed=False,
                   use_graph=use_graph)
    ng_prior, ng_post = _host(prior), _host(post)
    dprior = dpost = conc.double().cpu().numpy().reshape(-1)
    off = np.concatenate([[0], np.cumsum(lens)])
    Xh = X.double().cpu().numpy()
    for it in range(3):
        kl = O.normalgamma_kl(ng_post, ng_prior).sum() + O.dirichlet_kl(dpost, dprior).sum()
        want, acc_n, acc_d = 0.0, 0.0, 0.0
        for u in range(len(lens)):
            r = O.gmm_estep(Xh[off[u]:off[u + 1]], ng_post, dpost)
            want += O.elbo_value(r['exp_llh'], kl, N)
            acc_n = acc_n + r['acc_normal']
            acc_d = acc_d + r['acc_dirichlet']
        got = float(eng.step().item())
        assert abs(got - want) <= 1e-5 * abs(want), (it, got, want)
        acc = eng.acc.cpu().numpy()
        assert np.abs(acc - acc_n).max() <= 3e-5 * np.abs(acc_n).max()
        ng_post = O.natural_grad_update_normalgamma(ng_prior, ng_post, acc_n, 1.)
        dpost = O.natural_grad_update_dirichlet(dprior, dpost, acc_d, 1.)
    for g, w in zip(_host(em.post), ng_post):
        np.testing.assert_allclose(g, w, rtol=3e-4, atol=3e-4)
    np.testing.assert_allclose(groups[0].post.double().cpu().numpy().reshape(-1), dpost, rtol=3e-4, atol=1e-5)


def test_sparse_and_dense_statistics_give_the_same_iterations():
    """VBEngine(sparse_stats=...) at the cfg3 shape: the statistics kernel over the marked (frame tile, Gaussian tile)
    pairs only against the same kernel over every pair -- the skipped pairs are exact zeros, the rest is summed in another
    grouping: statistics and models agree to fp32 summation order; the fraction of marked pairs falls as the model fits."""
    from beer_b200 import ops, synthetic
    from beer_b200.engine import EmissionParams, Utterances, VBEngine, WeightGroup
    dev = torch.device('cuda', 0)
    P, S, D, C = 250, 4, 40, 8
    K, M = P * S, P * S * C
    lens = [300, 129, 210, 64]
    graph, _, _ = synthetic.phone_loop_graph(P, S)
    plan = ops.GraphPlan(graph.init_log_probs.numpy(), graph.final_log_probs.numpy(),
                         graph.trans_log_probs.numpy(), graph.pdf_id_mapping, n_pdfs=K)
    means = 2.0 * torch.randn(K, D, generator=torch.Generator().manual_seed(0))
    full = synthetic.sample_utterances(graph, means, len(lens), max(lens), seed=1, device=dev).reshape(len(lens), max(lens), D)
    X = torch.cat([full[i, :n] for i, n in enumerate(lens)])
    engines = []
    for sparse in (True, False):
        prior, post = synthetic.initial_normal_gamma(M, D, seed=2, device=dev)
        conc = torch.full((K, C), 1.0 / C, device=dev)
        em = EmissionParams(prior, post, comp_off=np.arange(K + 1) * C,
                            weight_groups=(WeightGroup(0, K, C, conc.clone(), conc.clone()),))
        engines.append(VBEngine(em, plan, Utterances(X, lens), datasize=float(sum(lens)), distributed=False,
                                sparse_stats=sparse, chunk_frames=450))
    fractions = []
    for it in range(4):
        engines[0].profile = {}
        a, b = (float(e.step().item()) for e in engines)
        fractions.append(float(engines[0].active_fraction))
        assert abs(a - b) <= 1e-9 * abs(b), (it, a, b)
        sa, da = engines[0].acc, engines[1].acc
        assert (sa - da).abs().max().item() <= 2e-6 * da.abs().max().item(), it
    for p, q in zip(engines[0].em.post, engines[1].em.post):
        assert (p - q).abs().max().item() <= 1e-5 * q.abs().max().item()
    assert fractions[-1] < 0.5 * fractions[0], fractions


@pytest.mark.parametrize('viterbi', [False, True])
def test_bigram_loop_on_the_fp16_mixture_kernels(viterbi, monkeypatch):
    """A bigram phone loop over mixtures at a shape the fp16-split kernels take (40-d, 8 Gaussians per pdf): the emission
    kernel then hands log2 llhs to the generic scan and to the transition-posterior kernel (scale x ln 2).  Against the
    same engine on the fp32 SIMT kernels (BEER_B200_NO_MIX16): ELBOs, bigram posteriors and rewritten arcs."""
    import beer_b200 as beer
    from beer_b200 import synthetic
    dev = torch.device('cuda', 0)
    P, S, D, C = 8, 4, 40, 8
    K, M = P * S, P * S * C
    lens = [150, 64, 97]
    g0, starts, ends = synthetic.phone_loop_graph(P, S)
    means = 2.0 * torch.randn(K, D, generator=torch.Generator().manual_seed(0))
    full = synthetic.sample_utterances(g0, means, len(lens), max(lens), seed=1, device=dev).reshape(len(lens), max(lens), D)
    X = torch.cat([full[i, :n] for i, n in enumerate(lens)])
    out = []
    for mix16 in (True, False):
        if mix16:
            monkeypatch.delenv('BEER_B200_NO_MIX16', raising=False)
        else:
            monkeypatch.setenv('BEER_B200_NO_MIX16', '1')
        graph, _, _ = synthetic.phone_loop_graph(P, S)
        cs = beer.CategoricalSet.create(torch.ones(P, P, device=dev) / P, 1.)
        units = beer.BigramUnitWeights(cs, graph, starts, ends)
        units.rewrite_graph()
        prior, post = synthetic.initial_normal_gamma(M, D, seed=2, device=dev)
        conc = torch.full((K, C), 1.0 / C, device=dev)
        em = beer.EmissionParams(prior, post, comp_off=np.arange(K + 1) * C,
                                 weight_groups=(beer.WeightGroup(0, K, C, conc.clone(), conc.clone()),))
        eng = beer.VBEngine(em, graph.plan(n_pdfs=K), beer.Utterances(X, lens), datasize=float(sum(lens)),
                            distributed=False, unit_weights=units, viterbi=viterbi)
        assert (eng.mix16 is not None) == mix16
        elbos = [float(eng.step().item()) for _ in range(3)]
        out.append((elbos, cs.weights.posterior.params.concentrations.double().cpu().numpy(),
                    graph.trans_log_probs.numpy().copy()))
    (ea, ca, ta), (eb, cb, tb) = out
    np.testing.assert_allclose(ea, eb, rtol=2e-6)
    np.testing.assert_allclose(ca, cb, rtol=1e-4, atol=1e-4)
    fin = np.isfinite(tb)
    assert np.array_equal(np.isfinite(ta), fin)
    np.testing.assert_allclose(ta[fin], tb[fin], rtol=1e-4, atol=1e-4)
    assert np.abs(cb - 1.0 / P).max() > 0.5          # the bigram counts did arrive
