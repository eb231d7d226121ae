"""GPU parity tests through the public, reference-shaped API (`import beer_b200 as beer`) against the
fp64 goldens dumped from the live reference by tests/golden/make_goldens.py.  Each test mirrors the
reference call sequence that produced its golden (same class names, same keyword arguments).

Tolerances: ELBO / expected log-likelihood sums 1e-5 relative (north_star), posteriors 1e-5
absolute, accumulated statistics 3e-5 of the largest entry, updated standard parameters 2e-4
(they are fp32 tensors compared with an fp64 run), alignments identical.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.fixture(scope='module')
def beer():
    import beer_b200
    beer_b200.ops.require_cuda() if hasattr(beer_b200, 'ops') else None
    return beer_b200


def t32(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device=DEV)


def set_ng(dist, g, prefix):
    p = dist.params
    for name in ('mean', 'scale', 'shape', 'rates'):
        getattr(p, name).copy_(t32(g[prefix + name]).reshape(getattr(p, name).shape))


def get_ng(dist):
    p = dist.params
    return [getattr(p, n).double().cpu().numpy() for n in ('mean', 'scale', 'shape', 'rates')]


def compiled(beer, g, prefix='g_'):
    return beer.CompiledGraph(torch.from_numpy(g[prefix + 'init']).float(), torch.from_numpy(g[prefix + 'final']).float(),
                              torch.from_numpy(g[prefix + 'trans']).float(), [int(i) for i in g[prefix + 'map']])


def normalset(beer, g, size, D, prior='prior_', post='post0_'):
    ns = beer.NormalSet.create(torch.zeros(D, device=DEV), torch.ones(D, device=DEV), size=size, prior_strength=1.,
                               noise_std=1., cov_type='diagonal')
    set_ng(ns.means_precisions.prior, g, prior)
    set_ng(ns.means_precisions.posterior, g, post)
    return ns


def vb_loop(beer, model, X, n_iter, **kw):
    optim = beer.VBConjugateOptimizer(model.mean_field_factorization(), lrate=1.)
    elbos = []
    for _ in range(n_iter):
        optim.init_step()
        elbo = beer.evidence_lower_bound(model, X, datasize=len(X), **kw)
        elbo.backward()
        elbos.append(float(elbo))
        optim.step()
    return np.asarray(elbos)


# ---------------------------------------------------------------------------------------------
def test_dists_against_reference(beer):
    g = load_golden('dists')
    q = beer.dists.NormalGamma.from_std_parameters(t32(g['ng_mean']), t32(g['ng_scale']), t32(g['ng_shape']),
                                                   t32(g['ng_rates']))
    p = beer.dists.NormalGamma.from_std_parameters(t32(g['ngp_mean']), t32(g['ngp_scale']), t32(g['ngp_shape']),
                                                   t32(g['ngp_rates']))
    np.testing.assert_allclose(q.natural_parameters().cpu().numpy(), g['ng_nat'], rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(q.expected_sufficient_statistics().cpu().numpy(), g['ng_ets'], rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(q.log_norm().cpu().numpy(), g['ng_lognorm'], rtol=1e-6)
    np.testing.assert_allclose(beer.dists.kl_div(q, p).item(), g['ng_kl'].sum(), rtol=1e-6)
    back = beer.dists.NormalGammaStdParams.from_natural_parameters(t32(g['ng_nat']))
    for name in ('mean', 'scale', 'shape', 'rates'):
        np.testing.assert_allclose(getattr(back, name).cpu().numpy(), g['ng_back_' + name], rtol=2e-5, atol=2e-6)
    lfn = q.conjugate()
    stats = lfn.sufficient_statistics(t32(g['X']))
    np.testing.assert_allclose(stats.cpu().numpy(), g['stats'], rtol=1e-6)
    llh = lfn(q.expected_sufficient_statistics(), stats)
    np.testing.assert_allclose(llh.cpu().numpy(), g['llh'], rtol=1e-5, atol=1e-4)
    assert len(q) == 6 and q.dim == (6, 5, 5)

    dq = beer.dists.Dirichlet.from_std_parameters(t32(g['dir_conc']))
    dp = beer.dists.Dirichlet.from_std_parameters(t32(g['dirp_conc']))
    np.testing.assert_allclose(dq.natural_parameters().cpu().numpy(), g['dir_nat'], rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(dq.expected_sufficient_statistics().cpu().numpy(), g['dir_ets'], rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(dq.log_norm().cpu().numpy(), g['dir_lognorm'], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(beer.dists.kl_div(dq, dp).item(), g['dir_kl'].sum(), rtol=1e-6)
    dback = beer.dists.DirichletStdParams.from_natural_parameters(t32(g['dir_nat']))
    np.testing.assert_allclose(dback.concentrations.cpu().numpy(), g['dir_back'], rtol=2e-6, atol=2e-6)
    cl = dq.conjugate()
    np.testing.assert_allclose(cl.sufficient_statistics(t32(g['cat_data'])).cpu().numpy(), g['cat_stats'], rtol=1e-6)
    np.testing.assert_allclose(dq.expected_log_weights().cpu().numpy(), g['dir_logw'], rtol=2e-6, atol=1e-6)
    with pytest.raises(beer.dists.DistributionTypeMismatch):
        beer.dists.kl_div(q, dq)


def test_mixture_cfg1(beer):
    """BASELINE configs[0]: 8-component diagonal Mixture on 2-D points, the loop of examples/HMM.ipynb
    cell 9 with a Mixture."""
    g = load_golden('gmm_cfg1')
    X = t32(g['X'])
    ns = normalset(beer, g, 8, 2)
    gmm = beer.Mixture.create(ns)
    w = gmm.categorical.weights
    w.prior.params.concentrations.copy_(t32(g['dprior']))
    w.posterior.params.concentrations.copy_(t32(g['dpost0']))
    par = gmm.modelset.means_precisions

    stats = gmm.sufficient_statistics(X)
    exp_llh = gmm.expected_log_likelihood(stats)
    np.testing.assert_allclose(exp_llh.double().cpu().numpy(), g['exp_llh'], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(gmm.kl_div_posterior_prior().item(), g['kl'], rtol=1e-6)
    acc = gmm.accumulate(stats)
    assert np.abs(acc[par].cpu().numpy() - g['acc_normal']).max() <= 3e-5 * np.abs(g['acc_normal']).max()
    np.testing.assert_allclose(acc[w].cpu().numpy(), g['acc_dirichlet'], rtol=2e-5)
    gmm.clear_cache()
    assert np.abs(gmm.posteriors(X).double().cpu().numpy() - g['resps']).max() <= 1e-5
    # labelled path (mixture.py:84-87)
    got = gmm.expected_log_likelihood(stats, labels=torch.from_numpy(g['labels']))
    np.testing.assert_allclose(got.double().cpu().numpy(), g['exp_llh_labels'], rtol=1e-5, atol=1e-4)
    gmm.clear_cache()
    elbos = vb_loop(beer, gmm, X, 6)
    np.testing.assert_allclose(elbos, g['elbos'], rtol=1e-5)
    for got, name in zip(get_ng(par.posterior), ('mean', 'scale', 'shape', 'rates')):
        np.testing.assert_allclose(got.reshape(g['post6_' + name].shape), g['post6_' + name], rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(w.posterior.params.concentrations.cpu().numpy(), g['dpost6'], rtol=2e-4)


@pytest.mark.parametrize('name,n_iter', [('hmm_small', 3), ('hmm_scaled', 3), ('hmm_cfg2_T200', 2)])
def test_hmm_against_reference(beer, name, n_iter):
    g = load_golden(name)
    scale = float(g['scale'])
    X = t32(g['X'])
    T, D = g['X'].shape
    graph = compiled(beer, g)
    K = graph.n_states
    ns = normalset(beer, g, K, D)
    hmm = beer.HMM.create(graph, ns)
    par = ns.means_precisions

    stats = hmm.sufficient_statistics(X)
    exp_llh = hmm.expected_log_likelihood(stats, inference_graph=hmm.graph, scale=scale)
    np.testing.assert_allclose(exp_llh.double().sum().item(), g['exp_llh'].sum(), rtol=1e-5)
    np.testing.assert_allclose(exp_llh.double().cpu().numpy(), g['exp_llh'], rtol=1e-4, atol=2e-3)
    np.testing.assert_allclose(hmm.kl_div_posterior_prior().item(), g['kl'], rtol=1e-6)
    acc = hmm.accumulate(stats)
    assert np.abs(acc[par].cpu().numpy() - g['acc_normal']).max() <= 3e-5 * np.abs(g['acc_normal']).max()
    hmm.clear_cache()
    elbo = beer.evidence_lower_bound(hmm, X, datasize=3 * T, inference_graph=hmm.graph, scale=scale)
    np.testing.assert_allclose(float(elbo), g['elbo_datasize3T'], rtol=1e-5)

    # Viterbi: alignment, decode, one-hot E-step
    pc = scale * ns.expected_log_likelihood(stats)[:, torch.as_tensor(graph.pdf_id_mapping, device=DEV)]
    np.testing.assert_array_equal(hmm.graph.best_path(pc).numpy(), g['viterbi_path'])
    np.testing.assert_array_equal(hmm.decode(X, scale=scale).numpy(), g['decode'])
    assert np.abs(hmm.posteriors(X).double().cpu().numpy() - g['posteriors']).max() <= 1e-5
    exp_llh_v = hmm.expected_log_likelihood(stats, inference_graph=hmm.graph, viterbi=True, scale=scale)
    np.testing.assert_allclose(exp_llh_v.double().cpu().numpy(), g['exp_llh_viterbi'], rtol=1e-5, atol=2e-4)
    acc_v = hmm.accumulate(stats)[par].cpu().numpy()
    assert np.abs(acc_v - g['acc_normal_viterbi']).max() <= 3e-5 * np.abs(g['acc_normal_viterbi']).max()
    hmm.clear_cache()
    # a given state path is the same code path as Viterbi (hmm.py:45-48)
    exp_llh_p = hmm.expected_log_likelihood(stats, inference_graph=hmm.graph, state_path=g['viterbi_path'],
                                            scale=scale)
    np.testing.assert_allclose(exp_llh_p.cpu().numpy(), exp_llh_v.cpu().numpy(), rtol=0, atol=0)
    hmm.clear_cache()

    elbos = vb_loop(beer, hmm, X, n_iter, inference_graph=hmm.graph, scale=scale)
    np.testing.assert_allclose(elbos, g['elbos'], rtol=1e-5)
    for got, pname in zip(get_ng(par.posterior), ('mean', 'scale', 'shape', 'rates')):
        want = g[f'post{n_iter}_' + pname]
        np.testing.assert_allclose(got.reshape(want.shape), want, rtol=2e-4, atol=2e-4)


def _joint_model(beer, g):
    C1, C2, K1 = int(g['C1']), int(g['C2']), int(g['K1'])
    K = len(g['g_map'])
    D = g['X1'].shape[1]
    ns1 = normalset(beer, g, K1 * C1, D, 'g1_prior_', 'g1_post0_')
    ns2 = normalset(beer, g, (K - K1) * C2, D, 'g2_prior_', 'g2_post0_')
    ms1 = beer.MixtureSet.create(K1, ns1, prior_strength=1.)
    ms2 = beer.MixtureSet.create(K - K1, ns2, prior_strength=1.)
    for ms, tag in ((ms1, 'g1'), (ms2, 'g2')):
        ms.categoricalset.weights.prior.params.concentrations.copy_(t32(g[tag + '_dprior']))
        ms.categoricalset.weights.posterior.params.concentrations.copy_(t32(g[tag + '_dpost0']))
    return beer.JointModelSet([ms1, ms2]), (ns1, ns2, ms1, ms2)


def test_joint_mixtureset_hmm_with_alignment_graph(beer):
    """The CLI emission stack JointModelSet([MixtureSet(NormalSet), MixtureSet(NormalSet)]) with two
    different numbers of components, driven through an alignment graph with repeated pdf ids and an
    acoustic scale (accumulate.py:47-57)."""
    g = load_golden('phoneloop_mixtureset')
    emissions, (ns1, ns2, ms1, ms2) = _joint_model(beer, g)
    hmm = beer.HMM.create(compiled(beer, g), emissions)
    ag = compiled(beer, g, 'ali_')
    X3 = t32(g['X3'])
    stats = hmm.sufficient_statistics(X3)
    exp_llh = hmm.expected_log_likelihood(stats, inference_graph=ag, scale=0.7)
    np.testing.assert_allclose(exp_llh.double().cpu().numpy(), g['u3_exp_llh'], rtol=1e-5, atol=1e-4)
    acc = hmm.accumulate(stats)
    for param, key in ((ns1.means_precisions, 'u3_acc_g1'), (ns2.means_precisions, 'u3_acc_g2'),
                       (ms1.categoricalset.weights, 'u3_acc_d1'), (ms2.categoricalset.weights, 'u3_acc_d2')):
        got = acc[param].cpu().numpy()
        assert np.abs(got - g[key]).max() <= 3e-5 * max(np.abs(g[key]).max(), 1.0), key
    hmm.clear_cache()
    # decoding graph (state posteriors of utterance 1)
    X1 = t32(g['X1'])
    assert np.abs(hmm.posteriors(X1).double().cpu().numpy() - g['u1_gamma']).max() <= 1e-5
    # standalone JointModelSet / MixtureSet calls (mixtureset.py:85-112)
    llh = emissions.expected_log_likelihood(emissions.sufficient_statistics(X1))
    np.testing.assert_allclose(llh.double().cpu().numpy(), g['u1_pdf_llh'], atol=1e-4)
    emissions.clear_cache()


def test_utterance_batch_equals_sum_of_utterances(beer):
    """`elbo += evidence_lower_bound(model, X_u, datasize=N)` over utterances (accumulate.py:39-59) ==
    one call on the ragged batch."""
    g = load_golden('hmm_cfg2_T200')
    D = g['X'].shape[1]
    graph = compiled(beer, g)
    hmm = beer.HMM.create(graph, normalset(beer, g, graph.n_states, D))
    lens = [200, 64, 37, 120]
    rng = np.random.default_rng(0)
    utts = [t32(g['X'][s:s + n]) for s, n in ((int(rng.integers(0, 200 - n + 1)), n) for n in lens)]
    N = sum(lens)
    total = beer.evidence_lower_bound(datasize=N)
    for X in utts:
        total += beer.evidence_lower_bound(hmm, X, datasize=N, inference_graph=graph)
    batch = beer.evidence_lower_bound(hmm, beer.Utterances.from_list(utts, device=DEV), datasize=N,
                                      inference_graph=graph)
    np.testing.assert_allclose(float(batch), float(total), rtol=1e-6)
    par = hmm.modelset.original_modelset.means_precisions
    np.testing.assert_allclose(batch._acc_stats[par].cpu().numpy(), total._acc_stats[par].cpu().numpy(),
                               rtol=1e-5, atol=1e-5)
    assert batch._minibatchsize == total._minibatchsize == N


def test_error_behaviour(beer):
    """Argument checks raise what the reference raises (objectives.py:79-83, 173-175)."""
    with pytest.raises(ValueError):
        beer.evidence_lower_bound()
    with pytest.raises(ValueError):
        beer.evidence_lower_bound(datasize=10) + beer.evidence_lower_bound(datasize=11)
    with pytest.raises(ValueError):
        beer.evidence_lower_bound(datasize=10) + 3
    with pytest.raises(NotImplementedError):
        beer.NormalSet.create(torch.zeros(2, device=DEV), torch.ones(2, device=DEV), 3, cov_type='full')
    with pytest.raises(beer._lib.BeerB200Error):      # CPU tensors: no fallback, fail loudly
        beer.NormalSet.create(torch.zeros(2), torch.ones(2), 3, cov_type='diagonal')
    ns = beer.NormalSet.create(torch.zeros(2, device=DEV), torch.ones(2, device=DEV), 3, cov_type='diagonal')
    with pytest.raises(beer._lib.BeerB200Error):
        ns.expected_log_likelihood(ns.sufficient_statistics(torch.zeros(4, 2)))


def test_phoneloop_unit_counts_and_training_loop(beer):
    """PhoneLoop over the CLI emission stack (golden 'phoneloop_mixtureset'): the transition-posterior
    path (inference_graph=None, hmm.py:76) with the unit counts reduced in the scan kernel, the in-place
    rewrite of the graph after every update (phoneloop.py:53-65) and the accumulate/update loop of
    `beer hmm accumulate` + `beer hmm update` over two utterances (accumulate.py:37-59, update.py:37-62)."""
    g = load_golden('phoneloop_mixtureset')
    emissions, (ns1, ns2, ms1, ms2) = _joint_model(beer, g)
    start_pdf = {f'u{i}': int(s) for i, s in enumerate(g['start_idxs'])}
    end_pdf = {f'u{i}': int(s) for i, s in enumerate(g['end_idxs'])}
    pl = beer.PhoneLoop.create(compiled(beer, g), start_pdf, end_pdf, emissions, prior_strength=1.)
    wu = pl.categorical.weights
    np.testing.assert_allclose(wu.posterior.params.concentrations.cpu().numpy(), g['u_dpost0'], rtol=1e-6)
    np.testing.assert_allclose(pl.graph.trans_log_probs.numpy(), g['g_trans'], rtol=1e-5, atol=1e-6)

    X1, X2 = t32(g['X1']), t32(g['X2'])
    stats = pl.sufficient_statistics(X1)
    exp_llh = pl.expected_log_likelihood(stats)
    np.testing.assert_allclose(exp_llh.double().cpu().numpy(), g['u1_exp_llh'], rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(pl.kl_div_posterior_prior().item(), g['kl'], rtol=1e-6)
    acc = pl.accumulate(stats)
    np.testing.assert_allclose(acc[wu].cpu().numpy(), g['u1_acc_units'], rtol=2e-5, atol=2e-5)
    for param, key in ((ns1.means_precisions, 'u1_acc_g1'), (ns2.means_precisions, 'u1_acc_g2'),
                       (ms1.categoricalset.weights, 'u1_acc_d1'), (ms2.categoricalset.weights, 'u1_acc_d2')):
        assert np.abs(acc[param].cpu().numpy() - g[key]).max() <= 3e-5 * max(np.abs(g[key]).max(), 1.0), key
    pl.clear_cache()
    # aligned training: no transition posteriors, zero unit statistics (phoneloop.py:98-100)
    ag = compiled(beer, g, 'ali_')
    s3 = pl.sufficient_statistics(t32(g['X3']))
    pl.expected_log_likelihood(s3, inference_graph=ag, scale=0.7)
    np.testing.assert_allclose(pl.accumulate(s3)[wu].cpu().numpy(), g['u3_acc_units'], atol=1e-12)
    pl.clear_cache()

    optim = beer.VBConjugateOptimizer(pl.conjugate_bayesian_parameters(keepgroups=True), lrate=1.)
    N = len(X1) + len(X2)
    elbos = []
    for it in range(3):
        optim.init_step()
        elbo = beer.evidence_lower_bound(datasize=N)
        for X in (X1, X2):
            elbo += beer.evidence_lower_bound(pl, X, datasize=N)
        elbo.backward()
        elbos.append(float(elbo))
        optim.step()
        want = g[f'it{it + 1}_trans']
        got = pl.graph.trans_log_probs.numpy()
        fin = np.isfinite(want)
        assert np.array_equal(np.isfinite(got), fin)
        np.testing.assert_allclose(got[fin], want[fin], rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(elbos, g['elbos'], rtol=1e-5)
    np.testing.assert_allclose(wu.posterior.params.concentrations.cpu().numpy(), g['u_dpost3'], rtol=2e-4)
    for ns, tag in ((ns1, 'g1'), (ns2, 'g2')):
        for got, pname in zip(get_ng(ns.means_precisions.posterior), ('mean', 'scale', 'shape', 'rates')):
            want = g[f'{tag}_post3_' + pname]
            np.testing.assert_allclose(got.reshape(want.shape), want, rtol=3e-4, atol=3e-4)


def test_compiled_graph_transition_posteriors(beer):
    """CompiledGraph.posteriors(llhs, trans_posteriors=True) (graph.py:289-326): dense ergodic graph with an
    unreachable state (the reference's NaN -> 0), and the full (T-1, K, K) tensor of a phone loop."""
    g = load_golden('dense_ergodic')
    cg = compiled(beer, g)
    (gamma, xi), lognorm = cg.posteriors(t32(g['llhs']), trans_posteriors=True)
    assert xi.shape == (len(g['llhs']) - 1, cg.n_states, cg.n_states)
    assert np.abs(gamma.double().cpu().numpy() - g['gamma']).max() <= 1e-5
    assert np.abs(xi.double().sum(dim=0).cpu().numpy() - g['xi_sum']).max() <= 2e-5 * len(g['llhs'])
    np.testing.assert_allclose(xi.double().sum(dim=(1, 2)).cpu().numpy(), 1.0, atol=1e-5)
    cg2 = compiled(beer, g, 'g2_')
    (_, xi2), _ = cg2.posteriors(t32(g['llhs2']), trans_posteriors=True)
    assert torch.isfinite(xi2).all() and (xi2[:, :, 6] == 0).all()          # unreachable destination
    b = load_golden('bigram_phoneloop')
    for tag in ('bg', 'un'):
        cgb = compiled(beer, b, tag + '_g_')
        ns = normalset(beer, b, cgb.n_states, b[tag + '_X1'].shape[1], prior=tag + '_prior_', post=tag + '_post0_')
        llhs = ns.expected_log_likelihood(ns.sufficient_statistics(t32(b[tag + '_X1'])))
        (gam, x), _ = cgb.posteriors(llhs[:, [int(i) for i in b[tag + '_g_map']]], trans_posteriors=True)
        assert np.abs(gam.double().cpu().numpy() - b[tag + '_gamma']).max() <= 1e-5
        assert np.abs(x.double().cpu().numpy() - b[tag + '_xi']).max() <= 1e-5


@pytest.mark.parametrize('tag', ['bg', 'un'])
def test_bigram_and_uneven_phoneloop(beer, tag):
    """BigramPhoneLoop (phoneloop.py:105-191) and a PhoneLoop whose units have different lengths: the models that
    read the transition posteriors themselves (golden 'bigram_phoneloop', two VB iterations over two utterances)."""
    g = load_golden('bigram_phoneloop')
    cg = compiled(beer, g, tag + '_g0_')          # the graph before the model's weight callback ran
    D = g[tag + '_X1'].shape[1]
    ns = normalset(beer, g, cg.n_states, D, prior=tag + '_prior_', post=tag + '_post0_')
    start_pdf = {f'u{i}': int(s) for i, s in enumerate(g[tag + '_start_idxs'])}
    end_pdf = {f'u{i}': int(s) for i, s in enumerate(g[tag + '_end_idxs'])}
    cls = beer.BigramPhoneLoop if tag == 'bg' else beer.PhoneLoop
    pl = cls.create(cg, start_pdf, end_pdf, ns, prior_strength=1.)
    wu = (pl.categoricalset if tag == 'bg' else pl.categorical).weights
    np.testing.assert_allclose(wu.posterior.params.concentrations.cpu().numpy(), g[tag + '_u_dpost0'], rtol=1e-6)
    np.testing.assert_allclose(pl.graph.trans_log_probs.numpy(), g[tag + '_g_trans'], rtol=1e-5, atol=1e-6)
    X1, X2 = t32(g[tag + '_X1']), t32(g[tag + '_X2'])
    stats = pl.sufficient_statistics(X1)
    exp_llh = pl.expected_log_likelihood(stats)
    np.testing.assert_allclose(exp_llh.double().cpu().numpy(), g[tag + '_exp_llh'], rtol=1e-5, atol=1e-4)
    assert np.abs(pl.cache['trans_resps'].double().cpu().numpy() - g[tag + '_xi_block']).max() <= 1e-5
    acc = pl.accumulate(stats)
    np.testing.assert_allclose(acc[wu].cpu().numpy(), g[tag + '_acc_units'], rtol=2e-5, atol=2e-5)
    want = g[tag + '_acc_normal']
    assert np.abs(acc[ns.means_precisions].cpu().numpy() - want).max() <= 3e-5 * np.abs(want).max()
    pl.clear_cache()
    # Viterbi training: one-hot transition posteriors (hmm.py:49-54) sum to T - 1 transitions in the block at most
    pl.expected_log_likelihood(stats, viterbi=True)
    assert pl.accumulate(stats)[wu].sum().item() <= 2 * (len(X1) - 1) + 2
    pl.clear_cache()

    optim = beer.VBConjugateOptimizer(pl.conjugate_bayesian_parameters(keepgroups=True), lrate=1.)
    N = len(X1) + len(X2)
    elbos = []
    for _ in range(2):
        optim.init_step()
        elbo = beer.evidence_lower_bound(datasize=N)
        for X in (X1, X2):
            elbo += beer.evidence_lower_bound(pl, X, datasize=N)
        elbo.backward()
        elbos.append(float(elbo))
        optim.step()
    np.testing.assert_allclose(elbos, g[tag + '_elbos'], rtol=1e-5)
    np.testing.assert_allclose(wu.posterior.params.concentrations.cpu().numpy(), g[tag + '_u_dpost2'], rtol=2e-4)
    got, want = pl.graph.trans_log_probs.numpy(), g[tag + '_trans2']
    fin = np.isfinite(want)
    assert np.array_equal(np.isfinite(got), fin)
    np.testing.assert_allclose(got[fin], want[fin], rtol=2e-4, atol=2e-4)
    for gotp, pname in zip(get_ng(ns.means_precisions.posterior), ('mean', 'scale', 'shape', 'rates')):
        wantp = g[tag + '_post2_' + pname]
        np.testing.assert_allclose(gotp.reshape(wantp.shape), wantp, rtol=3e-4, atol=3e-4)


@pytest.mark.parametrize('tag,viterbi', [('bg', False), ('un', False), ('bg', True), ('un', True)])
def test_engine_bigram_and_uneven_phoneloop(beer, tag, viterbi):
    """The batched engine on the phone loops that read the transition posteriors themselves (BigramPhoneLoop,
    phoneloop.py:105-191; PhoneLoop with units of different lengths): both utterances in one launch per kernel, the
    ends x starts block summed over the frames (VBEngine block counting) -- ELBOs, unit-weight posteriors, rewritten
    transitions and Normal-Gamma posteriors after two iterations against the live-reference golden 'bigram_phoneloop'.
    viterbi: the block of the one-hot transition posteriors of the best paths (hmm.py:49-54), against the model API
    driven utterance by utterance (BigramPhoneLoop / PhoneLoop of beer_b200.models, themselves pinned by the golden)."""
    g = load_golden('bigram_phoneloop')
    cg = compiled(beer, g, tag + '_g0_')
    D = g[tag + '_X1'].shape[1]
    ns = normalset(beer, g, cg.n_states, D, prior=tag + '_prior_', post=tag + '_post0_')
    start_pdf = {f'u{i}': int(s) for i, s in enumerate(g[tag + '_start_idxs'])}
    end_pdf = {f'u{i}': int(s) for i, s in enumerate(g[tag + '_end_idxs'])}
    # the unit weights stand alone (as in hmm_train): no model callback rewrites the graph behind the engine -- with a
    # one-state unit (start == end) the rewrite of phoneloop.py:53-65 is not idempotent, it must run once per update
    n = len(start_pdf)
    starts, ends = list(start_pdf.values()), list(end_pdf.values())
    if tag == 'bg':
        cs = beer.CategoricalSet.create(torch.ones(n, n, device=DEV) / n, 1.)
        units, wu = beer.BigramUnitWeights(cs, cg, starts, ends), cs.weights
    else:
        cat = beer.Categorical.create(torch.ones(n, device=DEV) / n, 1.)
        units, wu = beer.CategoricalUnitWeights(cat, cg, starts, ends), cat.weights
    units.rewrite_graph()                        # what the model's constructor does (phoneloop.py:49-50, 142-143)
    np.testing.assert_allclose(wu.posterior.params.concentrations.cpu().numpy(), g[tag + '_u_dpost0'], rtol=1e-6)
    np.testing.assert_allclose(cg.trans_log_probs.numpy(), g[tag + '_g_trans'], rtol=1e-5, atol=1e-6)

    def flat(dist):          # (mean [K, D], scale [K], shape [K], rates [K, D]) as the engine holds them
        K = cg.n_states
        return tuple(getattr(dist.params, n).detach().to(DEV, torch.float32).reshape((K, D) if n in ('mean', 'rates') else (K,))
                     .contiguous().clone() for n in ('mean', 'scale', 'shape', 'rates'))

    em = beer.EmissionParams(flat(ns.means_precisions.prior), flat(ns.means_precisions.posterior))
    X1, X2 = t32(g[tag + '_X1']), t32(g[tag + '_X2'])
    N = len(X1) + len(X2)
    eng = beer.VBEngine(em, cg.plan(n_pdfs=em.Kp), beer.Utterances(torch.cat([X1, X2]), [len(X1), len(X2)]),
                        datasize=float(N), distributed=False, unit_weights=units, viterbi=viterbi)
    elbos = [float(eng.step().item()) for _ in range(2)]
    if viterbi:
        cg_m = compiled(beer, g, tag + '_g0_')
        ns_m = normalset(beer, g, cg_m.n_states, D, prior=tag + '_prior_', post=tag + '_post0_')
        pl = (beer.BigramPhoneLoop if tag == 'bg' else beer.PhoneLoop).create(cg_m, start_pdf, end_pdf, ns_m, prior_strength=1.)
        optim = beer.VBConjugateOptimizer(pl.conjugate_bayesian_parameters(keepgroups=True), lrate=1.)
        want_elbos = []
        for _ in range(2):
            optim.init_step()
            elbo = beer.evidence_lower_bound(datasize=N)
            for X in (X1, X2):
                elbo += beer.evidence_lower_bound(pl, X, datasize=N, viterbi=True)
            elbo.backward()
            want_elbos.append(float(elbo))
            optim.step()
        want_conc = (pl.categoricalset if tag == 'bg' else pl.categorical).weights.posterior.params.concentrations.cpu().numpy()
        want = cg_m.trans_log_probs.numpy()
        want_post = get_ng(ns_m.means_precisions.posterior)
    else:
        want_elbos, want_conc, want = g[tag + '_elbos'], g[tag + '_u_dpost2'], g[tag + '_trans2']
        want_post = [g[tag + '_post2_' + pname] for pname in ('mean', 'scale', 'shape', 'rates')]
    np.testing.assert_allclose(elbos, want_elbos, rtol=1e-5)
    np.testing.assert_allclose(wu.posterior.params.concentrations.cpu().numpy(), want_conc, rtol=2e-4)
    got = cg.trans_log_probs.numpy()
    fin = np.isfinite(want)
    assert np.array_equal(np.isfinite(got), fin)
    np.testing.assert_allclose(got[fin], want[fin], rtol=2e-4, atol=2e-4)
    for gotp, wantp in zip(em.post, want_post):
        np.testing.assert_allclose(gotp.double().cpu().numpy().reshape(wantp.shape), wantp, rtol=3e-4, atol=3e-4)


def test_mixture_cfg5_shape(beer):
    """BASELINE configs[4] shape: the E-step + M-step of a 512-component diagonal GMM on 40-d frames (the inner
    model of the GSM-GMM example), two VB iterations against the oracle (mixture.py:70-102)."""
    from oracle import beer_oracle as O
    Cn, D, N = 512, 40, 6000
    gen = torch.Generator().manual_seed(11)
    centres = 3.0 * torch.randn(32, D, generator=gen)
    X = (centres[torch.randint(0, 32, (N,), generator=gen)] + torch.randn(N, D, generator=gen)).to(DEV)
    ns = beer.NormalSet.create(torch.zeros(D, device=DEV), torch.ones(D, device=DEV), size=Cn, prior_strength=1.,
                               noise_std=1., cov_type='diagonal')
    gmm = beer.Mixture.create(ns)
    par, w = ns.means_precisions, gmm.categorical.weights

    def host(dist):
        m, k, a, b = get_ng(dist)
        return m, k.reshape(-1, 1), a.reshape(-1, 1), b
    ng_prior, ng_post = host(par.prior), host(par.posterior)
    dprior = w.prior.params.concentrations.double().cpu().numpy()
    dpost = w.posterior.params.concentrations.double().cpu().numpy()
    Xh = X.double().cpu().numpy()
    optim = beer.VBConjugateOptimizer(gmm.mean_field_factorization(), lrate=1.)
    for _ in range(2):
        r = O.gmm_estep(Xh, ng_post, dpost)
        kl = O.normalgamma_kl(ng_post, ng_prior).sum() + O.dirichlet_kl(dpost, dprior).sum()
        want = O.elbo_value(r['exp_llh'], kl, N)
        optim.init_step()
        elbo = beer.evidence_lower_bound(gmm, X, datasize=N)
        elbo.backward()
        optim.step()
        np.testing.assert_allclose(float(elbo), want, rtol=1e-5)
        ng_post = O.natural_grad_update_normalgamma(ng_prior, ng_post, r['acc_normal'], 1.)
        dpost = O.natural_grad_update_dirichlet(dprior, dpost, r['acc_dirichlet'], 1.)
    for got, want in zip(host(par.posterior), ng_post):
        np.testing.assert_allclose(got, want, rtol=3e-4, atol=3e-4)
    np.testing.assert_allclose(w.posterior.params.concentrations.double().cpu().numpy(), dpost, rtol=3e-4, atol=1e-5)


@pytest.mark.parametrize('hyper', [False, True])
def test_stick_breaking_phoneloop(beer, hyper):
    """PhoneLoop.create(..., categorical=SBCategorical[HyperPrior].create(P, ...)) — the unit-weight priors of the CLI
    (mkphoneloop.py; `gamma_dirichlet_process`, its default, is the hyper-prior variant) — against the live-reference
    goldens: three accumulate/update iterations."""
    g = load_golden('sb_hyper_phoneloop' if hyper else 'sb_phoneloop')
    cg = compiled(beer, g, 'g0_')
    D = g['X1'].shape[1]
    ns = normalset(beer, g, cg.n_states, D)
    start_pdf = {f'u{i}': int(s) for i, s in enumerate(g['start_idxs'])}
    end_pdf = {f'u{i}': int(s) for i, s in enumerate(g['end_idxs'])}
    if hyper:
        sb = beer.SBCategoricalHyperPrior.create(len(start_pdf), prior_strength=2., hyper_prior_strength=1., device=DEV)
    else:
        sb = beer.SBCategorical.create(len(start_pdf), prior_strength=2., device=DEV)
    pl = beer.PhoneLoop.create(cg, start_pdf, end_pdf, ns, categorical=sb)
    w = sb.stickbreaking
    np.testing.assert_allclose(w.prior.params.concentrations.cpu().numpy(), g['sb_prior'], rtol=1e-6)
    np.testing.assert_allclose(sb.mean.cpu().numpy(), g['mean0'], rtol=1e-6)
    np.testing.assert_allclose(pl.graph.trans_log_probs.numpy(), g['g_trans'], rtol=1e-5, atol=1e-6)
    X1, X2 = t32(g['X1']), t32(g['X2'])
    optim = beer.VBConjugateOptimizer(pl.conjugate_bayesian_parameters(keepgroups=True), lrate=1.)
    N = len(X1) + len(X2)
    elbos = []
    for it in range(3):
        optim.init_step()
        elbo = beer.evidence_lower_bound(datasize=N)
        for X in (X1, X2):
            elbo += beer.evidence_lower_bound(pl, X, datasize=N)
        elbo.backward()
        elbos.append(float(elbo))
        optim.step()
        np.testing.assert_array_equal(sb.ordering.cpu().numpy(), g[f'it{it + 1}_ordering'])
        np.testing.assert_allclose(w.posterior.params.concentrations.cpu().numpy(), g[f'it{it + 1}_sb_post'], rtol=2e-4)
        if hyper:
            cp = sb.concentration.posterior.params
            np.testing.assert_allclose([float(cp.shape), float(cp.rate)], g[f'it{it + 1}_conc'], rtol=2e-4)
            np.testing.assert_allclose(w.prior.params.concentrations.cpu().numpy(), g[f'it{it + 1}_sb_prior'], rtol=2e-4)
        got, want = pl.graph.trans_log_probs.numpy(), g[f'it{it + 1}_trans']
        fin = np.isfinite(want)
        assert np.array_equal(np.isfinite(got), fin)
        np.testing.assert_allclose(got[fin], want[fin], rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(elbos, g['elbos'], rtol=1e-5)
    np.testing.assert_allclose(sb.mean.cpu().numpy(), g['mean3'], rtol=2e-4)


@pytest.mark.parametrize('name', ['hmm_small', 'hmm_scaled'])
def test_gradient_wrt_frames(beer, name):
    """`(exp_llh * w).sum().backward()` reaches the frames with the posteriors held fixed (hmm.py:79-87), the gradient
    an encoder in front of the HMM receives (HMM-VAE): against the live reference's autograd."""
    g, gg = load_golden(name), load_golden('hmm_input_grad')
    graph = compiled(beer, g)
    D = g['X'].shape[1]
    hmm = beer.HMM.create(graph, normalset(beer, g, graph.n_states, D))
    X = t32(g['X']).requires_grad_(True)
    exp_llh = hmm.expected_log_likelihood(hmm.sufficient_statistics(X), inference_graph=graph, scale=float(g['scale']))
    np.testing.assert_allclose(exp_llh.detach().double().cpu().numpy(), g['exp_llh'], rtol=1e-5, atol=1e-4)
    (exp_llh * t32(gg[name + '_upstream'])).sum().backward()
    want = gg[name + '_grad']
    assert np.abs(X.grad.double().cpu().numpy() - want).max() <= 2e-5 * np.abs(want).max()
    # the VB statistics of the same call are untouched by the autograd path
    acc = hmm.accumulate(None)
    par = hmm.modelset.original_modelset.means_precisions
    assert np.abs(acc[par].cpu().numpy() - g['acc_normal']).max() <= 3e-5 * np.abs(g['acc_normal']).max()


@pytest.mark.parametrize('tag', ['free', 'labels'])
def test_gradient_wrt_frames_mixture(beer, tag):
    """Mixture.expected_log_likelihood back-propagated to the frames (mixture.py:76-93: detached responsibilities x
    attached per-component llhs; the path a VAE encoder sees with a GMM prior, vae.py:63-89) against the live
    reference's autograd: 200 components (ragged chunk), D = 40, 300 frames (ragged tile) through the tcgen05
    backward kernel (csrc/emission_bwd.cu)."""
    g = load_golden('mixture_input_grad')
    C, D = g['post_mean'].shape
    ns = beer.NormalSet.create(torch.zeros(D, device=DEV), torch.ones(D, device=DEV), size=C, prior_strength=1.,
                               noise_std=1., cov_type='diagonal')
    set_ng(ns.means_precisions.posterior, g, 'post_')
    gmm = beer.Mixture.create(ns)
    gmm.categorical.weights.posterior.params.concentrations.copy_(t32(g['dpost']))
    X = t32(g['X']).requires_grad_(True)
    kw = {'labels': torch.from_numpy(g['labels'])} if tag == 'labels' else {}
    exp_llh = gmm.expected_log_likelihood(gmm.sufficient_statistics(X), **kw)
    np.testing.assert_allclose(exp_llh.detach().double().cpu().numpy(), g[tag + '_exp_llh'], rtol=1e-5, atol=1e-4)
    (exp_llh * t32(g['upstream'])).sum().backward()
    want = g[tag + '_grad']
    assert np.abs(X.grad.double().cpu().numpy() - want).max() <= 2e-5 * np.abs(want).max()


def test_gradient_wrt_frames_hmm_mixtureset(beer):
    """HMM over a MixtureSet: the gradient of sum_t go_t sum_k gamma_tk sum_c r_tkc llh_kc(x_t) w.r.t. the frames
    (posteriors and responsibilities held fixed) -- the mixture form of hmm.py:79-87.  The reference itself has no
    gradient here (mixtureset.py:92-98 returns a detached log-normaliser), so the check is the definition evaluated in
    fp64 from the posteriors / llhs the forward call cached."""
    from beer_b200 import ops, synthetic
    P, S, Cn, D, T = 6, 3, 4, 20, 170
    K, M = P * S, P * S * Cn
    graph, _, _ = synthetic.phone_loop_graph(P, S)
    means = 2.0 * torch.randn(K, D, generator=torch.Generator().manual_seed(0))
    X = synthetic.sample_utterances(graph, means, 1, T, seed=1, device=DEV).reshape(T, D)
    ns = beer.NormalSet.create(torch.zeros(D, device=DEV), torch.ones(D, device=DEV), size=M, prior_strength=1.,
                               noise_std=2., cov_type='diagonal')
    hmm = beer.HMM.create(graph, beer.MixtureSet.create(K, ns, prior_strength=1.))
    Xg = X.clone().requires_grad_(True)
    scale = 0.7
    exp_llh = hmm.expected_log_likelihood(hmm.sufficient_statistics(Xg), scale=scale)
    up = torch.linspace(0.5, 1.5, T, device=DEV)
    (exp_llh * up).sum().backward()
    c = hmm.cache
    w = (c['pdf_post'].double().repeat_interleave(Cn, dim=1)
         * torch.exp(c['comp_llh'].double() - c['pdf_llh'].double().repeat_interleave(Cn, dim=1)))
    ets = ops.normalgamma_expected_stats(*[t.float().contiguous() for t in
                                           ns.means_precisions.posterior.params.as_tuple()]).double()
    want = up.double()[:, None] * (w @ ets[:, :D] - X.double() * (w @ ets[:, D:2 * D]))
    assert float(w.sum()) == pytest.approx(scale * T, rel=1e-5)
    assert float((Xg.grad.double() - want).abs().max()) <= 2e-5 * float(want.abs().max())


class _Net(torch.nn.Sequential):
    def __init__(self, dim_in, dim_out):
        super().__init__(torch.nn.Linear(dim_in, dim_out), torch.nn.Tanh())
        self.dim_in, self.dim_out = dim_in, dim_out


@pytest.mark.parametrize('tag', ['hmm', 'gmm'])
def test_vae_against_reference(beer, tag, monkeypatch):
    """VAE (vae.py:27-89) over an HMM / a GMM prior: one evidence_lower_bound + backward() with the reference's
    reparameterisation noise replayed.  The value matrix, the ELBO, the statistics of the prior and the gradient of every
    network parameter -- which reaches the encoder through prior.expected_log_likelihood, i.e. through the backward
    kernel of the emission llhs -- against the live reference's fp64 run."""
    g = load_golden('vae')
    X = t32(g['X'])
    N, D = X.shape
    L = g[tag + '_post_mean'].shape[1]
    if tag == 'hmm':
        graph = compiled(beer, g, 'hmm_g_')
        ns = normalset(beer, g, graph.n_states, L, prior='hmm_prior_', post='hmm_post_')
        prior = beer.HMM.create(graph, ns)
    else:
        ns = normalset(beer, g, g['gmm_post_mean'].shape[0], L, prior='gmm_prior_', post='gmm_post_')
        prior = beer.Mixture.create(ns)
    H = g[tag + '_sd_encoder.0.bias'].shape[0]
    vae = beer.VAE(prior, _Net(D, H), _Net(L, H)).to(DEV)
    sd = {k[len(tag) + 4:]: t32(g[k]) for k in g if k.startswith(tag + '_sd_')}
    missing = vae.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys
    noise = t32(g[tag + '_noise'])
    monkeypatch.setattr(torch, 'randn', lambda *a, **k: noise.clone())
    value = vae.expected_log_likelihood(vae.sufficient_statistics(X))
    want = g[tag + '_value']
    assert value.shape == want.shape                     # [N, N]: the reference's broadcast (vae.py:86)
    assert np.abs(value.detach().double().cpu().numpy() - want).max() <= 2e-5 * np.abs(want).max()
    vae.clear_cache()
    elbo = beer.evidence_lower_bound(vae, X, datasize=N)
    np.testing.assert_allclose(float(elbo), float(g[tag + '_elbo']), rtol=2e-5)
    elbo.backward()
    acc = elbo._acc_stats[ns.means_precisions].cpu().numpy()
    assert np.abs(acc - g[tag + '_acc_normal']).max() <= 3e-5 * np.abs(g[tag + '_acc_normal']).max()
    n_checked = 0
    for name, p in vae.named_parameters():
        key = f'{tag}_grad_{name}'
        if key in g:
            want = g[key]
            assert p.grad is not None, name
            assert np.abs(p.grad.double().cpu().numpy() - want).max() <= 1e-4 * np.abs(want).max(), name
            n_checked += 1
    assert n_checked == 12


@pytest.mark.parametrize('P,S', [(96, 4), (130, 3), (250, 4)])
def test_phoneloop_unit_counts_many_units(beer, P, S):
    """Unit counts of phone loops around the limits of the fused reduction (phoneloop.py:83-101): 96 units and the
    250 x 4 = 1000-state graph of BASELINE configs[2] run one-warp left-to-right kernels with the counts fused into
    the backward sweep; 130 units of three states have no counting kernel, `n_units` reports 0 and the model reduces
    the ends x starts block of the transition posteriors instead.  All against the oracle's dense xi."""
    from beer_b200 import ops, synthetic
    from oracle import beer_oracle as O
    D, T = 8, 37
    K = P * S
    graph, starts, ends = synthetic.phone_loop_graph(P, S)
    means = 2.0 * torch.randn(K, D, generator=torch.Generator().manual_seed(3))
    X = synthetic.sample_utterances(graph, means, 1, T, seed=5, device=DEV)
    ns = beer.NormalSet.create(torch.zeros(D, device=DEV), torch.ones(D, device=DEV), size=K, prior_strength=1.,
                               noise_std=1., cov_type='diagonal')
    start_pdf = {f'u{i}': int(s) for i, s in enumerate(starts)}
    end_pdf = {f'u{i}': int(s) for i, s in enumerate(ends)}
    pl = beer.PhoneLoop.create(graph, start_pdf, end_pdf, ns, prior_strength=1.)
    plan = pl.graph.plan(n_pdfs=K)
    assert plan.n_units == (P if (P <= 128 or S == 4) else 0)
    if plan.n_units == 0:
        with pytest.raises(beer._lib.BeerB200Error):        # never dropped silently
            pdf = torch.zeros(T, K, device=DEV)
            ops.hmm_forward_backward(plan, pdf, None, torch.tensor([0, T], device=DEV),
                                     unit_counts=torch.zeros(P, dtype=torch.float64, device=DEV))
    stats = pl.sufficient_statistics(X)
    pl.expected_log_likelihood(stats)
    wu = pl.categorical.weights
    got = pl.accumulate(stats)[wu].cpu().numpy()
    post = tuple(a[:, None] if a.ndim == 1 else a for a in get_ng(ns.means_precisions.posterior))
    post = (post[0], post[1].reshape(-1, 1), post[2].reshape(-1, 1), post[3])
    og = (pl.graph.init_log_probs.double().numpy(), pl.graph.final_log_probs.double().numpy(),
          pl.graph.trans_log_probs.double().numpy(), pl.graph.pdf_id_mapping)
    r = O.hmm_estep(X.double().cpu().numpy(), post, None, og, trans_posteriors=True)
    want = O.phoneloop_counts(r['gamma'], r['xi'], starts, ends)
    np.testing.assert_allclose(got, want, rtol=3e-5, atol=3e-5)
