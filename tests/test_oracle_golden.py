"""Pin the numpy oracle to the fp64 outputs of the live reference (tests/golden/*.npz)."""
import numpy as np
import pytest

from oracle import beer_oracle as O
from conftest import load_golden

TOL = dict(rtol=1e-9, atol=1e-10)


def ng(g, p):
    return g[p + 'mean'], g[p + 'scale'], g[p + 'shape'], g[p + 'rates']


def graph(g, p='g_'):
    return g[p + 'init'], g[p + 'final'], g[p + 'trans'], g[p + 'map']


def test_normalgamma_and_dirichlet():
    g = load_golden('dists')
    q, p = ng(g, 'ng_'), ng(g, 'ngp_')
    np.testing.assert_allclose(O.normalgamma_natural_parameters(*q), g['ng_nat'], **TOL)
    np.testing.assert_allclose(O.normalgamma_expected_sufficient_statistics(*q), g['ng_ets'], **TOL)
    np.testing.assert_allclose(O.normalgamma_log_norm(*q), g['ng_lognorm'], **TOL)
    np.testing.assert_allclose(O.normalgamma_kl(q, p), g['ng_kl'], **TOL)
    back = O.normalgamma_from_natural_parameters(g['ng_nat'])
    for a, k in zip(back, ('mean', 'scale', 'shape', 'rates')):
        np.testing.assert_allclose(a, g['ng_back_' + k], **TOL)
    np.testing.assert_allclose(O.normal_diag_sufficient_statistics(g['X']), g['stats'], **TOL)
    np.testing.assert_allclose(O.normal_diag_llh(g['stats'], g['ng_ets'], g['X'].shape[1]), g['llh'], **TOL)
    c, cp = g['dir_conc'], g['dirp_conc']
    np.testing.assert_allclose(O.dirichlet_natural_parameters(c), g['dir_nat'], **TOL)
    np.testing.assert_allclose(O.dirichlet_expected_sufficient_statistics(c), g['dir_ets'], **TOL)
    np.testing.assert_allclose(O.dirichlet_log_norm(c), g['dir_lognorm'], **TOL)
    np.testing.assert_allclose(O.dirichlet_kl(c, cp), g['dir_kl'], **TOL)
    np.testing.assert_allclose(O.dirichlet_from_natural_parameters(g['dir_nat']), g['dir_back'], **TOL)
    np.testing.assert_allclose(O.categorical_sufficient_statistics(g['cat_data']), g['cat_stats'], **TOL)
    np.testing.assert_allclose(O.categorical_log_weights(c), g['dir_logw'], **TOL)


def test_gmm_cfg1():
    g = load_golden('gmm_cfg1')
    X = g['X'].astype(np.float64)
    prior, post = ng(g, 'prior_'), ng(g, 'post0_')
    dprior, dpost = g['dprior'], g['dpost0']
    r = O.gmm_estep(X, post, dpost)
    np.testing.assert_allclose(r['exp_llh'], g['exp_llh'], **TOL)
    np.testing.assert_allclose(r['resps'], g['resps'], **TOL)
    np.testing.assert_allclose(r['acc_normal'], g['acc_normal'], **TOL)
    np.testing.assert_allclose(r['acc_dirichlet'], g['acc_dirichlet'], **TOL)
    kl = O.normalgamma_kl(post, prior).sum() + O.dirichlet_kl(dpost, dprior).sum()
    np.testing.assert_allclose(kl, g['kl'], **TOL)
    rl = O.gmm_estep(X, post, dpost, labels=g['labels'])
    np.testing.assert_allclose(rl['exp_llh'], g['exp_llh_labels'], **TOL)
    elbos = []
    for _ in range(6):
        r = O.gmm_estep(X, post, dpost)
        kl = O.normalgamma_kl(post, prior).sum() + O.dirichlet_kl(dpost, dprior).sum()
        elbos.append(O.elbo_value(r['exp_llh'], kl, len(X)))
        post = O.natural_grad_update_normalgamma(prior, post, r['acc_normal'], 1.)
        dpost = O.natural_grad_update_dirichlet(dprior, dpost, r['acc_dirichlet'], 1.)
    np.testing.assert_allclose(elbos, g['elbos'], rtol=1e-9)
    for a, k in zip(post, ('mean', 'scale', 'shape', 'rates')):
        np.testing.assert_allclose(a, g['post6_' + k], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(dpost, g['dpost6'], rtol=1e-8)


def _check_hmm(name, n_iter):
    g = load_golden(name)
    X = g['X'].astype(np.float64)
    scale = float(g['scale'])
    prior, post = ng(g, 'prior_'), ng(g, 'post0_')
    gr = graph(g)
    r = O.hmm_estep(X, post, None, gr, scale=scale)
    np.testing.assert_allclose(r['pdf_llh'], g['pdf_llh'], **TOL)
    np.testing.assert_allclose(r['gamma'], g['gamma'], rtol=1e-7, atol=1e-12)
    np.testing.assert_allclose(r['exp_llh'], g['exp_llh'], rtol=1e-9)
    np.testing.assert_allclose(r['acc_normal'], g['acc_normal'], rtol=1e-7, atol=1e-9)
    kl = O.normalgamma_kl(post, prior).sum()
    np.testing.assert_allclose(kl, g['kl'], **TOL)
    np.testing.assert_allclose(O.elbo_value(r['exp_llh'], kl, 3 * len(X)), g['elbo_datasize3T'], rtol=1e-9)
    pc = scale * r['pdf_llh'][:, gr[3]]
    path = O.best_path(pc, *gr[:3])
    np.testing.assert_array_equal(path, g['viterbi_path'])
    np.testing.assert_array_equal(gr[3][path], g['decode'])
    rv = O.hmm_estep(X, post, None, gr, scale=scale, viterbi=True)
    np.testing.assert_allclose(rv['exp_llh'], g['exp_llh_viterbi'], rtol=1e-9)
    np.testing.assert_allclose(rv['acc_normal'], g['acc_normal_viterbi'], rtol=1e-9, atol=1e-12)
    # HMM.posteriors scales the statistics, not the llhs (hmm.py:119)
    stats = O.normal_diag_sufficient_statistics(X) * 1.0
    ets = O.normalgamma_expected_sufficient_statistics(*post)
    pcp = O.normal_diag_llh(stats, ets, X.shape[1])[:, gr[3]]
    gam, _ = O.posteriors(pcp, *gr[:3])
    np.testing.assert_allclose(gam, g['posteriors'], rtol=1e-7, atol=1e-12)
    elbos = []
    for _ in range(n_iter):
        e, post, _, _ = O.vb_iteration_hmm([X], prior, post, None, None, gr, scale=scale)
        elbos.append(e)
    np.testing.assert_allclose(elbos, g['elbos'], rtol=1e-9)
    for a, k in zip(post, ('mean', 'scale', 'shape', 'rates')):
        np.testing.assert_allclose(a, g[f'post{n_iter}_' + k], rtol=1e-7, atol=1e-9)


def test_hmm_small():
    _check_hmm('hmm_small', 3)


def test_hmm_scaled():
    _check_hmm('hmm_scaled', 3)


def test_hmm_cfg2_T200():
    _check_hmm('hmm_cfg2_T200', 2)


def test_dense_ergodic():
    g = load_golden('dense_ergodic')
    gr = graph(g)
    (gam, xi), ll = O.posteriors(g['llhs'], *gr[:3], trans_posteriors=True)
    np.testing.assert_allclose(gam, g['gamma'], rtol=1e-9, atol=1e-14)
    np.testing.assert_allclose(xi.sum(0), g['xi_sum'], rtol=1e-9, atol=1e-14)
    np.testing.assert_allclose(ll, g['lognorm_mean'], rtol=1e-12)
    np.testing.assert_array_equal(O.best_path(g['llhs'], *gr[:3]), g['path'])
    gr2 = graph(g, 'g2_')
    gam2, _ = O.posteriors(g['llhs2'], *gr2[:3])
    np.testing.assert_allclose(gam2, g['gamma2'], rtol=1e-9, atol=1e-14)
    np.testing.assert_array_equal(O.best_path(g['llhs2'], *gr2[:3]), g['path2'])


def test_graph_compile():
    g = load_golden('graph_compile')
    gr, starts, ends = O.phone_loop_graph(5, 3)
    for a, k in zip(gr, ('init', 'final', 'trans', 'map')):
        np.testing.assert_allclose(np.asarray(a, dtype=np.float64), g['pl_' + k], rtol=1e-6)
    assert starts == [0, 3, 6, 9, 12] and ends == [2, 5, 8, 11, 14]


def test_phoneloop_mixtureset():
    g = load_golden('phoneloop_mixtureset')
    C1, C2, K1 = int(g['C1']), int(g['C2']), int(g['K1'])
    gr = graph(g)
    K = len(gr[3])

    def estep(X, posts, dposts, gr, scale=1., xi=False):
        # JointModelSet of two MixtureSets with different C (modelset.py:71-85)
        l1, r1 = O.emission_llh(X, posts[0], dposts[0])
        l2, r2 = O.emission_llh(X, posts[1], dposts[1])
        pdf_llh = np.concatenate([l1, l2], axis=1)
        m = np.asarray(gr[3])
        pc = scale * pdf_llh[:, m]
        res, _ = O.posteriors(pc, *gr[:3], trans_posteriors=xi)
        gam, x = res if xi else (res, None)
        gp = np.zeros_like(pdf_llh)
        for i in range(len(m)):
            gp[:, m[i]] += scale * gam[:, i]
        st = O.normal_diag_sufficient_statistics(X)
        j1 = r1 * gp[:, :K1, None]
        j2 = r2 * gp[:, K1:, None]
        return dict(exp_llh=(pc * gam).sum(-1), gamma=gam, xi=x, pdf_llh=pdf_llh,
                    a1=j1.reshape(len(X), -1).T @ st, a2=j2.reshape(len(X), -1).T @ st,
                    d1=O.categorical_sufficient_statistics(j1).sum(0),
                    d2=O.categorical_sufficient_statistics(j2).sum(0))

    posts = [ng(g, 'g1_post0_'), ng(g, 'g2_post0_')]
    priors = [ng(g, 'g1_prior_'), ng(g, 'g2_prior_')]
    dposts = [g['g1_dpost0'], g['g2_dpost0']]
    dpriors = [g['g1_dprior'], g['g2_dprior']]
    upost, uprior = g['u_dpost0'], g['u_dprior']
    starts, ends = list(g['start_idxs']), list(g['end_idxs'])
    r = estep(g['X1'].astype(np.float64), posts, dposts, gr, xi=True)
    np.testing.assert_allclose(r['pdf_llh'], g['u1_pdf_llh'], **TOL)
    np.testing.assert_allclose(r['gamma'], g['u1_gamma'], rtol=1e-7, atol=1e-12)
    np.testing.assert_allclose(r['xi'].sum(0), g['u1_xi_sum'], rtol=1e-7, atol=1e-12)
    np.testing.assert_allclose(r['exp_llh'], g['u1_exp_llh'], rtol=1e-9)
    np.testing.assert_allclose(r['a1'], g['u1_acc_g1'], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(r['a2'], g['u1_acc_g2'], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(r['d1'], g['u1_acc_d1'], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(r['d2'], g['u1_acc_d2'], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(O.phoneloop_counts(r['gamma'], r['xi'], starts, ends), g['u1_acc_units'],
                               rtol=1e-7, atol=1e-9)
    ag = graph(g, 'ali_')
    r3 = estep(g['X3'].astype(np.float64), posts, dposts, ag, scale=0.7)
    np.testing.assert_allclose(r3['gamma'], g['u3_gamma'], rtol=1e-7, atol=1e-12)
    np.testing.assert_allclose(r3['exp_llh'], g['u3_exp_llh'], rtol=1e-9)
    np.testing.assert_allclose(r3['a1'], g['u3_acc_g1'], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(r3['d2'], g['u3_acc_d2'], rtol=1e-7, atol=1e-9)
    assert np.all(g['u3_acc_units'] == 0)
    # 3 iterations of accumulate/update over two utterances with the phone-loop weight callback
    trans = gr[2].copy()
    N = len(g['X1']) + len(g['X2'])
    elbos = []
    for it in range(3):
        kl = sum(O.normalgamma_kl(q, p).sum() for q, p in zip(posts, priors)) \
            + sum(O.dirichlet_kl(q, p).sum() for q, p in zip(dposts, dpriors)) \
            + O.dirichlet_kl(upost, uprior).sum()
        tot, a1, a2, d1, d2, du, frames = 0., 0., 0., 0., 0., 0., 0
        for X in (g['X1'], g['X2']):
            X = X.astype(np.float64)
            r = estep(X, posts, dposts, (gr[0], gr[1], trans, gr[3]), xi=True)
            tot += O.elbo_value(r['exp_llh'], kl, N)
            a1, a2, d1, d2 = a1 + r['a1'], a2 + r['a2'], d1 + r['d1'], d2 + r['d2']
            du = du + O.phoneloop_counts(r['gamma'], r['xi'], starts, ends)
            frames += len(X)
        elbos.append(tot)
        s = N / frames
        posts = [O.natural_grad_update_normalgamma(priors[0], posts[0], s * a1, 1.),
                 O.natural_grad_update_normalgamma(priors[1], posts[1], s * a2, 1.)]
        dposts = [O.natural_grad_update_dirichlet(dpriors[0], dposts[0], s * d1, 1.),
                  O.natural_grad_update_dirichlet(dpriors[1], dposts[1], s * d2, 1.)]
        upost = O.natural_grad_update_dirichlet(uprior, upost, s * du, 1.)
        trans = O.phoneloop_update_graph(trans, upost, starts, ends)
        np.testing.assert_allclose(trans, g[f'it{it + 1}_trans'], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(elbos, g['elbos'], rtol=1e-9)
    np.testing.assert_allclose(upost, g['u_dpost3'], rtol=1e-8)
    np.testing.assert_allclose(posts[0][0], g['g1_post3_mean'], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(dposts[1], g['g2_dpost3'], rtol=1e-8)


def test_bigram_and_uneven_phoneloop():
    """Models that read the dense transition posteriors: BigramPhoneLoop (phoneloop.py:105-191) and a
    PhoneLoop whose units differ in length, two VB iterations over two utterances each."""
    g = load_golden('bigram_phoneloop')
    for tag in ('bg', 'un'):
        gr = graph(g, tag + '_g_')
        starts, ends = list(g[tag + '_start_idxs']), list(g[tag + '_end_idxs'])
        post, prior = ng(g, tag + '_post0_'), ng(g, tag + '_prior_')
        upost, uprior = g[tag + '_u_dpost0'], g[tag + '_u_dprior']
        counts = (lambda r: O.bigram_counts(r['xi'], starts, ends)) if tag == 'bg' else \
            (lambda r: O.phoneloop_counts(r['gamma'], r['xi'], starts, ends))
        regraph = O.bigram_update_graph if tag == 'bg' else O.phoneloop_update_graph
        r = O.hmm_estep(g[tag + '_X1'].astype(np.float64), post, None, gr, trans_posteriors=True)
        np.testing.assert_allclose(r['gamma'], g[tag + '_gamma'], rtol=1e-7, atol=1e-12)
        np.testing.assert_allclose(r['xi'], g[tag + '_xi'], rtol=1e-7, atol=1e-12)
        np.testing.assert_allclose(r['exp_llh'], g[tag + '_exp_llh'], rtol=1e-9)
        np.testing.assert_allclose(r['acc_normal'], g[tag + '_acc_normal'], rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(counts(r), g[tag + '_acc_units'], rtol=1e-7, atol=1e-12)
        trans = gr[2].copy()
        N = len(g[tag + '_X1']) + len(g[tag + '_X2'])
        elbos = []
        for _ in range(2):
            kl = O.normalgamma_kl(post, prior).sum() + O.dirichlet_kl(upost, uprior).sum()
            tot, acc, du, frames = 0., 0., 0., 0
            for X in (g[tag + '_X1'], g[tag + '_X2']):
                r = O.hmm_estep(X.astype(np.float64), post, None, (gr[0], gr[1], trans, gr[3]), trans_posteriors=True)
                tot += O.elbo_value(r['exp_llh'], kl, N)
                acc, du, frames = acc + r['acc_normal'], du + counts(r), frames + len(X)
            elbos.append(tot)
            post = O.natural_grad_update_normalgamma(prior, post, N / frames * acc, 1.)
            upost = O.natural_grad_update_dirichlet(uprior, upost, N / frames * du, 1.)
            trans = regraph(trans, upost, starts, ends)
        np.testing.assert_allclose(elbos, g[tag + '_elbos'], rtol=1e-9)
        np.testing.assert_allclose(upost, g[tag + '_u_dpost2'], rtol=1e-8)
        np.testing.assert_allclose(trans, g[tag + '_trans2'], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(post[0], g[tag + '_post2_mean'], rtol=1e-7, atol=1e-9)


@pytest.mark.parametrize('hyper', [False, True])
def test_stick_breaking_phoneloop(hyper):
    """PhoneLoop with SBCategorical / SBCategoricalHyperPrior unit weights (categorical.py:82-209), three VB
    iterations over two utterances."""
    g = load_golden('sb_hyper_phoneloop' if hyper else 'sb_phoneloop')
    gr = graph(g)
    starts, ends = list(g['start_idxs']), list(g['end_idxs'])
    post, prior = ng(g, 'post0_'), ng(g, 'prior_')
    sb_post, sb_prior = g['sb_post0'], g['sb_prior'].copy()
    ordering = np.arange(len(starts))
    trans = gr[2].copy()
    N = len(g['X1']) + len(g['X2'])
    elbos = []
    for it in range(3):
        kl = O.normalgamma_kl(post, prior).sum() + O.dirichlet_kl(sb_post, sb_prior).sum()
        tot, acc, counts, frames = 0., 0., 0., 0
        for X in (g['X1'], g['X2']):
            r = O.hmm_estep(X.astype(np.float64), post, None, (gr[0], gr[1], trans, gr[3]), trans_posteriors=True)
            tot += O.elbo_value(r['exp_llh'], kl, N)
            tr = r['xi'].sum(axis=0)
            counts = counts + tr[:, starts][ends, :].sum(axis=0) + r['gamma'][0][starts]
            acc, frames = acc + r['acc_normal'], frames + len(X)
        elbos.append(tot)
        post = O.natural_grad_update_normalgamma(prior, post, N / frames * acc, 1.)
        stats, ordering = O.sb_transform_stats(N / frames * counts)
        sb_post = O.natural_grad_update_dirichlet(sb_prior, sb_post, stats, 1.)
        trans = O.sb_phoneloop_update_graph(trans, sb_post, ordering, starts, ends)
        if hyper:
            shape, rate = O.sb_hyper_update(sb_post, ordering, tuple(g['conc_prior']))
            np.testing.assert_allclose([shape, rate], g[f'it{it + 1}_conc'], rtol=1e-9)
            sb_prior[:, 1] = shape / rate
            np.testing.assert_allclose(sb_prior, g[f'it{it + 1}_sb_prior'], rtol=1e-9)
        np.testing.assert_array_equal(ordering, g[f'it{it + 1}_ordering'])
        np.testing.assert_allclose(sb_post, g[f'it{it + 1}_sb_post'], rtol=1e-8)
        np.testing.assert_allclose(trans, g[f'it{it + 1}_trans'], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(elbos, g['elbos'], rtol=1e-9)
    np.testing.assert_allclose(post[0], g['post3_mean'], rtol=1e-7, atol=1e-9)


def test_fbank_front_end():
    """beer/features.py restated (fbank, create_fbank, add_deltas) against the live-reference golden."""
    g = load_golden('fbank')
    for nf in (40, 26):
        np.testing.assert_allclose(O.create_fbank(nf, 512, lowfreq=20, highfreq=8000), g[f'filters{nf}'], atol=1e-15)
        np.testing.assert_allclose(O.fbank(g['signal'], nfilters=nf), g[f'fbank{nf}'], rtol=1e-12)
    np.testing.assert_allclose(O.add_deltas(g['fbank40']), g['deltas40'], atol=1e-12)
    # the front-end of `beer features extract` (short_term_mspec + filterbank + log(1e-6 + .))
    mspec, fft_len = O.short_term_mspec(g['signal'])
    assert fft_len == 512
    np.testing.assert_allclose(mspec, g['mspec'], rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(O.log_mel_spectrum(g['signal'], nfilters=40), g['cli_logmel40'], rtol=1e-12)
