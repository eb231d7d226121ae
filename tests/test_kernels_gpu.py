"""GPU parity tests, kernel level: every C-ABI entry point against the numpy oracle and
the fp64 goldens generated from the live reference (tests/golden/make_goldens.py).

Tolerances (north_star: "ELBO and state posteriors within 1e-5 relative in fp32"):
  * state posteriors: |gamma - gamma_ref| <= 1e-5 (posteriors live in [0, 1]);
  * expected log-likelihood sums / ELBO: 1e-5 relative (measured ~1e-7);
  * accumulated statistics: 2e-5 relative to the largest statistic of the row block;
  * Viterbi paths: identical.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import beer_oracle as O

pytestmark = pytest.mark.gpu

DEV = 'cuda'


def dev(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(DEV).contiguous()


def ng(g, p):
    return g[p + 'mean'], g[p + 'scale'], g[p + 'shape'], g[p + 'rates']


def ng_dev(t):
    m, k, a, b = t
    return dev(m), dev(k.reshape(-1)), dev(a.reshape(-1)), dev(b)


def graph(g, p='g_'):
    return g[p + 'init'], g[p + 'final'], g[p + 'trans'], g[p + 'map']


@pytest.fixture(scope='module')
def ops():
    from beer_b200 import ops
    ops.require_cuda()
    return ops


def test_normalgamma_dirichlet_math(ops):
    g = load_golden('dists')
    q, p = ng(g, 'ng_'), ng(g, 'ngp_')
    ets = ops.normalgamma_expected_stats(*ng_dev(q)).cpu().numpy()
    np.testing.assert_allclose(ets, g['ng_ets'], rtol=2e-6, atol=1e-6)
    kl = ops.normalgamma_kl(ng_dev(p), ng_dev(q)).item()
    np.testing.assert_allclose(kl, g['ng_kl'].sum(), rtol=1e-6)
    logw = ops.dirichlet_expected_logw(dev(g['dir_conc'])).cpu().numpy()
    np.testing.assert_allclose(logw, g['dir_logw'], rtol=2e-6, atol=1e-6)
    dkl = ops.dirichlet_kl(dev(g['dirp_conc']), dev(g['dir_conc'])).item()
    np.testing.assert_allclose(dkl, g['dir_kl'].sum(), rtol=1e-6)
    # natural-gradient step against the oracle (parameters.py:134-141)
    rng = np.random.default_rng(0)
    M, D = q[0].shape
    n = rng.random(M) * 20
    sx = rng.standard_normal((M, D)) * n[:, None]
    sxx = (rng.random((M, D)) + 1) * n[:, None] * 2
    acc = np.concatenate([sx, -.5 * sxx, -.5 * n[:, None], .5 * n[:, None]], axis=1)
    for lr, s in ((1.0, 1.0), (0.3, 2.5)):
        want = O.natural_grad_update_normalgamma(p, q, s * acc, lr)
        post = ng_dev(q)
        ops.normalgamma_update(ng_dev(p), post, dev(acc, torch.float64), s, lr)
        for got, w in zip(post, want):
            np.testing.assert_allclose(got.cpu().numpy().reshape(w.shape), w, rtol=1e-5, atol=1e-6)
        cacc = rng.random(g['dir_conc'].shape) * 10
        cacc[:, -1] = cacc.sum(axis=1)
        want = O.natural_grad_update_dirichlet(g['dirp_conc'], g['dir_conc'], s * cacc, lr)
        conc = dev(g['dir_conc'])
        ops.dirichlet_update(dev(g['dirp_conc']), conc, dev(cacc, torch.float64), s, lr)
        np.testing.assert_allclose(conc.cpu().numpy(), want, rtol=1e-5, atol=1e-6)


def _emission(ops, X, post, dir_post=None, comp_off=None, want_comp=False, tc=False):
    logw = None
    if dir_post is not None:
        logw = torch.cat([ops.dirichlet_expected_logw(dev(d)).reshape(-1) for d in dir_post])
    W, bias, ref = ops.emission_prepare(*ng_dev(post), logw=logw)
    co = None if comp_off is None else dev(comp_off, torch.int32)
    if tc:      # tcgen05 kernel (uniform number of Gaussians per pdf)
        M, D = post[0].shape
        C = 1 if comp_off is None else int(comp_off[1] - comp_off[0])
        if not ops.emission_tc_supported(M, D, C):
            pytest.skip('no tensor-core emission path for this shape')
        img = ops.emission_tc_pack(W, bias, C)
        return ops.emission_llh_tc(dev(X), img, ref, M, C, want_comp=want_comp)
    return ops.emission_llh(dev(X), W, bias, ref, comp_off=co, want_comp=want_comp)


@pytest.mark.parametrize('name', ['hmm_small', 'hmm_scaled', 'hmm_cfg2_T200'])
def test_emission_llh_normalset(ops, name):
    g = load_golden(name)
    pdf_llh, _, fref = _emission(ops, g['X'], ng(g, 'post0_'))
    got = pdf_llh.double().cpu().numpy() + fref.double().cpu().numpy()[:, None]
    np.testing.assert_allclose(got, g['pdf_llh'], rtol=0, atol=2e-4)   # absolute llh ~ -1e2: fp32 ulp
    # what the scan consumes is the offset form: differences between states must be accurate
    d_got = pdf_llh.double().cpu().numpy()
    d_ref = g['pdf_llh'] - g['pdf_llh'].max(axis=1, keepdims=True)
    d_got = d_got - d_got[np.arange(len(d_got)), g['pdf_llh'].argmax(axis=1)][:, None]
    near = d_ref > -30
    assert np.abs(d_got - d_ref)[near].max() < 3e-5


def test_emission_llh_wide_frames(ops):
    """A 64-d latent space (BASELINE configs[3], the HMM prior of an HMM-VAE): no tcgen05 emission tile for it, the
    SIMT kernel runs it with one block per SM.  Against the fp64 restatement of normalgamma.py:118-146 + :55-59."""
    rng = np.random.default_rng(5)
    M, D, N = 100, 64, 333
    post = (rng.standard_normal((M, D)), rng.uniform(0.5, 3, (M, 1)), rng.uniform(1, 5, (M, 1)), rng.uniform(0.3, 2, (M, D)))
    X = 2.0 * rng.standard_normal((N, D))
    pdf_llh, _, fref = _emission(ops, X, post)
    f32 = lambda a: a.astype(np.float32).astype(np.float64)          # the values the kernel is handed
    want = O.emission_llh(f32(X), tuple(f32(t) for t in post))[0]
    got = pdf_llh.double().cpu().numpy() + fref.double().cpu().numpy()[:, None]
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-6 * np.abs(want).max())


def _fb_case(ops, llh, gr, scale=1.0, factorize=True, **kw):
    plan = ops.GraphPlan(*gr, factorize=factorize)
    T = llh.shape[0]
    utt = torch.tensor([0, T], dtype=torch.int64, device=DEV)
    return plan, ops.hmm_forward_backward(plan, dev(llh), None, utt, scale=scale, want_state_post=True,
                                          want_frame_llh=True, want_logz=True, **kw)


@pytest.mark.parametrize('factorize', [True, False])
@pytest.mark.parametrize('name', ['hmm_small', 'hmm_scaled', 'hmm_cfg2_T200'])
def test_forward_backward_golden(ops, name, factorize):
    g = load_golden(name)
    scale = float(g['scale'])
    gr = graph(g)
    # feed the reference's own (fp64 -> fp32, max-subtracted) llhs: isolates the scan
    llh = g['pdf_llh'] - g['pdf_llh'].max(axis=1, keepdims=True)
    plan, r = _fb_case(ops, llh, gr, scale=scale, factorize=factorize)
    gamma = r['state_post'].double().cpu().numpy()
    assert np.abs(gamma - g['gamma']).max() <= 1e-5
    np.testing.assert_allclose(gamma.sum(axis=1), 1.0, atol=2e-6)
    exp_llh = r['frame_exp_llh'].double().cpu().numpy() + scale * g['pdf_llh'].max(axis=1)
    np.testing.assert_allclose(exp_llh.sum(), g['exp_llh'].sum(), rtol=1e-6)
    np.testing.assert_allclose(r['utt_exp_llh'].item() + scale * g['pdf_llh'].max(axis=1).sum(),
                               g['exp_llh'].sum(), rtol=1e-6)
    # log evidence against the oracle's unnormalised forward pass
    la = O.forward(scale * g['pdf_llh'], gr[0], gr[2])
    logz = O.logsumexp(la[-1] + gr[1], axis=0)
    np.testing.assert_allclose(r['utt_logz'].item() + scale * g['pdf_llh'].max(axis=1).sum(), logz, rtol=1e-6)
    if factorize and name == 'hmm_cfg2_T200':
        assert plan.info['junctions'] == 1 and plan.info['junction_in'] == 25 and plan.info['junction_out'] == 25


@pytest.mark.parametrize('path', ['lr', 'fast', 'generic'])
@pytest.mark.parametrize('name', ['hmm_scaled', 'hmm_cfg2_T200'])
def test_forward_backward_every_scan_kernel(ops, name, path, monkeypatch):
    """The three forward-backward kernels (aligned left-to-right loop in registers, register arc lists
    + shared memory, interpretive ELL) give the same posteriors / evidence on the same graph."""
    monkeypatch.setenv('BEER_B200_SCAN', path)
    g = load_golden(name)
    scale = float(g['scale'])
    llh = g['pdf_llh'] - g['pdf_llh'].max(axis=1, keepdims=True)
    plan, r = _fb_case(ops, llh, graph(g), scale=scale)
    if path == 'lr':
        assert plan.info['junctions'] == 1
    gamma = r['state_post'].double().cpu().numpy()
    assert np.abs(gamma - g['gamma']).max() <= 1e-5
    np.testing.assert_allclose(r['utt_exp_llh'].item() + scale * g['pdf_llh'].max(axis=1).sum(),
                               g['exp_llh'].sum(), rtol=1e-6)
    pdf_post = r['pdf_post'].double().cpu().numpy()
    np.testing.assert_allclose(pdf_post, scale * g['gamma'], atol=1e-5)


def test_forward_backward_dense_and_unreachable(ops):
    g = load_golden('dense_ergodic')
    _, r = _fb_case(ops, g['llhs'], graph(g))
    assert np.abs(r['state_post'].double().cpu().numpy() - g['gamma']).max() <= 1e-5
    _, r2 = _fb_case(ops, g['llhs2'], graph(g, 'g2_'))
    gam2 = r2['state_post'].double().cpu().numpy()
    assert np.isfinite(gam2).all()
    assert np.abs(gam2 - g['gamma2']).max() <= 1e-5
    assert (gam2[:, 6] == 0).all()                       # unreachable state: exactly zero, no NaN


@pytest.mark.parametrize('name', ['hmm_small', 'hmm_scaled', 'hmm_cfg2_T200'])
def test_viterbi_golden(ops, name):
    g = load_golden(name)
    scale = float(g['scale'])
    plan = ops.GraphPlan(*graph(g))
    T = len(g['X'])
    utt = torch.tensor([0, T], dtype=torch.int64, device=DEV)
    pdf_llh, _, _ = _emission(ops, g['X'], ng(g, 'post0_'))
    path = ops.hmm_viterbi(plan, pdf_llh, utt, scale=scale).cpu().numpy()
    np.testing.assert_array_equal(path, g['viterbi_path'])


def test_viterbi_ties_first_max(ops):
    g = load_golden('dense_ergodic')
    for sfx, gp in (('', 'g_'), ('2', 'g2_')):
        llh = g['llhs' + sfx]
        plan = ops.GraphPlan(*graph(g, gp))
        utt = torch.tensor([0, len(llh)], dtype=torch.int64, device=DEV)
        path = ops.hmm_viterbi(plan, dev(llh), utt).cpu().numpy()
        np.testing.assert_array_equal(path, g['path' + sfx])


@pytest.mark.parametrize('tc', [False, True])
@pytest.mark.parametrize('name', ['hmm_small', 'hmm_scaled', 'hmm_cfg2_T200'])
def test_estep_chain_normalset(ops, name, tc):
    """KA -> KB -> KC chained on the device against the golden E-step of the reference
    (tc: tcgen05 emission and statistics kernels; otherwise the SIMT ones)."""
    g = load_golden(name)
    scale = float(g['scale'])
    X = dev(g['X'])
    T, D = g['X'].shape
    pdf_llh, _, fref = _emission(ops, g['X'], ng(g, 'post0_'), tc=tc)
    plan = ops.GraphPlan(*graph(g))
    utt = torch.tensor([0, T], dtype=torch.int64, device=DEV)
    r = ops.hmm_forward_backward(plan, pdf_llh, fref, utt, scale=scale, want_state_post=True)
    assert np.abs(r['state_post'].double().cpu().numpy() - g['gamma']).max() <= 1e-5
    np.testing.assert_allclose(r['utt_exp_llh'].item(), g['exp_llh'].sum(), rtol=1e-6)
    K = g['gamma'].shape[1]
    acc = torch.zeros(K, 2 * D + 2, device=DEV, dtype=torch.float64)
    ops.accumulate_stats(X, acc, pdf_post=r['pdf_post'], tensor_cores=tc)
    got = acc.cpu().numpy()
    assert np.abs(got - g['acc_normal']).max() <= 2e-5 * np.abs(g['acc_normal']).max()
    # ELBO with datasize = 3T (objectives.py:176-184)
    prior, post = ng_dev(ng(g, 'prior_')), ng_dev(ng(g, 'post0_'))
    kl = ops.normalgamma_kl(prior, post).item()
    np.testing.assert_allclose(kl, g['kl'], rtol=1e-6)
    np.testing.assert_allclose(3.0 * r['utt_exp_llh'].item() - kl, g['elbo_datasize3T'], rtol=1e-6)


def test_mixtureset_two_groups(ops):
    """Two MixtureSet groups with different numbers of components (CSR component offsets),
    xi-free path over the decoding graph and over an alignment graph with repeated pdfs."""
    g = load_golden('phoneloop_mixtureset')
    C1, C2, K1 = int(g['C1']), int(g['C2']), int(g['K1'])
    K = len(g['g_map'])
    post = tuple(np.concatenate([g['g1_post0_' + k], g['g2_post0_' + k]]) for k in ('mean', 'scale', 'shape', 'rates'))
    comp_off = np.concatenate([np.arange(K1) * C1, K1 * C1 + np.arange(K - K1 + 1) * C2])
    M = comp_off[-1]
    D = g['X1'].shape[1]
    for X, gp, scale, tag in ((g['X1'], 'g_', 1.0, 'u1'), (g['X3'], 'ali_', 0.7, 'u3')):
        pdf_llh, comp, fref = _emission(ops, X, post, dir_post=(g['g1_dpost0'], g['g2_dpost0']),
                                        comp_off=comp_off, want_comp=True)
        if tag == 'u1':
            got = pdf_llh.double().cpu().numpy() + fref.double().cpu().numpy()[:, None]
            np.testing.assert_allclose(got, g['u1_pdf_llh'], atol=1e-4)
        plan = ops.GraphPlan(*graph(g, gp), n_pdfs=K)
        utt = torch.tensor([0, len(X)], dtype=torch.int64, device=DEV)
        r = ops.hmm_forward_backward(plan, pdf_llh, fref, utt, scale=scale, want_state_post=True,
                                     want_frame_llh=True)
        assert np.abs(r['state_post'].double().cpu().numpy() - g[tag + '_gamma']).max() <= 1e-5
        np.testing.assert_allclose(r['frame_exp_llh'].double().cpu().numpy(), g[tag + '_exp_llh'],
                                   rtol=1e-5, atol=1e-4)
        acc = torch.zeros(M, 2 * D + 2, device=DEV, dtype=torch.float64)
        ops.accumulate_stats(dev(X), acc, pdf_post=r['pdf_post'], pdf_llh=pdf_llh, comp_llh=comp,
                             comp_off=dev(comp_off, torch.int32))
        want = np.concatenate([g[tag + '_acc_g1'], g[tag + '_acc_g2']])
        assert np.abs(acc.cpu().numpy() - want).max() <= 2e-5 * np.abs(want).max()
        wst = ops.mixture_weight_stats(acc, D, comp_off=dev(comp_off, torch.int32)).cpu().numpy()
        want_w = np.concatenate([g[tag + '_acc_d1'].reshape(-1), g[tag + '_acc_d2'].reshape(-1)])
        np.testing.assert_allclose(wst, want_w, rtol=2e-5, atol=2e-5)


def test_gmm_cfg1(ops):
    """BASELINE configs[0]: 8-component diagonal Mixture on 2-D points."""
    g = load_golden('gmm_cfg1')
    X = g['X']
    post, dpost = ng(g, 'post0_'), g['dpost0']
    pdf_llh, comp, fref = _emission(ops, X, post, dir_post=(dpost[None],), comp_off=np.array([0, 8]),
                                    want_comp=True)
    exp_llh = pdf_llh.double().cpu().numpy()[:, 0] + fref.double().cpu().numpy()
    np.testing.assert_allclose(exp_llh, g['exp_llh'], rtol=1e-5, atol=1e-5)
    resps = torch.exp(comp - pdf_llh).double().cpu().numpy()
    assert np.abs(resps - g['resps']).max() <= 1e-5
    acc = torch.zeros(8, 2 * 2 + 2, device=DEV, dtype=torch.float64)
    ops.accumulate_stats(dev(X), acc, pdf_llh=pdf_llh, comp_llh=comp, comp_off=dev(np.array([0, 8]), torch.int32))
    assert np.abs(acc.cpu().numpy() - g['acc_normal']).max() <= 2e-5 * np.abs(g['acc_normal']).max()
    wst = ops.mixture_weight_stats(acc, 2, comp_off=dev(np.array([0, 8]), torch.int32)).cpu().numpy()
    np.testing.assert_allclose(wst, g['acc_dirichlet'], rtol=2e-5)


def test_ragged_batch_equals_single(ops):
    """A ragged batch (including an empty and a 1-frame utterance) gives, per utterance, what
    separate calls give (EvidenceLowerBoundInstance.__add__ semantics, objectives.py:78-90)."""
    g = load_golden('hmm_cfg2_T200')
    X = g['X']
    lens = [200, 0, 1, 37, 64, 200, 5]
    rng = np.random.default_rng(3)
    utts = [X[rng.integers(0, 200 - n + 1):][:n] if n else X[:0] for n in lens]
    Xc = np.concatenate(utts)
    off = np.concatenate([[0], np.cumsum(lens)])
    post = ng(g, 'post0_')
    pdf_llh, _, fref = _emission(ops, Xc, post)
    plan = ops.GraphPlan(*graph(g))
    r = ops.hmm_forward_backward(plan, pdf_llh, fref, dev(off, torch.int64), want_state_post=True)
    gam = r['state_post'].double().cpu().numpy()
    for i, n in enumerate(lens):
        if n == 0:
            assert r['utt_exp_llh'][i].item() == 0.0
            continue
        with np.errstate(all='ignore'):
            want = O.hmm_estep(utts[i].astype(np.float64), post, None, graph(g))
        if np.isnan(want['gamma']).any():
            # too short to reach a final state: the reference divides -inf by -inf (NaN,
            # graph.py:306-307); the kernel reports zero posteriors instead
            assert n < 4 and (gam[off[i]:off[i + 1]] == 0).all()
            continue
        assert np.abs(gam[off[i]:off[i + 1]] - want['gamma']).max() <= 1e-5
        np.testing.assert_allclose(r['utt_exp_llh'][i].item(), want['exp_llh'].sum(), rtol=1e-6)


@pytest.mark.parametrize('P,S,lrc', [(130, 4, '0'), (250, 4, '0'), (140, 3, '0'), (250, 4, '1'), (130, 4, '1'),
                                     (256, 4, '1'), (250, 4, 'w'), (130, 4, 'w'), (256, 4, 'w'), (140, 3, None)])
def test_forward_backward_many_units_block_kernel(ops, P, S, lrc, monkeypatch):
    """Loops with more than 128 units (BASELINE configs[2]: 250 units x 4 states) run W warps per utterance
    with one shared-memory exchange per reduction (lrc = '0': one unit per lane on eight warps, '1': two units per lane
    on four warps, 'w': ONE warp per utterance with eight units per lane, the kernel that also reduces unit counts;
    None: what the library picks); same posteriors / evidence as the generic kernel and as the fp64 oracle on a ragged batch."""
    if lrc is not None:
        monkeypatch.setenv('BEER_B200_SCAN_LRC', lrc)
    else:
        monkeypatch.delenv('BEER_B200_SCAN_LRC', raising=False)
    from beer_b200 import synthetic
    cg, _, _ = synthetic.phone_loop_graph(P, S)
    K = P * S
    rng = np.random.default_rng(P)
    lens = [37, 120, 1, 64]
    off = np.concatenate([[0], np.cumsum(lens)])
    llh = (rng.standard_normal((off[-1], K)) * 4).astype(np.float32)
    llh -= llh.max(axis=1, keepdims=True)
    gr = (cg.init_log_probs.numpy(), cg.final_log_probs.numpy(), cg.trans_log_probs.numpy(), np.arange(K))
    plan = ops.GraphPlan(*gr)
    assert plan.info['junctions'] == 1
    outs = {}
    for path in ('best', 'generic'):
        if path == 'generic':
            monkeypatch.setenv('BEER_B200_SCAN', 'generic')
        r = ops.hmm_forward_backward(plan, dev(llh), None, dev(off, torch.int64), scale=0.8, want_state_post=True,
                                     want_frame_llh=True, want_logz=True)
        outs[path] = {k: v.double().cpu().numpy() for k, v in r.items() if k != 'workspace' and v is not None}
    for k in ('state_post', 'pdf_post'):
        assert np.abs(outs['best'][k] - outs['generic'][k]).max() <= 1e-5, k
    for k in ('utt_exp_llh', 'utt_logz', 'frame_exp_llh'):
        np.testing.assert_allclose(outs['best'][k], outs['generic'][k], rtol=2e-6, atol=2e-4, err_msg=k)
    # fp64 oracle on the longest utterance
    g64 = tuple(np.asarray(x, dtype=np.float64) for x in gr[:3])
    u = 1
    with np.errstate(all='ignore'):
        gamma, _ = O.posteriors(0.8 * llh[off[u]:off[u + 1]].astype(np.float64), *g64)
    assert np.abs(outs['best']['state_post'][off[u]:off[u + 1]] - gamma).max() <= 1e-5
    # log2 posteriors straight from the loop kernel
    monkeypatch.delenv('BEER_B200_SCAN', raising=False)
    lp = torch.empty(int(off[-1]), K, device='cuda')
    ops.hmm_forward_backward(plan, dev(llh), None, dev(off, torch.int64), scale=0.8, want_pdf_post=False,
                             out_pdf_lpost=lp)
    assert np.abs(np.exp2(lp.double().cpu().numpy()) - outs['best']['pdf_post']).max() <= 2e-6


@pytest.mark.parametrize('P,SU', [(8, 4), (25, 4), (11, 3), (32, 4), (33, 4), (64, 4), (100, 3), (130, 3), (250, 4), (256, 4)])
def test_viterbi_loop_kernel_equals_generic_kernel_with_ties(ops, P, SU, monkeypatch):
    """The register-resident Viterbi kernels of aligned left-to-right loops (one unit per lane up to 32 units, several units
    per lane with a shared first maximum over the unit ends beyond: 250 x 4 is BASELINE configs[2]) against the generic one
    on llhs quantised to quarter units (thousands of exact ties): identical paths, i.e. the same first-max tie-breaking
    (graph.py:329-344), and identical to the oracle where fp32 and fp64 agree on the ties (uniform weights)."""
    gr, _, _ = O.phone_loop_graph(P, SU, self_loop=0.5)
    K = P * SU
    rng = np.random.default_rng(P * 10 + SU)
    lens = [1, 2, 33, 64, 257, 1000]
    off = np.concatenate([[0], np.cumsum(lens)])
    llh = np.round(rng.standard_normal((off[-1], K)) * 8) / 4 - 10.0
    plan = ops.GraphPlan(*gr)
    utt = torch.as_tensor(off, dtype=torch.int64, device=DEV)
    x = dev(llh)
    monkeypatch.delenv('BEER_B200_SCAN', raising=False)
    fast = ops.hmm_viterbi(plan, x, utt).cpu().numpy()
    monkeypatch.setenv('BEER_B200_SCAN', 'generic')
    slow = ops.hmm_viterbi(plan, x, utt).cpu().numpy()
    monkeypatch.delenv('BEER_B200_SCAN', raising=False)
    np.testing.assert_array_equal(fast, slow)
    # short utterances: no rounding drift yet, fp64 argmax sees the same ties
    for u in range(3):
        a, b = off[u], off[u + 1]
        want = O.best_path(llh[a:b].astype(np.float32).astype(np.float64), *[np.asarray(g, dtype=np.float64) for g in gr[:3]])
        np.testing.assert_array_equal(fast[a:b], want)
    if P > 32:
        # the same loop through the variant for learned unit weights (candidate ends ranked by the factored weights,
        # compared exactly with the dense ones) ...
        monkeypatch.setenv('BEER_B200_VIT_DENSE', '1')
        np.testing.assert_array_equal(ops.hmm_viterbi(plan, x, utt).cpu().numpy(), slow)
        monkeypatch.delenv('BEER_B200_VIT_DENSE', raising=False)
        # ... and with unit weights as PhoneLoop writes them: ln A[end, start u] = ln(1 - loop) + E[ln w_u] in fp32
        # (phoneloop.py:53-65); weights close to each other so that the rounding of the sums decides near-ties
        init, final, trans = [np.array(g, dtype=np.float32) for g in gr[:3]]
        logw = np.log(rng.dirichlet(np.full(P, 50.0))).astype(np.float32)
        for v in range(P):
            e = v * SU + SU - 1
            stay = np.float32(np.log1p(-np.exp(np.float32(trans[e, e]))))
            trans[e, np.arange(P) * SU] = stay + logw
        plan_w = ops.GraphPlan(init, final, trans, gr[3])
        fast_w = ops.hmm_viterbi(plan_w, x, utt).cpu().numpy()
        monkeypatch.setenv('BEER_B200_SCAN', 'generic')
        slow_w = ops.hmm_viterbi(plan_w, x, utt).cpu().numpy()
        monkeypatch.delenv('BEER_B200_SCAN', raising=False)
        np.testing.assert_array_equal(fast_w, slow_w)
        assert (fast_w != fast).any()


@pytest.mark.parametrize('N,M,D,C,scale', [(300, 200, 40, 1, 1.0), (129, 64, 20, 1, 0.5), (1000, 1000, 40, 8, 3.0),
                                           (77, 130, 13, 5, 1.0), (5, 3, 64, 1, 1.0)])
def test_emission_llh_bwd(N, M, D, C, scale):
    """KA backward (beer_emission_llh_bwd: fp16-split tcgen05, w as the tensor-memory operand) against the definition in
    fp64: grad_t = go_t (sum_j w_tj E[lambda mu]_j - x_t o sum_j w_tj E[lambda]_j), w = post x exp(comp - pdf);
    ragged tiles / chunks, every padded width of 2D, posteriors carrying a scale above and below one."""
    from beer_b200 import ops
    gen = torch.Generator().manual_seed(N + M)
    dev = torch.device('cuda', 0)
    X = (2.0 * torch.randn(N, D, generator=gen)).to(dev)
    ets = torch.randn(M, 2 * D + 2, generator=gen)
    ets[:, D:2 * D] = ets[:, D:2 * D].abs() * 3 + 0.1
    ets[:, :D] *= 50.0                                     # columns of very different magnitude
    ets = ets.to(dev)
    Kp = M // C
    post = torch.softmax(3.0 * torch.randn(N, Kp, generator=gen), dim=1) * scale
    post[post < 1e-4 * scale] = 0.0
    post = post.to(dev)
    go = torch.linspace(-1.0, 2.0, N, device=dev)
    comp = pdf = pdf_of = None
    w = post.double()
    if C > 1:
        comp = (4.0 * torch.randn(N, M, generator=gen)).to(dev)
        pdf = torch.logsumexp(comp.reshape(N, Kp, C).double(), dim=2).float()
        pdf_of = torch.arange(Kp, dtype=torch.int32, device=dev).repeat_interleave(C)
        w = w.repeat_interleave(C, dim=1) * torch.exp(comp.double() - pdf.double().repeat_interleave(C, dim=1))
    got = ops.emission_llh_bwd(X, ets, post, grad_out=go, comp_llh=comp, pdf_llh=pdf, pdf_of=pdf_of, scale=scale)
    e = ets.double()
    want = go.double()[:, None] * (w @ e[:, :D] - X.double() * (w @ e[:, D:2 * D]))
    assert float((got.double() - want).abs().max()) <= 5e-6 * float(want.abs().max())
    # no upstream gradient = ones
    got1 = ops.emission_llh_bwd(X, ets, post, comp_llh=comp, pdf_llh=pdf, pdf_of=pdf_of, scale=scale)
    want1 = w @ e[:, :D] - X.double() * (w @ e[:, D:2 * D])
    assert float((got1.double() - want1).abs().max()) <= 5e-6 * float(want1.abs().max())


def test_transition_posteriors_skip_empty_utterances():
    """A zero-length utterance inside the batch holds no transition: the block of the transition posteriors
    (graph.py:308-323 per utterance) is the one of the batch without it."""
    from beer_b200 import ops, synthetic
    dev = torch.device('cuda', 0)
    P, S, T1, T2 = 3, 3, 17, 9
    K = P * S
    graph, starts, ends = synthetic.phone_loop_graph(P, S)
    plan = graph.plan(n_pdfs=K)
    llh = torch.randn(T1 + T2, K, generator=torch.Generator().manual_seed(5)).to(dev)
    fref = torch.zeros(T1 + T2, device=dev)
    init = graph.init_log_probs.float().to(dev).contiguous()
    trans = graph.trans_log_probs.float().to(dev).contiguous()
    rows = torch.as_tensor(ends, dtype=torch.int32, device=dev)
    cols = torch.as_tensor(starts, dtype=torch.int32, device=dev)
    out = []
    for offs in ([0, T1, T1 + T2], [0, 0, T1, T1, T1 + T2, T1 + T2]):
        off = torch.tensor(offs, dtype=torch.int64, device=dev)
        r = ops.hmm_forward_backward(plan, llh, fref, off, want_state_post=True)
        out.append(ops.hmm_transition_posteriors(llh, r['state_post'], off, init, trans, rows=rows, cols=cols))
    assert out[0].shape == (T1 + T2 - 2, P, P)
    assert torch.equal(out[0], out[1])
