"""GPU parity tests of the mixture path that keeps the per-Gaussian llhs on chip (csrc/mix16.cu: fp16 hi / lo feature
images, KA16 emission kernel, KCF statistics kernel with the responsibilities recomputed in tensor memory) against
fp64 restatements of MixtureSet.expected_log_likelihood / accumulate (beer/models/mixtureset.py:85-112,
normalset.py:117-123)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda'
LN2 = float(np.log(2.0))


def _model(M, D, C, seed, spread=1.0):
    g = torch.Generator().manual_seed(seed)
    mean = spread * torch.randn(M, D, generator=g)
    scale = torch.rand(M, generator=g) * 3 + 0.5
    shape = torch.rand(M, generator=g) * 4 + 1.0
    rates = torch.rand(M, D, generator=g) * 2 + 0.3
    conc = torch.rand(M // C, C, generator=g) * 3 + 0.2
    return [t.to(DEV).contiguous() for t in (mean, scale, shape, rates)], conc.to(DEV)


def _reference(X, post, conc, C):
    """fp64: comp llh [N, M] (with E[ln pi]), pdf llh [N, Kp]."""
    from beer_b200 import ops
    mean, scale, shape, rates = [t.double() for t in post]
    D = mean.shape[1]
    a, k = shape[:, None], scale[:, None]
    lam = a / rates
    ets = torch.cat([lam * mean, lam, D / k + (lam * mean ** 2).sum(1, keepdim=True),
                     (torch.digamma(a) - rates.log()).sum(1, keepdim=True)], 1)
    Xd = X.double()
    stats = torch.cat([Xd, -0.5 * Xd ** 2, -0.5 * torch.ones(len(Xd), 1, device=DEV, dtype=torch.float64),
                       0.5 * torch.ones(len(Xd), 1, device=DEV, dtype=torch.float64)], 1)
    comp = stats @ ets.T - 0.5 * D * np.log(2 * np.pi)
    logw = torch.digamma(conc.double()) - torch.digamma(conc.double().sum(1, keepdim=True))
    comp = comp + logw.reshape(1, -1)
    pdf = torch.logsumexp(comp.reshape(len(Xd), -1, C), dim=2)
    return stats, comp, pdf


@pytest.mark.parametrize('M,D,C,N', [(8000, 40, 8, 700), (256, 40, 8, 1000), (160, 20, 8, 333), (96, 40, 4, 130),
                                     (320, 40, 16, 64), (100, 40, 1, 1000), (256, 40, 1, 333), (12, 20, 1, 130)])
def test_emission_and_statistics_match_fp64(M, D, C, N):
    from beer_b200 import ops
    Kp = M // C
    post, conc = _model(M, D, C, seed=M + N)
    g = torch.Generator().manual_seed(N)
    X = (2.0 * torch.randn(N, D, generator=g)).to(DEV)
    X[:, 0] *= 10.0                     # dimensions of different ranges: the per-dimension scales matter
    X[:, 1] *= 0.01
    logw = ops.dirichlet_expected_logw(conc).reshape(-1).contiguous()      # (C = 1: psi(a) - psi(a) = 0)
    W, bias, ref = ops.emission_prepare(*post, logw=logw)
    mx = ops.Mix16(M, D, C, DEV)
    images = mx.build_images(X)
    # the images decode to the scaled statistics
    KP = mx.KP
    alpha = images['alpha'].double().cpu().numpy()
    img1 = images['img1'].double().cpu().numpy().reshape(-1, 2, 8, KP // 8, 8, 8)     # tile, hi/lo, f8, k8, f, k
    dec = (img1[:, 0] + img1[:, 1]).transpose(0, 1, 3, 2, 4).reshape(-1, KP)[:N, :2 * D] / alpha
    Xd = X.double().cpu().numpy()
    want = np.concatenate([Xd, -0.5 * Xd ** 2], 1)
    assert np.abs(dec - want).max() <= 2.0 ** -20 * np.abs(want).max(axis=0).max()
    img2 = images['img2'].double().cpu().numpy().reshape(-1, 2, KP // 8, 8, 8, 8)     # tile, hi/lo, k8, f8, k, f
    dec2 = (img2[:, 0] + img2[:, 1]).transpose(0, 2, 4, 1, 3).reshape(-1, KP)[:N, :2 * D] / alpha
    np.testing.assert_array_equal(dec2, dec)

    mx.pack(W, bias, images['alpha'])
    llh2 = mx.emission(images)
    fref = mx.frame_ref(X, ref)
    stats, comp, pdf = _reference(X, post, conc, C)
    got = llh2.double() * LN2 + fref.double()[:, None]
    err = (got - pdf).abs().max().item()
    assert err <= 3e-6 * pdf.abs().max().item() + 1e-4, (err, pdf.abs().max().item())

    # statistics with random pdf posteriors (some exactly zero), responsibilities recomputed on chip
    pp = torch.rand(N, Kp, generator=g).to(DEV)
    pp = torch.where(pp < 0.6, torch.zeros_like(pp), pp).contiguous()
    acc = torch.zeros(M, 2 * D + 2, device=DEV, dtype=torch.float64)
    lpp = mx.log2_posteriors(pp) if C > 1 else pp      # single-Gaussian pdfs: the posteriors themselves
    mx.accumulate(images, lpp, llh2, acc)
    resp = (comp.reshape(N, Kp, C) - pdf[:, :, None]).exp()
    w = (resp * pp.double()[:, :, None]).reshape(N, M)
    want_acc = w.T @ stats
    tol = 3e-5 * want_acc.abs().max().item()
    assert (acc - want_acc).abs().max().item() <= tol, ((acc - want_acc).abs().max().item(), tol)
    # the z of the statistics kernel is the z the emission kernel normalised: the responsibilities of a pdf sum to one
    cnt = 2.0 * acc[:, 2 * D + 1].reshape(Kp, C).sum(1)
    want_cnt = pp.double().sum(0)
    assert (cnt - want_cnt).abs().max().item() <= 5e-6 * want_cnt.abs().max().item() + 1e-9
    # accumulation semantics: a second call adds
    mx.accumulate(images, lpp, llh2, acc)
    assert (acc - 2 * want_acc).abs().max().item() <= 2 * tol
    if C > 1:
        # relative form: ONE array log2 posterior - llh2 (what the forward-backward writes with lpost_relative);
        # its rounding is half an ulp of |llh2| in the exponent, zero-mean over the frames
        acc_r = torch.zeros_like(acc)
        mx.accumulate(images, (lpp - llh2).contiguous(), None, acc_r, relative=True)
        assert (acc_r - want_acc).abs().max().item() <= tol, ((acc_r - want_acc).abs().max().item(), tol)
        cnt = 2.0 * acc_r[:, 2 * D + 1].reshape(Kp, C).sum(1)
        assert (cnt - want_cnt).abs().max().item() <= 2e-5 * want_cnt.abs().max().item() + 1e-9
        with pytest.raises(ValueError):
            mx.accumulate(images, lpp, llh2, acc_r, relative=True)


def test_forward_backward_accepts_log2_llhs():
    """beer_hmm_forward_backward_ex(llh_log2=1) on llh / ln 2 == the call on llh (same posteriors, same ELBO terms)."""
    from beer_b200 import ops, synthetic
    P, S, T = 250, 4, 90
    K = P * S
    graph, _, _ = synthetic.phone_loop_graph(P, S)
    plan = ops.GraphPlan(graph.init_log_probs.numpy(), graph.final_log_probs.numpy(),
                         graph.trans_log_probs.numpy(), graph.pdf_id_mapping, n_pdfs=K)
    g = torch.Generator().manual_seed(1)
    llh = (5.0 * torch.randn(2 * T, K, generator=g)).to(DEV)
    fref = torch.randn(2 * T, generator=g).to(DEV)
    off = torch.tensor([0, T, 2 * T], device=DEV)
    a = ops.hmm_forward_backward(plan, llh, fref, off, want_frame_llh=True)
    b = ops.hmm_forward_backward(plan, (llh / LN2).contiguous(), fref, off, want_frame_llh=True, llh_log2=True)
    assert (a['pdf_post'] - b['pdf_post']).abs().max().item() <= 2e-5
    # log2 posteriors straight from the loop kernel
    assert plan.writes_log2_posteriors
    lp = torch.empty(2 * T, K, device=DEV)
    ops.hmm_forward_backward(plan, llh, fref, off, want_pdf_post=False, out_pdf_lpost=lp, scale=0.5)
    c = ops.hmm_forward_backward(plan, llh, fref, off, scale=0.5)
    assert (torch.exp2(lp) - c['pdf_post']).abs().max().item() <= 2e-6
    small = torch.zeros(40, 12, device=DEV)            # 3 units x 4 states: the one-warp loop kernel
    g3, _, _ = synthetic.phone_loop_graph(3, 4)
    p3 = ops.GraphPlan(g3.init_log_probs.numpy(), g3.final_log_probs.numpy(), g3.trans_log_probs.numpy(),
                       g3.pdf_id_mapping, n_pdfs=12)
    small.copy_(3.0 * torch.randn(40, 12, generator=g))
    lp3 = torch.empty(40, 12, device=DEV)
    o3 = torch.tensor([0, 40], device=DEV)
    ops.hmm_forward_backward(p3, small, None, o3, want_pdf_post=False, out_pdf_lpost=lp3)
    c3 = ops.hmm_forward_backward(p3, small, None, o3)
    assert (torch.exp2(lp3) - c3['pdf_post']).abs().max().item() <= 2e-6
    np.testing.assert_allclose(a['utt_exp_llh'].cpu().numpy(), b['utt_exp_llh'].cpu().numpy(), rtol=2e-6)
    # relative form (BEER_FB_LPOST_RELATIVE): log2 posterior - log2 llh, from the eight-warp and the one-warp kernel
    llh2 = (llh / LN2).contiguous()
    lp_abs, lp_rel = torch.empty_like(lp), torch.empty_like(lp)
    ops.hmm_forward_backward(plan, llh2, fref, off, want_pdf_post=False, out_pdf_lpost=lp_abs, llh_log2=True)
    ops.hmm_forward_backward(plan, llh2, fref, off, want_pdf_post=False, out_pdf_lpost=lp_rel, llh_log2=True,
                             lpost_relative=True)
    fin = torch.isfinite(lp_abs)
    assert ((lp_rel - (lp_abs - llh2))[fin]).abs().max().item() <= 2e-5
    assert torch.equal(torch.isfinite(lp_rel), fin)
    # ... and under an acoustic scale (the eight-warp kernel keeps the llhs scaled: it divides the scale out again)
    ops.hmm_forward_backward(plan, llh2, fref, off, want_pdf_post=False, out_pdf_lpost=lp_abs, llh_log2=True, scale=0.5)
    ops.hmm_forward_backward(plan, llh2, fref, off, want_pdf_post=False, out_pdf_lpost=lp_rel, llh_log2=True, scale=0.5,
                             lpost_relative=True)
    fin = torch.isfinite(lp_abs)
    assert ((lp_rel - (lp_abs - llh2))[fin]).abs().max().item() <= 2e-5
    lp3r = torch.empty_like(lp3)
    ops.hmm_forward_backward(p3, small, None, o3, want_pdf_post=False, out_pdf_lpost=lp3r, lpost_relative=True, scale=2.0)
    c3s = ops.hmm_forward_backward(p3, small, None, o3, scale=2.0)
    assert (torch.exp2(lp3r + small / LN2) - c3s['pdf_post']).abs().max().item() <= 1e-5


def test_statistics_skip_inactive_blocks_exactly():
    """Activity map (beer_hmm_forward_backward_blocks -> beer_mix16_accumulate_blocks): the forward-backward marks the
    (tile of 64 frames, pdfs of one Gaussian tile) pairs in which a posterior is large enough to be non-zero in the
    statistics kernel's fp16 operands; the kernel skips the others: exact zeros.  What remains are the same products,
    grouped differently into the fp32 partial sums (drain every four worked tiles): equal to summation order."""
    from beer_b200 import ops, synthetic
    P, S, C, D, T, U = 40, 4, 8, 40, 150, 5
    K, M = P * S, P * S * C
    graph, _, _ = synthetic.phone_loop_graph(P, S)
    plan = ops.GraphPlan(graph.init_log_probs.numpy(), graph.final_log_probs.numpy(), graph.trans_log_probs.numpy(),
                         graph.pdf_id_mapping, n_pdfs=K)
    assert plan.marks_active_blocks()
    means = 2.0 * torch.randn(K, D, generator=torch.Generator().manual_seed(0))
    X = synthetic.sample_utterances(graph, means, U, T, seed=1, device=DEV)
    # a model that fits: component means at the state means (+ noise), so that the posteriors are peaked
    g = torch.Generator().manual_seed(3)
    mean = (means.repeat_interleave(C, 0) + 0.3 * torch.randn(M, D, generator=g)).to(DEV)
    post = [mean, torch.full((M,), 2.0, device=DEV), torch.full((M,), 3.0, device=DEV), torch.full((M, D), 3.0, device=DEV)]
    conc = torch.ones(K, C, device=DEV)
    logw = ops.dirichlet_expected_logw(conc).reshape(-1).contiguous()
    W, bias, ref = ops.emission_prepare(*post, logw=logw)
    mx = ops.Mix16(M, D, C, DEV)
    images = mx.build_images(X)
    mx.pack(W, bias, images['alpha'])
    llh2 = mx.emission(images)
    fref = mx.frame_ref(X, ref)
    off = torch.arange(U + 1, device=DEV) * T
    N = U * T
    nb = (K + mx.pdfs_per_block - 1) // mx.pdfs_per_block
    blocks = torch.zeros((N + 63) // 64, nb, dtype=torch.uint8, device=DEV)
    lrel, labs = torch.empty(N, K, device=DEV), torch.empty(N, K, device=DEV)
    ops.hmm_forward_backward(plan, llh2, fref, off, want_pdf_post=False, out_pdf_lpost=labs, llh_log2=True)
    ops.hmm_forward_backward(plan, llh2, fref, off, want_pdf_post=False, out_pdf_lpost=lrel, llh_log2=True,
                             lpost_relative=True, block_active=blocks, pdfs_per_block=mx.pdfs_per_block)
    # the marks are the blocks whose largest log2 posterior reaches -(25 + 14 + 2)
    pad = (-N) % 64
    lp = torch.nn.functional.pad(labs, (0, nb * mx.pdfs_per_block - K, 0, pad), value=float('-inf'))
    top = lp.reshape(-1, 64, nb, mx.pdfs_per_block).amax(dim=(1, 3))
    assert ((top >= -40.9) <= (blocks > 0)).all() and ((blocks > 0) <= (top >= -41.1)).all()
    frac = blocks.float().mean().item()
    assert 0.0 < frac < 0.5, frac          # peaked posteriors: most pairs carry no weight
    dense = torch.zeros(M, 2 * D + 2, device=DEV, dtype=torch.float64)
    sparse = torch.zeros_like(dense)
    mx.accumulate(images, lrel, None, dense, relative=True)
    mx.accumulate(images, lrel, None, sparse, relative=True, block_active=blocks)
    mom_d, mom_s = dense[:, :2 * D], sparse[:, :2 * D]                   # first and second moments
    assert (mom_d - mom_s).abs().max().item() <= 1e-6 * mom_d.abs().max().item()
    assert torch.equal(mom_d == 0, mom_s == 0)
    assert (dense[:, 2 * D:] - sparse[:, 2 * D:]).abs().max().item() <= 1e-9      # counts: the fp32 sum of weights < 2^-41
    # ... and every block marked: the dense result again, bit for bit
    full = torch.zeros_like(dense)
    mx.accumulate(images, lrel, None, full, relative=True, block_active=torch.ones_like(blocks))
    assert torch.equal(full, dense)
