"""tcgen05 emission kernel (3xTF32) against the fp32 SIMT kernel and the fp64 golden llhs."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _params(M, D, seed):
    g = torch.Generator().manual_seed(seed)
    mean = torch.randn(M, D, generator=g) * 2
    scale = torch.rand(M, generator=g) * 3 + .5
    shape = torch.rand(M, generator=g) * 4 + 1.
    rates = torch.rand(M, D, generator=g) * 2 + .3
    return tuple(t.to(DEV).contiguous() for t in (mean, scale, shape, rates))


@pytest.mark.parametrize('M,C,D,N', [(100, 1, 40, 1000), (100, 1, 40, 77), (16, 1, 40, 300), (128, 2, 40, 515),
                                     (200, 2, 40, 400), (1024, 8, 40, 260), (96, 4, 20, 333), (48, 16, 40, 129),
                                     # streamed chunks with the staged per-Gaussian tile; partial last chunk
                                     (1000, 8, 40, 300), (520, 4, 40, 200), (2048, 16, 40, 1000), (640, 8, 20, 131)])
def test_tc_matches_simt(M, C, D, N):
    from beer_b200 import ops
    ops.require_cuda()
    assert ops.emission_tc_supported(M, D, C)
    post = _params(M, D, seed=M + C)
    Kp = M // C
    logw = None
    if C > 1:
        conc = torch.rand(Kp, C, generator=torch.Generator().manual_seed(1)).to(DEV) * 3 + .2
        logw = ops.dirichlet_expected_logw(conc).reshape(-1).contiguous()
    W, bias, ref = ops.emission_prepare(*post, logw=logw)
    X = (torch.randn(N, D, generator=torch.Generator().manual_seed(5)) * 2.5).to(DEV)
    comp_off = None if C == 1 else torch.arange(Kp + 1, dtype=torch.int32, device=DEV) * C
    want, want_comp, want_ref = ops.emission_llh(X, W, bias, ref, comp_off=comp_off, Kp=Kp, want_comp=C > 1)
    img = ops.emission_tc_pack(W, bias, C)
    got, got_comp, got_ref = ops.emission_llh_tc(X, img, ref, M, C, want_comp=C > 1)
    torch.cuda.synchronize()
    # fp64 restatement of the same contraction from the fp32 weights
    S = torch.cat([X, -0.5 * X * X], dim=1).double()
    exact = S @ W.double().t() + bias.double()[None]
    if C > 1:
        exact_pdf = torch.logsumexp(exact.reshape(N, Kp, C), dim=-1)
    else:
        exact_pdf = exact
    err_tc = (got.double() - exact_pdf).abs().max().item()
    err_simt = (want.double() - exact_pdf).abs().max().item()
    scale = exact.abs().max().item()
    # the fp32 accumulation in TMEM truncates (30 accumulations per output): measured 7.6e-7 * scale,
    # against 2.2e-7 * scale for the round-to-nearest SIMT FMAs
    assert err_tc <= 1.5e-6 * scale + 1e-5, (err_tc, err_simt, scale)
    np.testing.assert_allclose(got_ref.cpu().numpy(), want_ref.cpu().numpy(), rtol=1e-5, atol=1e-4)
    if C > 1:
        assert (got_comp.double() - exact).abs().max().item() <= 1.5e-6 * scale + 1e-5


def test_tc_golden_cfg2():
    from beer_b200 import ops
    g = load_golden('hmm_cfg2_T200')
    post = tuple(torch.as_tensor(g['post0_' + k], dtype=torch.float32).reshape(
        (100, 40) if k in ('mean', 'rates') else (100,)).to(DEV).contiguous() for k in ('mean', 'scale', 'shape', 'rates'))
    W, bias, ref = ops.emission_prepare(*post)
    X = torch.as_tensor(g['X']).to(DEV)
    img = ops.emission_tc_pack(W, bias, 1)
    pdf_llh, _, fref = ops.emission_llh_tc(X, img, ref, 100, 1)
    d_got = pdf_llh.double().cpu().numpy()
    d_ref = g['pdf_llh'] - g['pdf_llh'].max(axis=1, keepdims=True)
    d_got = d_got - d_got[np.arange(len(d_got)), g['pdf_llh'].argmax(axis=1)][:, None]
    near = d_ref > -30
    assert np.abs(d_got - d_ref)[near].max() < 6e-5
    got = pdf_llh.double().cpu().numpy() + fref.double().cpu().numpy()[:, None]
    np.testing.assert_allclose(got, g['pdf_llh'], rtol=0, atol=2e-4)
