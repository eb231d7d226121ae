"""tcgen05 statistics kernel (KC, 3xTF32, operands transposed while staged) against an fp64 torch restatement of
resps.T @ stats (beer/models/normalset.py:121-123, mixtureset.py:100-112) and the SIMT kernel."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _exact(X, w):
    """fp64 [M, 2D+2] statistics of weights w [N, M]."""
    Xd, wd = X.double(), w.double()
    n = wd.sum(0)
    return torch.cat([wd.t() @ Xd, -0.5 * (wd.t() @ (Xd * Xd)), -0.5 * n[:, None], 0.5 * n[:, None]], dim=1)


# (M, C, D, N, use_post): partial frame tiles, Gaussian tiles that are not full, several tiles
CASES = [(100, 1, 40, 4000, True), (100, 1, 40, 77, True), (7, 1, 40, 1000, True), (130, 1, 40, 900, True),
         (100, 1, 40, 64 * 148 * 2 + 13, True), (100, 1, 40, 148 * 32 * 40 + 5, True), (64, 8, 40, 1500, True), (200, 4, 40, 1031, True),
         (96, 3, 20, 555, True), (8, 8, 40, 3000, False), (512, 1, 40, 2000, False), (48, 1, 64, 700, True),
         (40, 2, 80, 333, True),
         # mixtures on the bulk-staged path: several Gaussian tiles, a partial last tile, ragged last stage
         (1024, 8, 40, 3000, True), (2112, 8, 40, 2531, True), (160, 4, 40, 777, True), (512, 16, 40, 1000, True),
         (256, 32, 20, 413, True), (128, 1, 40, 500, True)]


@pytest.mark.parametrize('M,C,D,N,use_post', CASES)
def test_tc_statistics(M, C, D, N, use_post):
    from beer_b200 import ops
    ops.require_cuda()
    assert ops.accumulate_tc_supported(M, D)
    g = torch.Generator().manual_seed(M * 7 + C + N)
    Kp = M // C
    X = (torch.randn(N, D, generator=g) * 2 + 0.5).to(DEV)
    post = None
    if use_post:
        post = torch.softmax(torch.randn(N, Kp, generator=g) * 4, dim=1)
        post[post < 1e-4] = 0.              # exact zeros happen (unreachable states)
        post = post.to(DEV).contiguous()
    comp = pdf = comp_off = None
    if C > 1:
        comp = (torch.randn(N, M, generator=g) * 3 - 50).to(DEV).contiguous()
        pdf = torch.logsumexp(comp.reshape(N, Kp, C), dim=-1).contiguous()
        # uniform C: no CSR offsets (the bulk-staged mixture path); three cases keep the CSR form of the same layout
        comp_off = torch.arange(Kp + 1, dtype=torch.int32, device=DEV) * C if N in (1031, 555, 333) else None
        resp = torch.exp(comp.double().reshape(N, Kp, C) - pdf.double()[:, :, None])
        w = resp * (post.double()[:, :, None] if use_post else 1.)
        w = w.reshape(N, M)
    else:
        w = post.double() if use_post else torch.ones(N, M, dtype=torch.float64, device=DEV)
    want = _exact(X, w)
    got = torch.zeros(M, 2 * D + 2, device=DEV, dtype=torch.float64)
    ops.accumulate_stats(X, got, pdf_post=post, pdf_llh=pdf, comp_llh=comp, comp_off=comp_off, Kp=Kp,
                         tensor_cores=True)
    simt = torch.zeros_like(got)
    if 2 * D <= 128:        # the SIMT kernel's register tile stops at 2D = 128
        ops.accumulate_stats(X, simt, pdf_post=post, pdf_llh=pdf, comp_llh=comp, comp_off=comp_off, Kp=Kp,
                             tensor_cores=False)
    torch.cuda.synchronize()
    scale = want.abs().max(dim=1, keepdim=True).values.clamp(min=1e-3)
    err_tc = ((got - want).abs() / scale).max().item()
    err_simt = ((simt - want).abs() / scale).max().item() if 2 * D <= 128 else float('nan')
    # fp32 products of fp32 data, fp32 accumulation: both kernels sit at a few 1e-7 relative
    assert err_tc <= 3e-6, (err_tc, err_simt)
    # accumulation (+=) semantics: a second launch doubles the buffer
    ops.accumulate_stats(X, got, pdf_post=post, pdf_llh=pdf, comp_llh=comp, comp_off=comp_off, Kp=Kp,
                         tensor_cores=True)
    torch.cuda.synchronize()
    assert ((got - 2 * want).abs() / scale).max().item() <= 6e-6


def test_tc_ragged_components():
    """JointModelSet of mixtures with different numbers of Gaussians per pdf (comp_off CSR)."""
    from beer_b200 import ops
    ops.require_cuda()
    g = torch.Generator().manual_seed(3)
    counts = [10, 4, 4, 1, 7, 4, 130 - 30]
    comp_off_h = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    M, Kp, D, N = int(comp_off_h[-1]), len(counts), 40, 700
    X = torch.randn(N, D, generator=g).to(DEV)
    post = torch.softmax(torch.randn(N, Kp, generator=g), dim=1).to(DEV).contiguous()
    comp = (torch.randn(N, M, generator=g) * 2).to(DEV).contiguous()
    pdf = torch.stack([torch.logsumexp(comp[:, a:b], dim=1) for a, b in zip(comp_off_h[:-1], comp_off_h[1:])],
                      dim=1).contiguous()
    w = torch.cat([torch.exp(comp[:, a:b].double() - pdf[:, k:k + 1].double()) * post[:, k:k + 1].double()
                   for k, (a, b) in enumerate(zip(comp_off_h[:-1], comp_off_h[1:]))], dim=1)
    want = _exact(X, w)
    got = torch.zeros(M, 2 * D + 2, device=DEV, dtype=torch.float64)
    ops.accumulate_stats(X, got, pdf_post=post, pdf_llh=pdf, comp_llh=comp,
                         comp_off=torch.as_tensor(comp_off_h, device=DEV), Kp=Kp, tensor_cores=True)
    torch.cuda.synchronize()
    scale = want.abs().max(dim=1, keepdim=True).values.clamp(min=1e-3)
    assert ((got - want).abs() / scale).max().item() <= 3e-6


@pytest.mark.parametrize('M,D,N,scale', [(100, 40, 5013, 1.0), (100, 40, 148 * 32 * 9 + 1, 0.7), (7, 20, 77, 1.0),
                                         (128, 40, 4096, 1.0)])
def test_tc_statistics_of_a_path(M, D, N, scale):
    """Viterbi training: statistics of one-hot posteriors read from pdf ids (never materialised) == the dense kernel
    fed the one-hot matrix == the fp64 restatement."""
    from beer_b200 import ops
    ops.require_cuda()
    g = torch.Generator().manual_seed(M + N)
    X = (torch.randn(N, D, generator=g) * 2 + 0.5).to(DEV)
    ids_h = torch.randint(0, M, (N,), generator=g, dtype=torch.int32)
    buf = torch.zeros(N + 4, dtype=torch.int32, device=DEV)
    buf[:N] = ids_h.to(DEV)
    onehot = torch.zeros(N, M, device=DEV)
    onehot[torch.arange(N, device=DEV), ids_h.to(DEV).long()] = scale
    want = _exact(X, onehot)
    got = torch.zeros(M, 2 * D + 2, device=DEV, dtype=torch.float64)
    ops.accumulate_stats_path(X, got, buf[:N], scale=scale)
    torch.cuda.synchronize()
    sc = want.abs().max(dim=1, keepdim=True).values.clamp(min=1e-3)
    assert ((got - want).abs() / sc).max().item() <= 3e-6
    ops.accumulate_stats_path(X, got, buf[:N], scale=scale)        # += semantics
    torch.cuda.synchronize()
    assert ((got - 2 * want).abs() / sc).max().item() <= 6e-6
