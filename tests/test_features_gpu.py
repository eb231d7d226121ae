"""fbank front-end kernel (beer_fbank, beer_add_deltas) against the golden of the live reference
(beer/features.py:145-204, 82-100).  fp32 FFT vs the reference's float64: log-energies within 2e-4."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('nfilters', [40, 26])
def test_fbank_matches_reference(nfilters):
    from beer_b200 import features as F
    g = load_golden('fbank')
    got = F.fbank(g['signal'], nfilters=nfilters)
    want = g[f'fbank{nfilters}']
    assert got.shape == want.shape
    np.testing.assert_allclose(got.double().cpu().numpy(), want, rtol=0, atol=2e-4)


def test_fbank_edge_cases():
    from beer_b200 import features as F
    g = load_golden('fbank')
    sig = g['signal']
    assert F.fbank(sig[:399], nfilters=40).shape == (0, 40)        # shorter than one frame
    one = F.fbank(sig[:400], nfilters=40)
    np.testing.assert_allclose(one.double().cpu().numpy(), g['fbank40'][:1], atol=2e-4)
    # a CUDA float tensor is accepted as is
    t = torch.as_tensor(sig.astype(np.float32), device='cuda')
    np.testing.assert_allclose(F.fbank(t, nfilters=40).cpu().numpy(), F.fbank(sig, nfilters=40).cpu().numpy())


def test_add_deltas_matches_reference():
    from beer_b200 import features as F
    g = load_golden('fbank')
    fea = torch.as_tensor(g['fbank40'], dtype=torch.float32, device='cuda')
    got = F.add_deltas(fea).double().cpu().numpy()
    np.testing.assert_allclose(got, g['deltas40'], rtol=0, atol=5e-6)
    got1 = F.add_deltas(fea[:1], winlens=(3,)).cpu().numpy()       # a single frame: zero deltas
    assert got1.shape == (1, 80) and np.all(got1[:, 40:] == 0)


def test_cli_front_end_matches_reference():
    """The features `beer features extract` writes (features.py:102-143 short_term_mspec, extract.py:107-127):
    DC removal, per-frame pre-emphasis, |rFFT|, mel filterbank, log(1e-6 + .)."""
    from beer_b200 import features as F
    g = load_golden('fbank')
    mspec, fft_len = F.short_term_mspec(g['signal'])
    assert fft_len == 512 and mspec.shape == g['mspec'].shape
    want = g['mspec']
    assert np.abs(mspec.double().cpu().numpy() - want).max() <= 2e-6 * np.abs(want).max()
    got = F.log_mel_spectrum(g['signal'], nfilters=40)
    np.testing.assert_allclose(got.double().cpu().numpy(), g['cli_logmel40'], rtol=0, atol=2e-4)
    assert F.short_term_mspec(g['signal'][:399])[0].shape == (0, 256)
