"""Speech front-end with the call surface of beer/features.py: `fbank`, `create_fbank`, `add_deltas`,
`hz2mel`, `mel2hz`.  The filterbank matrix and the window are host-side constants (built once, as the
reference's `lru_cache`d `create_fbank` does); framing, pre-emphasis, windowing, FFT, filterbank and log
run in the sm_100a kernel `beer_fbank`, the regression filter in `beer_add_deltas`.  Results are CUDA
tensors (the reference returns numpy arrays)."""
from functools import lru_cache

import numpy as np
import torch

from . import ops

__all__ = ['hz2mel', 'mel2hz', 'create_fbank', 'fbank', 'short_term_mspec', 'log_mel_spectrum', 'add_deltas']


def hz2mel(freq_hz):
    return 1127 * np.log(1 + freq_hz / 700.0)


def mel2hz(mel):
    return 700.0 * (np.exp(mel / 1127.0) - 1)


def _slope(lo, hi, bins, rising):
    """One side of a triangular filter sampled on FFT bins: a linspace between the values at the first
    and the last bin inside [lo, hi] (features.py:30-43)."""
    sel = (bins >= lo) & (bins <= hi)
    vals = np.zeros(len(bins))
    if sel.any():
        f = bins[sel]
        with np.errstate(divide='ignore', invalid='ignore'):
            ends = ((f[0] - lo) / (hi - lo), (f[-1] - lo) / (hi - lo)) if rising else \
                   ((hi - f[0]) / (hi - lo), (hi - f[-1]) / (hi - lo))
        vals[sel] = np.linspace(ends[0], ends[1], len(f))
    return sel, vals


@lru_cache(maxsize=8)
def create_fbank(nfilters, fft_len=512, srate=16000, lowfreq=0, highfreq=None):
    """[nfilters, fft_len / 2] mel filterbank, filter centres aligned to FFT bins (features.py:47-79)."""
    highfreq = highfreq or srate / 2
    centers = np.floor(fft_len * mel2hz(np.linspace(hz2mel(lowfreq), hz2mel(highfreq), nfilters + 2)) / srate)
    bins = np.arange(0, fft_len // 2)
    filters = np.zeros((nfilters, fft_len // 2))
    for i in range(1, nfilters + 1):
        up, v_up = _slope(centers[i - 1], centers[i], bins, True)
        down, v_down = _slope(centers[i], centers[i + 1], bins, False)
        filters[i - 1, up] = v_up[up]
        filters[i - 1, down] = v_down[down]      # the falling side overwrites the shared centre bin
    return filters


@lru_cache(maxsize=8)
def _constants(flen_samp, fft_len, nfilters, srate, lowfreq, hifreq, device):
    window = torch.as_tensor(np.hamming(flen_samp), dtype=torch.float32, device=device)
    filt = create_fbank(nfilters, fft_len, srate=srate, lowfreq=lowfreq, highfreq=hifreq)
    filt_t = torch.as_tensor(np.ascontiguousarray(filt.T), dtype=torch.float32, device=device)
    return window, filt_t


def fbank(signal, flen=0.025, frate=0.01, hifreq=8000, lowfreq=20, nfilters=26, preemph=0.97, srate=16000,
          device='cuda'):
    """FBANK features [n_frames, nfilters] of a raw signal (features.py:145-204)."""
    frate_samp, flen_samp = int(srate * frate), int(srate * flen)
    fft_len = int(2 ** np.floor(np.log2(flen_samp) + 1))
    if isinstance(signal, torch.Tensor):
        sig = signal.to(device=device, dtype=torch.float32).contiguous()
    else:
        sig = torch.as_tensor(np.asarray(signal, dtype=np.float32), device=device)
    # the reference builds the filterbank with create_fbank's default srate (16000) whatever `srate` is
    # (features.py:198): kept, so that results are identical for every sampling rate
    window, filt_t = _constants(flen_samp, fft_len, nfilters, 16000, lowfreq, hifreq, str(sig.device))
    return ops.fbank(sig, window, filt_t, frate_samp, preemph, fft_len)


def _signal_and_dc(signal, device):
    if isinstance(signal, torch.Tensor):
        sig = signal.to(device=device, dtype=torch.float32).contiguous()
        dc = float(signal.double().mean().item()) if signal.numel() else 0.0
    else:
        arr = np.asarray(signal)
        dc = float(arr.mean()) if arr.size else 0.0          # as the reference: mean of the samples in float64
        sig = torch.as_tensor(arr.astype(np.float32), device=device)
    return sig, dc


def short_term_mspec(signal, flen=0.025, frate=0.01, preemph=0.97, srate=16000, window=np.hamming, device='cuda'):
    """Short-term magnitude spectrum (features.py:102-143): DC removed, pre-emphasis inside every frame, window,
    |rFFT| without the Nyquist bin.  Returns (mspec [n_frames, fft_len / 2] on `device`, fft_len)."""
    frate_samp, flen_samp = int(srate * frate), int(srate * flen)
    fft_len = int(2 ** np.floor(np.log2(flen_samp) + 1))
    sig, dc = _signal_and_dc(signal, device)
    win = torch.as_tensor(window(flen_samp), dtype=torch.float32, device=sig.device)
    return ops.short_term_mspec(sig, win, frate_samp, preemph, fft_len, dc), fft_len


def log_mel_spectrum(signal, nfilters=40, flen=0.025, frate=0.01, preemph=0.97, srate=16000, lowfreq=20, hifreq=8000,
                     device='cuda'):
    """The features `beer features extract` writes for an fbank configuration (extract.py:107-127):
    log(1e-6 + short_term_mspec(signal) @ create_fbank(nfilters, fft_len, lowfreq, highfreq).T), one fused kernel."""
    frate_samp, flen_samp = int(srate * frate), int(srate * flen)
    fft_len = int(2 ** np.floor(np.log2(flen_samp) + 1))
    sig, dc = _signal_and_dc(signal, device)
    window, filt_t = _constants(flen_samp, fft_len, nfilters, 16000, lowfreq, hifreq, str(sig.device))
    return ops.short_term_mspec(sig, window, frate_samp, preemph, fft_len, dc, filters_t=filt_t, log_offset=1e-6)


def add_deltas(fea, winlens=(2, 2)):
    """Append deltas, delta-deltas, ... (features.py:82-100)."""
    fea = fea.to(torch.float32).contiguous()
    feats = [fea]
    for wlen in winlens:
        fea = ops.add_deltas(fea, wlen)
        feats.append(fea)
    return torch.cat(feats, dim=1)
