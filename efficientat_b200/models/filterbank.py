"""Kaldi-style triangular mel filterbank and its banded (start, length, taps) form.

Host-side (CPU, tiny) table construction for the fused mel kernel.  Same arithmetic, in the same
fp32 op order, as torchaudio.compliance.kaldi.get_mel_banks(vtln_warp_factor=1.0), which the
reference calls on every forward (models/preprocess.py:52-55); here the table is cached per
(fmin, fmax) and kept on the device.
"""
import math

import torch


def kaldi_mel_banks(n_mels, n_fft, sr, fmin, fmax):
    """-> [n_mels, n_fft//2 + 1] fp32 (last column is the zero pad of preprocess.py:54)."""
    assert n_mels > 3 and n_fft % 2 == 0
    nyquist = 0.5 * sr
    if fmax <= 0.0:
        fmax += nyquist
    assert 0.0 <= fmin < nyquist and 0.0 < fmax <= nyquist and fmin < fmax, \
        f"Bad values in options: low-freq {fmin} and high-freq {fmax} vs. nyquist {nyquist}"
    bin_width = sr / n_fft
    mel_lo = 1127.0 * math.log(1.0 + fmin / 700.0)
    mel_hi = 1127.0 * math.log(1.0 + fmax / 700.0)
    delta = (mel_hi - mel_lo) / (n_mels + 1)
    idx = torch.arange(n_mels).unsqueeze(1)
    left = mel_lo + idx * delta
    center = mel_lo + (idx + 1.0) * delta
    right = mel_lo + (idx + 2.0) * delta
    mel = (1127.0 * (1.0 + bin_width * torch.arange(n_fft / 2) / 700.0).log()).unsqueeze(0)
    up = (mel - left) / (center - left)
    down = (right - mel) / (right - center)
    fb = torch.max(torch.zeros(1), torch.min(up, down))
    return torch.nn.functional.pad(fb, (0, 1), mode="constant", value=0)


def to_bands(fb):
    """Dense [n_mels, n_bins] -> (start[int32 n_mels], length[int32 n_mels], taps[max_len, n_mels])."""
    n_mels = fb.shape[0]
    nz = fb != 0
    start = torch.zeros(n_mels, dtype=torch.int32)
    length = torch.zeros(n_mels, dtype=torch.int32)
    for m in range(n_mels):
        cols = torch.nonzero(nz[m]).flatten()
        if cols.numel():
            start[m] = int(cols[0])
            length[m] = int(cols[-1]) - int(cols[0]) + 1
    max_len = max(int(length.max()), 1)
    taps = torch.zeros(max_len, n_mels, dtype=torch.float32)
    for m in range(n_mels):
        n = int(length[m])
        if n:
            taps[:n, m] = fb[m, int(start[m]):int(start[m]) + n]
    return start, length, taps
