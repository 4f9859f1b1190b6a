"""Oracle (test infrastructure, not product code): log-mel front end.

Restates ``AugmentMelSTFT.forward`` (reference models/preprocess.py:40-67) step by
step with explicit framing + rFFT so that every stage can be evaluated in fp64
("truth") or fp32.  The Kaldi filterbank follows
``torchaudio.compliance.kaldi.get_mel_banks`` (torchaudio 2.11, kaldi.py:436-511;
the reference calls it at models/preprocess.py:52-53) for vtln_warp == 1.0.
"""
import math

import torch


def kaldi_mel_banks(n_mels, n_fft, sr, fmin, fmax, dtype=torch.float32):
    """[n_mels, n_fft//2 + 1] triangular filters; last column is the zero pad the
    reference appends (models/preprocess.py:54)."""
    nyq = 0.5 * sr
    if fmax <= 0.0:
        fmax += nyq
    bin_width = sr / n_fft
    mel_lo = 1127.0 * math.log(1.0 + fmin / 700.0)
    mel_hi = 1127.0 * math.log(1.0 + fmax / 700.0)
    delta = (mel_hi - mel_lo) / (n_mels + 1)
    b = torch.arange(n_mels).unsqueeze(1)
    left = mel_lo + b * delta
    center = mel_lo + (b + 1.0) * delta
    right = mel_lo + (b + 2.0) * delta
    if dtype == torch.float64:
        left, center, right = left.double(), center.double(), right.double()
        mel = 1127.0 * (1.0 + bin_width * torch.arange(n_fft // 2, dtype=torch.float64) / 700.0).log()
    else:
        mel = 1127.0 * (1.0 + bin_width * torch.arange(n_fft / 2) / 700.0).log()
    mel = mel.unsqueeze(0)
    up = (mel - left) / (center - left)
    down = (right - mel) / (right - center)
    fb = torch.clamp(torch.minimum(up, down), min=0.0)
    return torch.nn.functional.pad(fb, (0, 1)).to(dtype)


def mel_forward(x, n_mels=128, sr=32000, win_length=800, hopsize=320, n_fft=1024,
                fmin=0.0, fmax=15000.0, dtype=torch.float32, window=None):
    """x: [B, N] waveform -> [B, n_mels, 1 + (N-1)//hopsize] normalised log-mel (eval path)."""
    x = x.to(dtype)
    # pre-emphasis, valid cross-correlation with [-0.97, 1]  (preprocess.py:41, buffer :30)
    p = x[:, 1:] - 0.97 * x[:, :-1]
    L = p.shape[1]
    # torch.stft(center=True) reflect-pads n_fft//2 each side  (preprocess.py:42-43)
    half = n_fft // 2
    idx = torch.arange(-half, L + half)
    idx = idx.abs()
    idx = torch.where(idx >= L, 2 * (L - 1) - idx, idx)
    padded = p[:, idx]
    n_frames = 1 + L // hopsize
    frames = padded.unfold(1, n_fft, hopsize)[:, :n_frames]            # [B, T, n_fft]
    if window is None:
        window = torch.hann_window(win_length, periodic=False, dtype=torch.float64)   # preprocess.py:22-24
    w = torch.zeros(n_fft, dtype=torch.float64)
    lp = (n_fft - win_length) // 2
    w[lp:lp + win_length] = window.double()
    spec = torch.fft.rfft(frames * w.to(dtype), dim=-1)                # [B, T, 513]
    power = spec.real ** 2 + spec.imag ** 2                            # preprocess.py:44
    fb = kaldi_mel_banks(n_mels, n_fft, sr, fmin, fmax, dtype=dtype)
    mel = torch.matmul(fb, power.transpose(1, 2))                      # preprocess.py:56-57
    mel = (mel + 0.00001).log()                                        # preprocess.py:59
    return (mel + 4.5) / 5.0                                           # preprocess.py:65
