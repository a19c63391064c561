"""On-device feature front-end: wav -> x_data, replacing local/make_fbank.py:24-28
(`psf.fbank(y, samplerate=sr, nfilt=80)[0]`) and utils.py:35-46 (`feat_norm`, `feat_reshape`).

The filterbank matrix is host-side setup (80 x 257 floats, built once, float64 then cast);
all per-sample arithmetic runs in csrc/fbank.cu.
"""
from __future__ import annotations

import math
from typing import List, Sequence

import numpy as np
import torch

from . import ops

NFFT, NFILT, FRAME_LEN, FRAME_STEP, SAMPLERATE = 512, 80, 400, 160, 16000


def hz2mel(hz):
    return 2595.0 * np.log10(1.0 + hz / 700.0)


def mel2hz(mel):
    return 700.0 * (10.0 ** (mel / 2595.0) - 1.0)


def mel_filterbank(nfilt: int = NFILT, nfft: int = NFFT, samplerate: int = SAMPLERATE) -> np.ndarray:
    """python_speech_features.get_filterbanks: (nfilt, nfft//2+1), triangles on floor()ed bins."""
    pts = np.linspace(hz2mel(0.0), hz2mel(samplerate / 2), nfilt + 2)
    b = np.floor((nfft + 1) * mel2hz(pts) / samplerate)
    fb = np.zeros((nfilt, nfft // 2 + 1))
    for j in range(nfilt):
        lo, mid, hi = int(b[j]), int(b[j + 1]), int(b[j + 2])
        for i in range(lo, mid):
            fb[j, i] = (i - b[j]) / (b[j + 1] - b[j])
        for i in range(mid, hi):
            fb[j, i] = (b[j + 2] - i) / (b[j + 2] - b[j + 1])
    return fb


def num_frames(n_samples: int) -> int:
    if n_samples <= FRAME_LEN:
        return 1
    return 1 + int(math.ceil((n_samples - FRAME_LEN) / FRAME_STEP))


_FB_CACHE = {}


def fbank_batch(wavs: Sequence[np.ndarray], max_len: int, device="cuda", return_raw: bool = False):
    """List of 16 kHz float waveforms -> x_data (B, max_len, 80, 1) on device: fbank ->
    per-utterance min-max -> truncate / zero-pad (make_fbank.py:27 -> utils.py:91,102)."""
    dev = torch.device(device)
    lens = [int(len(w)) for w in wavs]
    if min(lens) < 1:
        raise ValueError("empty waveform")
    offs = np.zeros(len(wavs) + 1, dtype=np.int64)
    offs[1:] = np.cumsum(lens)
    flat = torch.from_numpy(np.concatenate([np.asarray(w, dtype=np.float32) for w in wavs])).to(dev)
    key = str(dev)
    if key not in _FB_CACHE:
        _FB_CACHE[key] = torch.from_numpy(np.ascontiguousarray(mel_filterbank().T, dtype=np.float32)).to(dev)
    fmax = max(num_frames(n) for n in lens)
    x, feat = ops.fbank(flat, torch.from_numpy(offs).to(dev), _FB_CACHE[key], fmax, max_len)
    x = x.unsqueeze(-1)
    return (x, feat) if return_raw else x
