"""Float64 numpy restatement of the feature front-end (TEST INFRASTRUCTURE ONLY).

Follows local/make_fbank.py:24-28 (`psf.fbank(y, samplerate=sr, nfilt=80)[0]`)
and utils.py:35-46 (`feat_norm`, `feat_reshape`).  The arithmetic of `psf.fbank`
lives in the un-vendored, un-pinned dependency `python_speech_features`
(0.6 is the only release series; not installed here) -- restated below from its
published algorithm (sigproc.preemphasis / framesig / powspec and
base.get_filterbanks).  `feat_norm` is sklearn's MinMaxScaler, which IS
importable here and is cross-checked in tests/test_oracle_fbank.py.
Parity status for this file: unpinned for psf itself (no copy of the package or of its outputs is available offline);
its stages are cross-checked against independent implementations in tests/test_oracle_fbank.py -- framing + rectangular
window + |rfft_512|^2 / 512 against torch.stft, the mel scale and filter placement against torchaudio's HTK filterbank --
and MinMaxScaler is pinned against sklearn.
"""
from __future__ import annotations

import math

import numpy as np


def _round_half_up(x: float) -> int:
    return int(math.floor(x + 0.5))


def hz2mel(hz):
    return 2595.0 * np.log10(1.0 + hz / 700.0)


def mel2hz(mel):
    return 700.0 * (10.0 ** (mel / 2595.0) - 1.0)


def num_frames(n_samples: int, frame_len: int = 400, frame_step: int = 160) -> int:
    """psf.sigproc.framesig frame count: 1 + ceil((n - len)/step), 1 if n <= len."""
    if n_samples <= frame_len:
        return 1
    return 1 + int(math.ceil((1.0 * n_samples - frame_len) / frame_step))


def get_filterbanks(nfilt: int = 80, nfft: int = 512, samplerate: int = 16000,
                    lowfreq: float = 0.0, highfreq: float | None = None) -> np.ndarray:
    """psf.base.get_filterbanks: (nfilt, nfft//2+1) triangular filters on floor()ed bins."""
    highfreq = highfreq or samplerate / 2
    melpoints = np.linspace(hz2mel(lowfreq), hz2mel(highfreq), nfilt + 2)
    bins = np.floor((nfft + 1) * mel2hz(melpoints) / samplerate)
    fb = np.zeros([nfilt, nfft // 2 + 1])
    for j in range(nfilt):
        for i in range(int(bins[j]), int(bins[j + 1])):
            fb[j, i] = (i - bins[j]) / (bins[j + 1] - bins[j])
        for i in range(int(bins[j + 1]), int(bins[j + 2])):
            fb[j, i] = (bins[j + 2] - i) / (bins[j + 2] - bins[j + 1])
    return fb


def fbank(signal: np.ndarray, samplerate: int = 16000, nfilt: int = 80, nfft: int = 512,
          winlen: float = 0.025, winstep: float = 0.01, preemph: float = 0.97) -> np.ndarray:
    """`psf.fbank(...)[0]` as called at local/make_fbank.py:27: LINEAR mel filterbank
    energies (no log), rectangular window, |rfft|^2 / nfft, zeros replaced by eps."""
    signal = np.asarray(signal, dtype=np.float64)
    sig = np.append(signal[0], signal[1:] - preemph * signal[:-1])
    frame_len = _round_half_up(winlen * samplerate)
    frame_step = _round_half_up(winstep * samplerate)
    n = num_frames(len(sig), frame_len, frame_step)
    padlen = (n - 1) * frame_step + frame_len
    pad = np.concatenate([sig, np.zeros(padlen - len(sig))])
    idx = np.arange(frame_len)[None, :] + frame_step * np.arange(n)[:, None]
    frames = pad[idx]
    pspec = (1.0 / nfft) * np.square(np.abs(np.fft.rfft(frames, nfft)))
    feat = pspec @ get_filterbanks(nfilt, nfft, samplerate).T
    return np.where(feat == 0, np.finfo(float).eps, feat)


def feat_norm(feat: np.ndarray) -> np.ndarray:
    """utils.py:35-36: MinMaxScaler().fit_transform -- per column over time to [0,1];
    constant columns map to 0 (sklearn replaces a zero range by 1)."""
    feat = np.asarray(feat, dtype=np.float64)
    mn, mx = feat.min(axis=0), feat.max(axis=0)
    rng = mx - mn
    rng = np.where(rng == 0.0, 1.0, rng)
    scale = 1.0 / rng
    return feat * scale + (0.0 - mn * scale)


def feat_reshape(feat: np.ndarray, max_len: int = 1200) -> np.ndarray:
    """utils.py:39-46: truncate to max_len rows or zero-pad at the end."""
    h, w = feat.shape
    if h >= max_len:
        return feat[:max_len]
    out = np.zeros((max_len, w))
    out[:h] = feat
    return out


def wav_to_x_data(signal: np.ndarray, max_len: int) -> np.ndarray:
    """make_fbank.py:27 -> utils.py:91,102: (max_len, 80, 1) float32 model input."""
    return np.float32(feat_reshape(feat_norm(fbank(signal)), max_len)[:, :, None])


def data_loader(lst, ctc_enable=False, ar_enable=False, disc_enable=False, data_dct=None, accent_dct=None, trans_dct=None,
                max_input_len=1200, max_ctc_len=72, encoder_len=100, accent_classes=8, bn=0):
    """utils.py:71-117 restated with numpy (feat_norm / feat_reshape above; text_ids_norm utils.py:57-63 with EOS_ID = 2;
    to_categorical): `data_dct[utt]` is the (frames, dims) feature matrix itself (the reference unpickles it, utils.py:91)."""
    inputs, in_len, lab_len, labels, accents = [], [], [], [], []
    for utt in lst:
        inputs.append(feat_reshape(feat_norm(np.asarray(data_dct[utt], dtype=np.float64)), max_input_len))
        if ctc_enable and trans_dct:
            ids = list(trans_dct[utt])[:max_ctc_len]
            in_len.append(encoder_len)
            lab_len.append(min(len(trans_dct[utt]), max_ctc_len))
            labels.append(ids + [2] * (max_ctc_len - len(ids)))
        if ar_enable and accent_dct:
            oh = np.zeros(accent_classes, dtype=np.float32)      # keras to_categorical: dtype="float32"
            oh[int(accent_dct[utt])] = 1.0
            accents.append(oh)
    input_data = {"x_data": np.float32(np.expand_dims(np.asarray(inputs), axis=3))}
    output_data = {}
    if ctc_enable:
        input_data["x_ctc_in_len"] = np.int32(np.expand_dims(np.asarray(in_len), axis=1))
        input_data["x_ctc_out_len"] = np.int32(np.expand_dims(np.asarray(lab_len), axis=1))
        input_data["x_ctc_label"] = np.float32(np.asarray(labels))
        output_data["y_ctc_loss"] = np.zeros([len(lst)])
    if ar_enable:
        output_data["y_accent"] = np.asarray(accents)
    if disc_enable:
        input_data["x_accent"] = np.asarray(accents)
        output_data["y_disc"] = np.asarray(accents)
        if bn:
            output_data["y_disc_bn"] = np.asarray(accents)
    return input_data, output_data
