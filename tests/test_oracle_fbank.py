"""fbank front-end oracle: frame-count rule, filterbank shape, tone localisation, MinMaxScaler."""
import numpy as np
from sklearn.preprocessing import MinMaxScaler

from oracle import fbank_oracle as FO


def test_frame_count_formula():
    assert FO.num_frames(100) == 1 and FO.num_frames(400) == 1 and FO.num_frames(401) == 2
    assert FO.num_frames(560) == 2 and FO.num_frames(561) == 3
    assert FO.num_frames(16000 * 5) == 499                  # 5 s @ 16 kHz -> 499 frames
    assert FO.fbank(np.random.RandomState(0).randn(16000)).shape == (99, 80)


def test_filterbank_structure():
    fb = FO.get_filterbanks()
    assert fb.shape == (80, 257) and fb.min() >= 0 and fb.max() <= 1.0
    # floor()ed bin edges collide at the low end for nfilt=80 / nfft=512: filter 2 is EMPTY, so its
    # energy is always 0 -> eps -> a constant column -> all zeros after feat_norm (a reference quirk)
    assert np.flatnonzero(fb.sum(1) == 0).tolist() == [2]
    feat = FO.fbank(np.random.RandomState(3).randn(4000))
    assert (feat[:, 2] == np.finfo(float).eps).all() and (FO.feat_norm(feat)[:, 2] == 0).all()


def test_pure_tone_peaks_in_expected_filter():
    sr, f0 = 16000, 1000.0
    t = np.arange(sr) / sr
    feat = FO.fbank(np.sin(2 * np.pi * f0 * t))
    fb = FO.get_filterbanks()
    k = int(round(f0 * 512 / sr))
    expect = int(np.argmax(fb[:, k]))
    assert int(np.argmax(feat[10])) == expect
    assert (feat > 0).all()                                  # zeros replaced by eps, never log-ed


def test_feat_norm_equals_sklearn_minmax_and_constant_column():
    rng = np.random.RandomState(1)
    f = np.abs(rng.randn(50, 80)) * 1e-3
    f[:, 7] = 0.125                                          # constant column -> all zeros
    want = MinMaxScaler().fit_transform(f)                   # utils.py:5-6,35-36
    got = FO.feat_norm(f)
    assert np.array_equal(got, want)
    assert (got[:, 7] == 0).all() and got.min() == 0.0 and abs(got.max() - 1.0) < 1e-12


def test_feat_reshape_pad_and_truncate():
    f = np.arange(12.0).reshape(6, 2)
    assert FO.feat_reshape(f, 4).shape == (4, 2) and np.array_equal(FO.feat_reshape(f, 4), f[:4])
    p = FO.feat_reshape(f, 9)
    assert p.shape == (9, 2) and np.array_equal(p[:6], f) and (p[6:] == 0).all()
    x = FO.wav_to_x_data(np.random.RandomState(2).randn(8000), 60)
    assert x.shape == (60, 80, 1) and x.dtype == np.float32 and (x[49:] == 0).all()


def test_data_loader_oracle_matches_sklearn_and_reference_contract():
    """oracle data_loader (utils.py:71-117): x_data against sklearn's MinMaxScaler itself (the reference's MMS), label
    packing against the reference's text_ids_norm / to_categorical semantics, shapes and dtypes of utils.py:102-116."""
    from sklearn.preprocessing import MinMaxScaler
    from oracle import fbank_oracle as FO
    rng = np.random.RandomState(3)
    lst = ["u%d" % i for i in range(5)]
    frames = [37, 120, 64, 200, 1]
    data = {u: rng.rand(n, 80) * 50 for u, n in zip(lst, frames)}
    data["u2"][:, 7] = 3.25                                 # constant column -> zeros
    acc = {u: str(i % 8) for i, u in enumerate(lst)}
    trans = {u: list(rng.randint(3, 999, size=n)) for u, n in zip(lst, [4, 90, 1, 72, 10])}
    x, y = FO.data_loader(lst, True, True, True, data, acc, trans, max_input_len=100, max_ctc_len=72, encoder_len=13, bn=1)
    assert x["x_data"].shape == (5, 100, 80, 1) and x["x_data"].dtype == np.float32
    for i, u in enumerate(lst):
        want = MinMaxScaler().fit_transform(data[u])
        n = min(frames[i], 100)
        assert np.allclose(x["x_data"][i, :n, :, 0], want[:n], atol=1e-6)
        assert not x["x_data"][i, n:].any()
    assert x["x_ctc_label"].dtype == np.float32 and x["x_ctc_label"].shape == (5, 72)
    assert x["x_ctc_label"][0].tolist() == [float(v) for v in trans["u0"]] + [2.0] * 68
    assert x["x_ctc_label"][1].tolist() == [float(v) for v in trans["u1"][:72]]
    assert x["x_ctc_out_len"].reshape(-1).tolist() == [4, 72, 1, 72, 10] and x["x_ctc_out_len"].dtype == np.int32
    assert x["x_ctc_in_len"].reshape(-1).tolist() == [13] * 5 and x["x_ctc_in_len"].shape == (5, 1)
    assert x["x_accent"].argmax(1).tolist() == [0, 1, 2, 3, 4] and x["x_accent"].sum() == 5
    assert set(y) == {"y_ctc_loss", "y_accent", "y_disc", "y_disc_bn"}


def test_power_spectrum_matches_torch_stft():
    """Independent implementation of the framing + rectangular window + |rfft_512|^2 / 512 stage: torch.stft.  (stft centres the
    400-sample window in its 512-sample buffer -- a 56-sample shift that changes only the phase -- and frames the signal it is
    given, so the pre-emphasised signal is offered with 56 leading zeros and psf's zero padding at the end.)"""
    import torch
    from oracle import fbank_oracle as F
    rng = np.random.RandomState(0)
    for n in (16000, 16123, 399, 401):
        sig = rng.randn(n)
        pre = np.append(sig[0], sig[1:] - 0.97 * sig[:-1])
        nf = F.num_frames(n)
        padlen = (nf - 1) * 160 + 400
        buf = np.concatenate([np.zeros(56), pre, np.zeros(padlen - n + 56)])
        st = torch.stft(torch.from_numpy(buf), n_fft=512, hop_length=160, win_length=400, window=torch.ones(400, dtype=torch.float64),
                        center=False, return_complex=True)
        pspec_t = (st.abs() ** 2 / 512).T.numpy()[:nf]                      # (frames, 257)
        fb = F.get_filterbanks(80, 512, 16000)
        want = pspec_t @ fb.T
        got = F.fbank(sig)
        assert got.shape == (nf, 80)
        assert np.allclose(got, np.where(want == 0, np.finfo(float).eps, want), rtol=1e-9, atol=1e-12)


def test_mel_scale_matches_torchaudio_htk():
    """The mel <-> Hz maps (2595 log10(1 + f/700)) and the 82 filter edge frequencies against torchaudio's HTK mel scale; the
    oracle's FFT-bin edges are psf's floor((nfft + 1) f / sr) of exactly those frequencies, and every triangular filter peaks
    within one bin of torchaudio's continuous triangle of the same index."""
    import torch
    import torchaudio.functional as AF
    from oracle import fbank_oracle as F
    hz = np.array([0.0, 100.0, 440.0, 1000.0, 4000.0, 8000.0])
    m = F.hz2mel(hz)
    assert np.allclose(m, 2595.0 * np.log10(1 + hz / 700.0)) and np.allclose(F.mel2hz(m), hz)
    ta = AF.melscale_fbanks(n_freqs=257, f_min=0.0, f_max=8000.0, n_mels=80, sample_rate=16000, norm=None, mel_scale="htk").numpy()  # (257, 80)
    fb = F.get_filterbanks(80, 512, 16000)                                   # (80, 257)
    assert fb.shape == ta.T.shape
    # psf's floor()ed bin edges collapse where the mel spacing is finer than an FFT bin: at nfilt = 80 / nfft = 512 filter 2 is
    # EMPTY (its three edges fall on bins 1, 2, 2), so that feature is identically eps -- a property of the reference's front-end
    # the device kernel reproduces
    empty = [j for j in range(80) if not fb[j].any()]
    assert empty == [2]
    live = np.array([j for j in range(80) if j not in empty])
    pk_o, pk_t = fb.argmax(1), ta.argmax(0)
    assert np.all(np.abs(pk_o[live] - pk_t[live]) <= 1), (pk_o, pk_t)
    edges_hz = F.mel2hz(np.linspace(F.hz2mel(0.0), F.hz2mel(8000.0), 82))
    bins = np.floor((512 + 1) * edges_hz / 16000)
    for j in live:                                                           # support of filter j = [bin_j, bin_{j+2}]
        nz = np.nonzero(fb[j])[0]
        assert nz.min() >= bins[j] and nz.max() <= bins[j + 2]
    # torchaudio's own edge frequencies (all_freqs where a triangle starts) agree with the oracle's to a fraction of a bin
    f_axis = np.linspace(0, 8000, 257)
    starts_t = np.array([f_axis[np.nonzero(ta[:, j])[0].min()] for j in range(80)])
    assert np.all(np.abs(starts_t - edges_hz[:80]) <= 2 * (8000 / 256))
