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
