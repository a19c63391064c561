"""GPU parity: every C-ABI kernel against the float64 CPU oracle on the same seeded inputs.
Bar: 1e-3 relative (north_star); the fp32 kernels land orders of magnitude below it."""
import numpy as np
import pytest
import torch

from helpers import rel_err, norm_err, t64, dev
from oracle import sarnet_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("H,W,Cin,Cout,k,stride", [
    (37, 80, 1, 64, 7, 2),      # stem shape (Cin=1, scalar gather path)
    (25, 20, 64, 32, 3, 1),     # s1b1.conv1 thin-res34
    (26, 11, 32, 64, 3, 2),     # strided, even/odd sizes -> asymmetric SAME pads
    (13, 5, 64, 128, 1, 2),     # 1x1 projection shortcut, 'valid'
    (7, 3, 8, 12, 3, 1),        # ragged channel counts (Cout % 64 != 0)
])
def test_conv2d_same_vs_oracle(cuda_device, H, W, Cin, Cout, k, stride):
    from aesrc2020_b200 import ops
    from aesrc2020_b200.config import same_pad
    rng = np.random.RandomState(H * 131 + Cout)
    B = 3
    x = rng.randn(B, H, W, Cin)
    w = rng.randn(k, k, Cin, Cout) / np.sqrt(k * k * Cin)
    b = rng.randn(Cout) * 0.1
    padding = "valid" if k == 1 else "same"
    want = O.conv2d(t64(x), t64(w), t64(b), stride, padding)
    if padding == "same":
        Ho, pt, _ = same_pad(H, k, stride)
        Wo, pl, _ = same_pad(W, k, stride)
    else:
        Ho, Wo, pt, pl = (H - 1) // stride + 1, (W - 1) // stride + 1, 0, 0
    got = ops.conv2d(dev(x), dev(w), dev(b), stride=stride, pad_t=pt, pad_l=pl, out_hw=(Ho, Wo))
    assert norm_err(got, want) < 1e-5


def test_conv2d_fused_pre_post_residual(cuda_device):
    """BN->ReLU on the conv input must NOT be applied to the SAME padding (pads are zeros of
    the activated tensor, resnet.py:47-65), residual add and post BN->ReLU in the epilogue."""
    from aesrc2020_b200 import ops
    rng = np.random.RandomState(5)
    B, H, W, Cin, Cout = 2, 9, 6, 16, 24
    x = rng.randn(B, H, W, Cin)
    w = rng.randn(3, 3, Cin, Cout) * 0.1
    b = rng.randn(Cout) * 0.1
    ps, pt_ = rng.uniform(0.5, 1.5, Cin), rng.randn(Cin) + 0.5       # shift > 0: relu(shift) != 0 at the pads
    qs, qt = rng.uniform(0.5, 1.5, Cout), rng.randn(Cout) * 0.2
    res = rng.randn(B, H, W, Cout)
    act_in = torch.relu(t64(x) * t64(ps) + t64(pt_))
    want = O.conv2d(act_in, t64(w), t64(b), 1, "same") + t64(res)
    want = torch.relu(want * t64(qs) + t64(qt))
    got = ops.conv2d(dev(x), dev(w), dev(b), stride=1, pad_t=1, pad_l=1, out_hw=(H, W),
                     pre=(dev(ps), dev(pt_)), post=(dev(qs), dev(qt)), residual=dev(res), act="relu")
    assert norm_err(got, want) < 1e-5


def test_maxpool_same(cuda_device):
    from aesrc2020_b200 import ops
    from aesrc2020_b200.config import same_pad
    rng = np.random.RandomState(1)
    for H, W in ((150, 40), (11, 7)):
        x = (rng.randn(2, H, W, 8) - 2.0).astype(np.float32).astype(np.float64)   # negative: padded cells must never win
        want = O.maxpool_same(t64(x))
        Ho, pt, _ = same_pad(H, 3, 2)
        Wo, pl, _ = same_pad(W, 3, 2)
        got = ops.maxpool2d(dev(x), k=3, stride=2, pad_t=pt, pad_l=pl, out_hw=(Ho, Wo))
        assert norm_err(got, want) == 0.0


def test_layernorm_eps_1e14(cuda_device):
    from aesrc2020_b200 import ops
    rng = np.random.RandomState(2)
    for C in (256, 512):
        x = rng.randn(37, C) * 3 + 1
        x[3] = 0.25                               # constant row -> exactly beta
        g, b = rng.uniform(0.5, 1.5, C), rng.randn(C)
        want = O.layernorm(t64(x), {"n/gamma": t64(g), "n/beta": t64(b)}, "n")
        got = ops.layernorm(dev(x), dev(g), dev(b))
        assert norm_err(got, want) < 1e-5
        assert torch.allclose(got[3].cpu(), torch.from_numpy(b).float(), atol=0, rtol=0)


@pytest.mark.parametrize("seq", [True, False])
@pytest.mark.parametrize("B,S,Din", [(5, 21, 256), (2, 7, 512), (61, 9, 256), (130, 4, 256)])   # 61: 16 utterances per cluster, ragged last group; 130: the 32-utterance variant
def test_bigru_vs_oracle(cuda_device, B, S, Din, seq):
    from aesrc2020_b200 import model as mdl
    rng = np.random.RandomState(B * 7 + S)
    u = 256
    x = rng.randn(B, S, Din)
    w = {}
    for d in ("forward", "backward"):
        w[d + "/kernel"] = rng.uniform(-0.08, 0.08, (Din, 3 * u))
        w[d + "/recurrent_kernel"] = rng.randn(u, 3 * u) / np.sqrt(u)
        w[d + "/bias"] = rng.randn(6 * u) * 0.1
    layer = mdl.BIGRU(u, seq=seq, name="g")
    layer.set_weights_dict(w)
    got = layer(dev(x))
    want = O.bigru(t64(x), {"g/" + k: t64(v) for k, v in w.items()}, "g", seq=seq)
    assert norm_err(got, want) < 2e-5
    # 16 or 32 utterances per cluster (sar_bigru_nb_fwd) is a scheduling choice: bitwise the same result
    outs = []
    for nb in (16, 32):
        layer.nb = nb
        outs.append(layer(dev(x)))
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], got)


@pytest.mark.parametrize("mode,K,G,S,D", [("gvlad", 64, 8, 48, 256), ("vlad", 64, 0, 48, 256),
                                          ("gvlad", 8, 2, 21, 256), ("gvlad", 10, 3, 114, 256),
                                          ("vlad", 4, 0, 5, 256), ("gvlad", 64, 8, 75, 256)])
def test_vlad_vs_literal_5d_formula(cuda_device, mode, K, G, S, D):
    """Fused kernel vs the literal (B,1,S,K+G,D) broadcast of VLAD.py:33-48."""
    from aesrc2020_b200 import ops, VLAD
    rng = np.random.RandomState(K + S)
    B = 3
    feat = rng.randn(B, 1, S, D)
    wa = rng.randn(D, K + G) / np.sqrt(D) * 3
    ba = rng.randn(K + G) * 0.1
    cen = rng.randn(K + G, D) / np.sqrt(D)
    score = t64(feat) @ t64(wa) + t64(ba)
    want = O.vlad_pooling(t64(feat), score, t64(cen), mode, K)
    got = ops.vlad(dev(feat.reshape(B, S, D)), dev(wa), dev(ba), dev(cen), K, G)
    assert norm_err(got, want) < 2e-5
    # the bare VladPooling([feat, cluster_score]) call surface (external scores)
    layer = VLAD.VladPooling(mode, K, G, centers=cen)
    got2 = layer([dev(feat), score.float().cuda()])
    assert tuple(got2.shape) == layer.compute_output_shape([feat.shape, score.shape])
    assert norm_err(got2, want) < 2e-5


def test_vlad_zero_input_stays_zero(cuda_device):
    from aesrc2020_b200 import ops
    K, G, S, D = 8, 2, 6, 256
    cen = np.zeros((K + G, D))
    got = ops.vlad(torch.zeros(2, S, D, device="cuda"), dev(np.zeros((D, K + G))), dev(np.zeros(K + G)), dev(cen), K, G)
    assert float(got.abs().max()) == 0.0         # 0 / sqrt(max(0, 1e-12)) = 0


@pytest.mark.parametrize("M,K,N", [(64, 16384, 256), (5, 2048, 256), (130, 4100, 64)])
def test_gemm_splitk(cuda_device, M, K, N):
    from aesrc2020_b200 import ops
    rng = np.random.RandomState(M)
    a, w, b = rng.randn(M, K), rng.randn(K, N) / np.sqrt(K), rng.randn(N)
    want = t64(a) @ t64(w) + t64(b)
    got = ops.gemm_splitk(dev(a), dev(w), dev(b))
    assert norm_err(got, want) < 1e-5
    assert torch.equal(got, ops.gemm_splitk(dev(a), dev(w), dev(b)))     # deterministic


@pytest.mark.parametrize("kind,m", [("softmax", 0.3), ("sphereface", 1.35), ("cosface", 0.35), ("arcface", 0.5),
                                    ("arcface", 0.3), ("circleloss", 0.2)])
def test_heads_vs_oracle(cuda_device, kind, m):
    from aesrc2020_b200 import ops
    rng = np.random.RandomState(11)
    B, D, n = 9, 256, 8
    x = rng.randn(B, D)
    x[0] *= 1e-3
    W = rng.uniform(-0.15, 0.15, (D, n))
    W[:, 0] = x[1] / np.linalg.norm(x[1]) * 0.7       # cos ~ +1 for sample 1 -> exercises the acos clip
    lab = rng.randint(0, n, B)
    lab[1] = 0
    y = np.eye(n)[lab]
    key = "h/W" if kind in ("sphereface", "cosface", "arcface") else "h/kernel"
    want_out, want_logits = O.disc_head(t64(x), {key: t64(W)}, t64(y), kind, m, "h")
    r = ops.head(None, None, emb_d=dev(x), wd=dev(W), onehot=dev(y), n_classes=n, head_kind=kind, margin=m)
    assert rel_err(r["y_disc_logits"], want_logits, floor=1e-3) < 1e-3
    assert rel_err(r["y_disc"], want_out, floor=1e-6) < 1e-3
    if kind == "circleloss":
        want_loss = O.circle_loss(t64(y), want_out, 256.0, m)
    else:
        want_loss = O.categorical_crossentropy(t64(y), want_out)
    assert rel_err(r["sample_stats"][:, 1], want_loss, floor=1e-4) < 1e-3
    want_acc = (want_out.argmax(-1) == t64(y).argmax(-1)).double()
    assert torch.equal(r["sample_stats"][:, 3].cpu().double(), want_acc)


def test_losses_module_surface(cuda_device):
    """losses.ArcFace/CosFace/SphereFace([x, y]) and losses.circle_loss(y_true, y_pred)."""
    from aesrc2020_b200 import losses as ls
    rng = np.random.RandomState(3)
    B, D, n = 6, 256, 8
    x, W = rng.randn(B, D), rng.uniform(-0.1, 0.1, (D, n))
    y = np.eye(n)[rng.randint(0, n, B)]
    for cls, kind in ((ls.ArcFace, "arcface"), (ls.CosFace, "cosface"), (ls.SphereFace, "sphereface")):
        layer = cls(n_classes=n, m=0.3, name="y_disc", W=W)
        got = layer([dev(x), dev(y)])
        want, _ = O.disc_head(t64(x), {"h/W": t64(W)}, t64(y), kind, 0.3, "h")
        assert rel_err(got, want) < 1e-3
        assert layer.compute_output_shape(None) == (None, n)
    cos = np.tanh(rng.randn(B, n))
    got = ls.circle_loss(dev(y), dev(cos), gamma=256, margin=0.25)
    want = O.circle_loss(t64(y), t64(cos), 256.0, 0.25)
    assert rel_err(got, want, floor=1e-4) < 1e-3


def test_classifier_and_ce(cuda_device):
    from aesrc2020_b200 import ops
    rng = np.random.RandomState(4)
    B, D, n = 7, 256, 8
    x = rng.randn(B, D)
    p = {"a/kernel": rng.randn(D, 64) / 16, "a/bias": rng.randn(64) * 0.1, "b/kernel": rng.randn(64, 64) / 8,
         "b/bias": rng.randn(64) * 0.1, "c/kernel": rng.randn(64, n) / 8, "c/bias": rng.randn(n) * 0.1}
    p64 = {k: t64(v) for k, v in p.items()}
    logits = O.dense(O.dense(O.dense(t64(x), p64, "a", "relu"), p64, "b", "relu"), p64, "c", None)
    y = np.eye(n)[rng.randint(0, n, B)]
    r = ops.head(dev(x), tuple(dev(p[k]) for k in ("a/kernel", "a/bias", "b/kernel", "b/bias", "c/kernel", "c/bias")),
                 onehot=dev(y), n_classes=n, head_kind=None)
    assert rel_err(r["y_accent_logits"], logits, floor=1e-3) < 1e-4
    assert rel_err(r["y_accent"], torch.softmax(logits, -1)) < 1e-4
    want_loss = O.categorical_crossentropy(t64(y), torch.softmax(logits, -1))
    assert rel_err(r["sample_stats"][:, 0], want_loss, floor=1e-4) < 1e-4


@pytest.mark.parametrize("S,C,Lmax", [(48, 1000, 72), (114, 1000, 72), (9, 6, 4)])
def test_ctc_vs_oracle(cuda_device, S, C, Lmax):
    from aesrc2020_b200 import ops
    rng = np.random.RandomState(S + C)
    B = 6
    logits = rng.randn(B, S, C) * 2
    labels = np.full((B, Lmax), 2.0)
    lab_len = np.zeros(B, np.int32)
    for b in range(B):
        L = int(rng.randint(1, min(Lmax, S // 2) + 1))
        ids = rng.randint(0, C - 1, L)
        if b == 1 and L >= 2:
            ids[1] = ids[0]                          # repeated label needs a blank between
        labels[b, :L] = ids
        lab_len[b] = L
    lab_len[2] = min(Lmax, S // 2)
    labels[2, :lab_len[2]] = rng.randint(0, C - 1, lab_len[2])
    in_len = np.full(B, S, np.int32)
    in_len[3] = min(S, max(S - 3, 2 * int(lab_len[3]) + 1))  # shorter input_length (API allows it)
    probs = torch.softmax(t64(logits), -1)
    want = O.ctc_batch_cost(t64(labels), probs, in_len, lab_len)
    loss, status, p = ops.ctc(dev(logits), dev(labels), torch.from_numpy(in_len).cuda(), torch.from_numpy(lab_len).cuda(),
                              want_probs=True)
    assert int(status.abs().sum()) == 0
    assert rel_err(loss.reshape(-1, 1), want) < 1e-4
    assert norm_err(p, probs) < 1e-5
    # padded rows (sar_ctc_ld_fwd: the tensor-core ctc_pred Dense pads its columns to a multiple of 32): garbage in
    # the pad columns must not matter, and the result is bitwise the dense-row one
    ld = (C + 31) // 32 * 32 + 32
    padded = torch.full((B, S, ld), 1e30, device="cuda", dtype=torch.float32)
    padded[..., :C] = dev(logits)
    loss2, status2, p2 = ops.ctc(padded, dev(labels), torch.from_numpy(in_len).cuda(), torch.from_numpy(lab_len).cuda(),
                                 want_probs=True, classes=C)
    assert torch.equal(loss2, loss) and torch.equal(status2, status) and torch.equal(p2, p)


def test_ctc_infeasible_and_bad_labels(cuda_device):
    from aesrc2020_b200 import ops
    S, C, Lmax = 4, 6, 4
    logits = torch.zeros(3, S, C, device="cuda")
    labels = dev(np.array([[1, 1, 1, 2], [0, 1, 2, 3], [5, 0, 0, 0]], dtype=np.float32))   # repeats need 6 frames; 5 == blank
    lab_len = torch.tensor([3, 2, 1], dtype=torch.int32, device="cuda")
    in_len = torch.full((3,), S, dtype=torch.int32, device="cuda")
    loss, status, _ = ops.ctc(logits, labels, in_len, lab_len)
    assert status.cpu().tolist() == [1, 0, 2]
    assert torch.isinf(loss[0]) and torch.isfinite(loss[1])


def test_loss_reduce(cuda_device):
    from aesrc2020_b200 import ops
    rng = np.random.RandomState(0)
    B = 300
    stats, ctc, bn = rng.rand(B, 4), rng.rand(B) * 100, rng.rand(B, 4)
    v = ops.loss_reduce(dev(stats), dev(ctc), dev(bn)).cpu().double().numpy()
    want = [stats[:, 0].sum(), stats[:, 1].sum(), ctc.sum(), bn[:, 1].sum(), stats[:, 2].sum(), stats[:, 3].sum(), B,
            bn[:, 3].sum()]
    assert np.allclose(v, want, rtol=1e-5)


def test_fbank_vs_oracle(cuda_device):
    from aesrc2020_b200 import fbank as fb
    from oracle import fbank_oracle as FO
    rng = np.random.RandomState(9)
    wavs = []
    for n in (16000, 5000, 400, 23456, 161):
        t = np.arange(n) / 16000.0
        wavs.append(0.3 * np.sin(2 * np.pi * (200 + 37 * len(wavs)) * t * (1 + t)) + 0.05 * rng.randn(n))
    T = 120
    x, raw = fb.fbank_batch(wavs, T, return_raw=True)
    for i, w in enumerate(wavs):
        want_raw = FO.fbank(w.astype(np.float32).astype(np.float64))
        nf = want_raw.shape[0]
        assert nf == fb.num_frames(len(w))
        assert norm_err(raw[i, :nf], want_raw) < 1e-4
        want = FO.wav_to_x_data(w.astype(np.float32).astype(np.float64), T)
        assert float(np.max(np.abs(x[i].cpu().numpy() - want))) < 2e-4


def test_bad_arguments_are_rejected(cuda_device):
    from aesrc2020_b200 import ops, _shim
    with pytest.raises(_shim.SarnetError):
        ops.layernorm(torch.zeros(4, 6, device="cuda"), torch.ones(6, device="cuda"), torch.zeros(6, device="cuda"))
    with pytest.raises(_shim.SarnetError):
        ops.bigru(torch.zeros(1, 3, 2, 384, device="cuda"), torch.zeros(2, 128, 384, device="cuda"),
                  torch.zeros(2, 384, device="cuda"))
    with pytest.raises(_shim.SarnetError):
        ops.layernorm(torch.zeros(4, 8), torch.ones(8), torch.zeros(8))        # host tensors at the boundary


def test_ctc_greedy_decode_vs_oracle(cuda_device):
    """sar_ctc_greedy_fwd on padded logit rows vs the oracle's restatement of K.ctc_decode(greedy=True): exact ids."""
    from aesrc2020_b200 import ops
    rng = np.random.RandomState(11)
    B, S, C, ld = 9, 37, 50, 64
    logits = rng.randn(B, S, C).astype(np.float32)
    peak = rng.randint(0, C, size=(B, S))
    peak[:, 5:9] = peak[:, 5:6]                          # repeats to merge
    peak[0] = C - 1                                      # all blank -> empty row
    for b in range(B):
        for t in range(S):
            logits[b, t, peak[b, t]] += 6.0
    logits[3, 2, :] = 1.0                                # exact tie -> class 0
    padded = torch.full((B, S, ld), 1e30, device="cuda", dtype=torch.float32)
    padded[..., :C] = dev(logits)
    for T in (S, 20, 1):
        dec, n = ops.ctc_greedy(padded, fixed_len=T, classes=C)
        dec, n = dec.cpu().numpy(), n.cpu().numpy()
        want = O.ctc_greedy_decode(torch.softmax(t64(logits), -1), T)
        L = want.shape[1]
        assert int(n.max()) == (want >= 0).sum(1).max()
        assert np.array_equal(dec[:, :L], want) and (dec[:, L:] == -1).all()
    lens = torch.from_numpy(rng.randint(0, S + 1, size=B).astype(np.int32)).cuda()
    dec, n = ops.ctc_greedy(padded, lens, classes=C)
    for b in range(B):
        want = O.ctc_greedy_decode(torch.softmax(t64(logits[b:b + 1]), -1), int(lens[b]))[0]
        want = want[want >= 0]
        assert dec[b, :int(n[b])].cpu().numpy().tolist() == want.tolist()


def test_data_loader_on_device_vs_oracle(cuda_device, tmp_path):
    """utils.data_loader: sar_feat_batch_fwd + sar_labels_pack_fwd against the oracle's restatement of
    utils.py:71-117 -- ragged utterances (1 frame ... longer than max_input_len), a constant column, pickle paths and
    in-memory arrays, truncated transcripts; the result feeds model inputs unchanged."""
    from aesrc2020_b200 import utils as us, _shim
    from oracle import fbank_oracle as FO
    import pickle
    rng = np.random.RandomState(5)
    lst = ["u%d" % i for i in range(7)]
    frames = [37, 120, 64, 333, 1, 100, 99]
    data = {u: (rng.rand(n, 80) * 40 - 5).astype(np.float32) for u, n in zip(lst, frames)}
    data["u2"][:, 7] = 3.25
    acc = {u: str((3 * i) % 8) for i, u in enumerate(lst)}
    trans = {u: [int(v) for v in rng.randint(3, 999, size=n)] for u, n in zip(lst, [4, 90, 1, 72, 10, 73, 30])}
    paths = dict(data)
    for u in lst[:3]:                                        # the reference stores pickles (utils.py:14-22,91)
        p = tmp_path / (u + ".pkl")
        with open(p, "wb") as f:
            pickle.dump(data[u], f)
        paths[u] = str(p)
    kw = dict(max_input_len=100, max_ctc_len=72, encoder_len=13, accent_classes=8, bn=1)
    got_x, got_y = us.data_loader(lst, True, True, True, paths, acc, trans, **kw)
    want_x, want_y = FO.data_loader(lst, True, True, True, data, acc, trans, **kw)
    assert set(got_x) == set(want_x) and set(got_y) == set(want_y)
    for k, w in want_x.items():
        g = got_x[k].cpu().numpy()
        assert g.shape == w.shape and g.dtype == w.dtype, k
        if k == "x_data":
            assert np.abs(g - w).max() < 2e-6
            assert not g[4, 1:].any() and not g[0, 37:].any()      # zero padding
        else:
            assert np.array_equal(g, w), k
    assert np.array_equal(got_y["y_disc_bn"].cpu().numpy(), want_y["y_disc_bn"])
    with pytest.raises(_shim.SarnetError):
        us.data_loader(lst, False, True, True, data, dict(acc, u3="8"), None, **kw)
