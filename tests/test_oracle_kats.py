"""Known-answer tests that pin the CPU oracle.  The reference pins nothing
but S(1200,80)=114, so every primitive is checked against a closed form, a brute-force
enumeration or an independent implementation (torch.nn.GRU, F.ctc_loss, sklearn)."""
import itertools
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import sarnet_oracle as O

D = torch.float64


def t64(a):
    return torch.as_tensor(np.asarray(a), dtype=D)


# ------------------------------------------------------------------ shape rule
def test_shape_known_answer_114():
    """train.py:69 ENCODER_LEN = 114 for (1200, 80); utils.py:156-159 cal_descriptors."""
    assert O.cal_descriptors(1200, 80) == 114
    rng = np.random.RandomState(0)
    from aesrc2020_b200 import weights as W
    from aesrc2020_b200.config import SARConfig, encoder_len
    assert encoder_len(1200, 80) == 114
    for T, S in ((200, 21), (300, 30), (500, 48), (800, 75)):
        assert encoder_len(T, 80) == S == O.cal_descriptors(T, 80)


def test_resnet_output_shape_runs_through_oracle():
    from aesrc2020_b200 import weights as W
    from aesrc2020_b200.config import SARConfig
    cfg = SARConfig(input_shape=(97, 80, 1), res_type="res34", res_filters=8, mto="avg")
    w = {k: t64(v) for k, v in W.init_weights(cfg).items()}
    y = O.resnet(torch.rand(1, 97, 80, 1, dtype=D), w, "res34", 8)
    assert tuple(y.shape) == (1, 4, 3, 64) and O.cal_descriptors(97, 80) == 12
    with pytest.raises(NotImplementedError):
        O.resnet(torch.rand(1, 97, 80, 1, dtype=D), w, "res50", 8)


# ------------------------------------------------------------------ TF-SAME conv
@pytest.mark.parametrize("H,W,k,s", [(7, 6, 3, 2), (8, 5, 3, 2), (10, 10, 7, 2), (5, 4, 3, 1), (9, 80, 7, 2)])
def test_same_conv_vs_index_loop(H, W, k, s):
    rng = np.random.RandomState(H * 10 + W)
    x = rng.randn(1, H, W, 2)
    w = rng.randn(k, k, 2, 3)
    b = rng.randn(3)
    got = O.conv2d(t64(x), t64(w), t64(b), s, "same").numpy()
    Ho, Wo = -(-H // s), -(-W // s)
    ph = max((Ho - 1) * s + k - H, 0)
    pw = max((Wo - 1) * s + k - W, 0)
    pt, pl = ph // 2, pw // 2                           # extra padding goes AFTER
    want = np.zeros((1, Ho, Wo, 3))
    for ho in range(Ho):
        for wo in range(Wo):
            for r in range(k):
                for c in range(k):
                    hi, wi = ho * s - pt + r, wo * s - pl + c
                    if 0 <= hi < H and 0 <= wi < W:
                        want[0, ho, wo] += x[0, hi, wi] @ w[r, c]
    want += b
    assert got.shape == want.shape and np.allclose(got, want, atol=1e-12)


def test_same_pad_is_asymmetric_for_stride2():
    assert O.same_pad(500, 7, 2) == (250, 2, 3)
    assert O.same_pad(250, 3, 2) == (125, 0, 1)
    assert O.same_pad(125, 3, 2) == (63, 1, 1)
    assert O.same_pad(32, 3, 2) == (16, 0, 1)
    assert O.same_pad(20, 3, 1) == (20, 1, 1)


def test_maxpool_same_ignores_padding():
    x = -torch.ones(1, 4, 4, 1, dtype=D) * 5
    y = O.maxpool_same(x)
    assert tuple(y.shape) == (1, 2, 2, 1) and float(y.max()) == -5.0     # zero padding would give 0


# ------------------------------------------------------------------ BN / LN / basic block
def test_bn_inference_formula_and_eps():
    rng = np.random.RandomState(1)
    x = rng.randn(3, 5)
    p = {"b/gamma": t64(rng.rand(5) + 0.5), "b/beta": t64(rng.randn(5)), "b/moving_mean": t64(rng.randn(5)),
         "b/moving_variance": t64(rng.rand(5) * 0.01)}
    got = O.batchnorm(t64(x), p, "b").numpy()
    want = p["b/gamma"].numpy() * (x - p["b/moving_mean"].numpy()) / np.sqrt(p["b/moving_variance"].numpy() + 1e-3) \
        + p["b/beta"].numpy()
    assert np.allclose(got, want, atol=1e-13)
    other = p["b/gamma"].numpy() * (x - p["b/moving_mean"].numpy()) / np.sqrt(p["b/moving_variance"].numpy() + 1e-5) \
        + p["b/beta"].numpy()
    assert np.max(np.abs(other - want)) > 1e-3          # eps=1e-3 (Keras) vs 1e-5 (torch) is visible


def test_layernorm_constant_row_gives_beta():
    p = {"n/gamma": t64([2.0, 3.0, 4.0, 5.0]), "n/beta": t64([0.1, 0.2, 0.3, 0.4])}
    y = O.layernorm(t64([[7.0, 7.0, 7.0, 7.0], [1.0, 2.0, 3.0, 4.0]]), p, "n")
    assert torch.equal(y[0], p["n/beta"])
    x = np.array([1.0, 2.0, 3.0, 4.0])
    want = (x - 2.5) / np.sqrt(1.25 + 1e-14) * p["n/gamma"].numpy() + p["n/beta"].numpy()
    assert np.allclose(y[1].numpy(), want, atol=1e-12)


def test_basic_block_identity_kernels():
    """Identity-kernel convs + zero biases: out = shortcut + relu(bn2(relu(bn1(x)))); the first
    block of stage 1 skips the leading BN->ReLU (resnet.py:111-117)."""
    rng = np.random.RandomState(2)
    C = 4
    x = t64(rng.randn(1, 5, 3, C))
    eye = torch.zeros(3, 3, C, C, dtype=D)
    eye[1, 1] = torch.eye(C, dtype=D)
    p = {"b/conv1/kernel": eye, "b/conv1/bias": torch.zeros(C, dtype=D), "b/conv2/kernel": eye,
         "b/conv2/bias": torch.zeros(C, dtype=D)}
    for n in ("bn1", "bn2"):
        p["b/%s/gamma" % n] = t64(rng.rand(C) + 0.5)
        p["b/%s/beta" % n] = t64(rng.randn(C))
        p["b/%s/moving_mean" % n] = t64(rng.randn(C))
        p["b/%s/moving_variance" % n] = t64(rng.rand(C) + 0.5)
    got = O.basic_block(x, p, "b", C, 1, first_of_first=False)
    want = x + O.bn_relu(O.bn_relu(x, p, "b/bn1"), p, "b/bn2")
    assert torch.allclose(got, want, atol=1e-13)
    got1 = O.basic_block(x, p, "b", C, 1, first_of_first=True)
    assert torch.allclose(got1, x + O.bn_relu(x, p, "b/bn2"), atol=1e-13)


def test_projection_shortcut_taps_raw_input_with_valid_1x1():
    rng = np.random.RandomState(3)
    x = t64(rng.randn(1, 7, 5, 2))
    p = {"b/conv1/kernel": torch.zeros(3, 3, 2, 4, dtype=D), "b/conv1/bias": torch.zeros(4, dtype=D),
         "b/conv2/kernel": torch.zeros(3, 3, 4, 4, dtype=D), "b/conv2/bias": torch.zeros(4, dtype=D),
         "b/short/kernel": t64(rng.randn(1, 1, 2, 4)), "b/short/bias": t64(rng.randn(4))}
    for n, c in (("bn1", 2), ("bn2", 4)):
        p["b/%s/gamma" % n] = torch.ones(c, dtype=D); p["b/%s/beta" % n] = torch.zeros(c, dtype=D)
        p["b/%s/moving_mean" % n] = torch.zeros(c, dtype=D); p["b/%s/moving_variance" % n] = torch.ones(c, dtype=D)
    got = O.basic_block(x, p, "b", 4, 2, first_of_first=False)
    want = x[:, ::2, ::2, :] @ p["b/short/kernel"][0, 0] + p["b/short/bias"]      # raw x, stride 2, no BN
    assert tuple(got.shape) == (1, 4, 3, 4) and torch.allclose(got, want, atol=1e-13)


# ------------------------------------------------------------------ GRU
def _gru_weights(rng, din, u):
    return (t64(rng.randn(din, 3 * u) * 0.3), t64(rng.randn(u, 3 * u) * 0.3), t64(rng.randn(6 * u) * 0.3))


def test_gru_zero_weights_stay_zero_and_bias_only_closed_form():
    u, S = 3, 4
    x = torch.zeros(2, S, 5, dtype=D)
    out, h = O.gru_direction(x, torch.zeros(5, 3 * u, dtype=D), torch.zeros(u, 3 * u, dtype=D),
                             torch.zeros(6 * u, dtype=D), False)
    assert float(out.abs().max()) == 0.0
    bias = t64(np.random.RandomState(4).randn(6 * u))
    out, h = O.gru_direction(x, torch.zeros(5, 3 * u, dtype=D), torch.zeros(u, 3 * u, dtype=D), bias, False)
    bz = bias[:u] + bias[3 * u:4 * u]
    br = bias[u:2 * u] + bias[4 * u:5 * u]
    hh = torch.tanh(bias[2 * u:3 * u] + torch.sigmoid(br) * bias[5 * u:])
    z = torch.sigmoid(bz)
    href = torch.zeros(u, dtype=D)
    for _ in range(S):
        href = z * href + (1 - z) * hh
    assert torch.allclose(h[0], href, atol=1e-14)


@pytest.mark.parametrize("reverse", [False, True])
def test_gru_vs_torch_nn_gru_with_permuted_gates(reverse):
    """Keras/cuDNN gate order z|r|h vs torch r|z|n; same reset_after equations."""
    rng = np.random.RandomState(5)
    din, u, B, S = 6, 4, 3, 7
    k, r, b = _gru_weights(rng, din, u)
    x = t64(rng.randn(B, S, din))
    out, h = O.gru_direction(x, k, r, b, reverse)
    perm = torch.cat([torch.arange(u, 2 * u), torch.arange(0, u), torch.arange(2 * u, 3 * u)])
    g = torch.nn.GRU(din, u, batch_first=True).double()
    with torch.no_grad():
        g.weight_ih_l0.copy_(k[:, perm].T); g.weight_hh_l0.copy_(r[:, perm].T)
        g.bias_ih_l0.copy_(b[:3 * u][perm]); g.bias_hh_l0.copy_(b[3 * u:][perm])
        xin = x.flip(1) if reverse else x
        want, hn = g(xin)
    if reverse:
        want = want.flip(1)                                # Bidirectional re-reverses the backward outputs
    assert torch.allclose(out, want, atol=1e-12) and torch.allclose(h, hn[0], atol=1e-12)


def test_bigru_concat_and_last_state_semantics():
    rng = np.random.RandomState(6)
    din, u = 5, 3
    p = {}
    for d in ("forward", "backward"):
        k, r, b = _gru_weights(rng, din, u)
        p["g/%s/kernel" % d], p["g/%s/recurrent_kernel" % d], p["g/%s/bias" % d] = k, r, b
    x = t64(rng.randn(2, 6, din))
    seq = O.bigru(x, p, "g", seq=True)
    last = O.bigru(x, p, "g", seq=False)
    assert tuple(seq.shape) == (2, 6, 2 * u)
    assert torch.equal(last[:, :u], seq[:, -1, :u])        # forward final state = output at t = S-1
    assert torch.equal(last[:, u:], seq[:, 0, u:])         # backward final state = output at t = 0


# ------------------------------------------------------------------ VLAD
def test_vlad_literal_vs_closed_form_and_ghosts_last():
    rng = np.random.RandomState(7)
    B, S, Dd, K, G = 2, 9, 6, 4, 2
    feat = t64(rng.randn(B, 1, S, Dd)); score = t64(rng.randn(B, 1, S, K + G)); cen = t64(rng.randn(K + G, Dd))
    got = O.vlad_pooling(feat, score, cen, "gvlad", K)
    A = torch.softmax(score[:, 0], -1)
    V = A.transpose(1, 2) @ feat[:, 0] - A.sum(1).unsqueeze(-1) * cen           # A^T X - diag(sum A) c
    V = V[:, :K]
    V = V / torch.sqrt(torch.clamp((V * V).sum(-1, keepdim=True), min=1e-12))
    assert torch.allclose(got, V.reshape(B, K * Dd), atol=1e-13)
    # G = 0: 'gvlad' == 'vlad'
    a = O.vlad_pooling(feat, score[..., :K], cen[:K], "vlad", K)
    b = O.vlad_pooling(feat, score[..., :K], cen[:K], "gvlad", K)
    assert torch.equal(a, b)
    # constant scores -> uniform assignment
    u = O.vlad_pooling(feat, torch.zeros_like(score), cen, "gvlad", K)
    Vu = (feat[:, 0].sum(1, keepdim=True) / (K + G) - S / (K + G) * cen[None, :K])
    Vu = Vu / torch.sqrt((Vu * Vu).sum(-1, keepdim=True))
    assert torch.allclose(u, Vu.reshape(B, K * Dd), atol=1e-13)
    # zero residual stays zero under the 1e-12 clamp
    z = O.vlad_pooling(torch.zeros_like(feat), score, torch.zeros_like(cen), "gvlad", K)
    assert float(z.abs().max()) == 0.0


# ------------------------------------------------------------------ heads / losses
def test_face_heads_degenerate_identities():
    rng = np.random.RandomState(8)
    x, W = t64(rng.randn(5, 16)), t64(rng.randn(16, 4))
    y0 = torch.zeros(5, 4, dtype=D)
    y = t64(np.eye(4)[rng.randint(0, 4, 5)])
    cos = O.l2_normalize(x, 1) @ O.l2_normalize(W, 0)
    for kind in ("sphereface", "cosface", "arcface"):
        assert torch.allclose(O.face_logits(x, W, y0, kind, 0.3), 30 * cos, atol=1e-12)   # y = 0: plain s*cos
    a = O.face_logits(x, W, y, "arcface", 0.0)
    c = O.face_logits(x, W, y, "cosface", 0.0)
    s = O.face_logits(x, W, y, "sphereface", 1.0)
    assert torch.allclose(a, c, atol=1e-5) and torch.allclose(s, c, atol=1e-5)
    # clip: cos = 1 -> theta = acos(1 - 1e-7), not 0
    x1 = W[:, 0:1].T.clone()
    lg = O.face_logits(x1, W, t64([[1, 0, 0, 0]]), "arcface", 0.5)
    assert abs(float(lg[0, 0]) / 30 - math.cos(math.acos(1 - 1e-7) + 0.5)) < 1e-12


def test_circle_loss_two_class_hand_example():
    y = t64([[1.0, 0.0]]); s = t64([[0.8, 0.3]]); m, g = 0.25, 256.0
    lp = g * max(1 + m - 0.8, 0) * (0.8 - (1 - m))
    ln = g * max(0.3 + m, 0) * (0.3 - m)
    want = math.log(1 + math.exp(ln - lp))
    assert abs(float(O.circle_loss(y, s, g, m)[0]) - want) < 1e-12


def test_categorical_crossentropy_clip_path():
    y = t64([[0.0, 1.0], [1.0, 0.0]])
    p = t64([[1.0, 0.0], [2.0, 2.0]])                    # row 0: target prob 0 -> clip 1e-7; row 1 renormalised
    l = O.categorical_crossentropy(y, p)
    assert abs(float(l[0]) + math.log(1e-7)) < 1e-12 and abs(float(l[1]) + math.log(0.5)) < 1e-12


def test_loss_weights_double_assignment():
    """model.py:360-361: second assignment wins."""
    assert O.loss_weights(True, True, True, 0) == {"y_accent": 0.01, "y_disc": 0.6, "y_ctc_loss": 0.01}
    assert O.loss_weights(True, True, False, 0) == {"y_accent": 1.0, "y_ctc_loss": 0.6}
    assert O.loss_weights(False, True, True, 0) == {"y_accent": 0.01, "y_disc": 1.0}


# ------------------------------------------------------------------ CTC
def _ctc_brute(q, lab, blank):
    """-log sum over all alignments (path enumeration), q (T,C) probabilities."""
    T, C = q.shape
    tot = 0.0
    for path in itertools.product(range(C), repeat=T):
        col, prev = [], None
        for c in path:
            if c != prev and c != blank:
                col.append(c)
            prev = c
        if col == list(lab):
            pr = 1.0
            for t, c in enumerate(path):
                pr *= q[t, c]
            tot += pr
    return -math.log(tot)


@pytest.mark.parametrize("lab", [[0], [1, 2], [1, 1], [0, 1, 0], [2, 2, 1]])
def test_ctc_brute_force(lab):
    rng = np.random.RandomState(len(lab) * 3 + lab[0])
    T, C = 6, 4
    p = torch.softmax(t64(rng.randn(1, T, C)), -1)
    labels = t64([lab + [2] * (5 - len(lab))])           # EOS(2)-padded like utils.py:57-63
    got = float(O.ctc_batch_cost(labels, p, np.array([T]), np.array([len(lab)]))[0, 0])
    q = (p[0] + 1e-7) / (p[0] + 1e-7).sum(-1, keepdim=True)     # K.ctc_batch_cost double normalisation
    assert abs(got - _ctc_brute(q.numpy(), lab, C - 1)) < 1e-10


def test_ctc_vs_torch_ctc_loss_and_renormalisation_shift():
    rng = np.random.RandomState(9)
    B, S, C, L = 3, 30, 50, 8
    p = torch.softmax(t64(rng.randn(B, S, C)), -1)
    labels = t64(rng.randint(0, C - 1, (B, L)))
    lens = np.array([8, 5, 1])
    got = O.ctc_batch_cost(labels, p, np.full(B, S), lens)[:, 0]
    logq = torch.log_softmax(torch.log(p + 1e-7), -1)
    want = F.ctc_loss(logq.transpose(0, 1), labels.long(), torch.full((B,), S), torch.as_tensor(lens), blank=C - 1,
                      reduction="none")
    assert torch.allclose(got, want, atol=1e-9)
    plain = F.ctc_loss(torch.log(p).transpose(0, 1), labels.long(), torch.full((B,), S), torch.as_tensor(lens),
                       blank=C - 1, reduction="none")
    assert float((got - plain).abs().max()) > 1e-6        # the (p+eps)/Z re-normalisation is visible


def test_ctc_infeasible_raises():
    p = torch.full((1, 3, 4), 0.25, dtype=D)
    with pytest.raises(ValueError):
        O.ctc_batch_cost(t64([[1, 1, 1]]), p, np.array([3]), np.array([3]))


# ------------------------------------------------------------------ whole forward
def test_forward_output_contract_and_order():
    from aesrc2020_b200 import weights as W, utils as us
    from aesrc2020_b200.config import SARConfig
    cfg = SARConfig(input_shape=(64, 80, 1), ctc_enable=True, disc_enable=True, res_type="res18", res_filters=8,
                    mto="gvlad", vlad_clusters=4, ghost_clusters=2, metric_loss="cosface", bn_dim=16)
    w = W.init_weights(cfg)
    x, y = us.synthetic_batch(cfg, 2, seed=1, label_len_range=(1, 2))
    out = O.sar_net_forward(w, x, **cfg.model_kwargs())
    assert cfg.output_names() == ["y_accent", "y_disc", "y_ctc_loss", "y_disc_bn"]
    assert tuple(out["y_accent"].shape) == (2, 8) and tuple(out["y_ctc_loss"].shape) == (2, 1)
    assert tuple(out["y_disc_bn"].shape) == (2, 8)
    assert torch.allclose(out["y_accent"].sum(-1), torch.ones(2, dtype=D))
    with pytest.raises(SystemExit):
        O.sar_net_forward(w, x, **{**cfg.model_kwargs(), "mto": None})       # model.py:136-138


def test_ctc_greedy_decode_known_answers():
    """ctc_pred (model.py:385-389): merge repeats THEN drop blanks (blank = C-1), first maximum on ties, frames beyond
    input_len ignored, -1 padding."""
    C = 5                                    # blank = 4
    path = [[1, 1, 4, 1, 2, 2, 4, 4, 3], [4, 4, 4, 4, 4, 4, 4, 4, 4], [0, 4, 0, 0, 3, 3, 3, 1, 4]]
    pr = np.full((3, 9, C), 0.1)
    for b, row in enumerate(path):
        for t, c in enumerate(row):
            pr[b, t, c] = 0.6
    pr[2, 1, :] = 0.2                        # a tie over all classes -> class 0 (first maximum), merges with the 0 before
    got = O.ctc_greedy_decode(pr, 9)
    assert got.tolist() == [[1, 1, 2, 3], [-1, -1, -1, -1], [0, 3, 1, -1]]
    assert O.ctc_greedy_decode(pr, 4).tolist() == [[1, 1], [-1, -1], [0, -1]]
