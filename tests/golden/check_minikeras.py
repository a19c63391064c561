"""Tie the fixtures to an INDEPENDENT implementation of the third-party primitives.

`minikeras` restates the Keras/TF primitives the reference calls (Conv2D 'same'/'valid', BatchNormalization,
MaxPooling2D 'same', Dense, CuDNNGRU / Bidirectional, LayerNormalization, K.ctc_batch_cost, K.l2_normalize) in numpy, and
the oracle restates them again -- both by the same author.  `assert_primitives_match_torch()` checks every one of them
against PyTorch's own kernels (F.conv2d with the TF-SAME pads made explicit, F.batch_norm, F.max_pool2d, nn.GRU with the
gates permuted from Keras' z|r|h to torch's r|z|n, F.layer_norm(eps=1e-14), F.ctc_loss(blank=C-1), F.normalize) in float64.
It runs INSIDE make_golden.py before any fixture is written, and as a CPU test (tests/test_golden.py), so a drift of the
stand-in from the independent implementation fails the generator and the suite.

What torch cannot vouch for is only which of its options Keras 2.2 / TF 1.13 correspond to (SAME's extra pad goes to the
bottom/right; CuDNNGRU is the reset_after form with two bias sets; ctc_batch_cost feeds log(p + 1e-7) to a loss that
softmaxes again) -- those choices are the [KERAS-SEMANTICS] constants documented in the oracle header.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)
import minikeras as mk                                   # noqa: E402

TOL = 1e-10


def _close(a, b, what, tol=TOL):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = float(np.max(np.abs(a - b)) / max(1.0, float(np.max(np.abs(b))))) if a.size else 0.0
    assert err < tol, "%s: minikeras differs from torch by %.3e" % (what, err)
    return err


def _t(a):
    return torch.as_tensor(np.asarray(a), dtype=torch.float64)


def check_conv2d(rng):
    worst = 0.0
    for (H, W, Cin, Cout, k, s, pad) in ((13, 9, 3, 5, 3, 1, "same"), (13, 9, 3, 5, 3, 2, "same"), (12, 10, 2, 4, 7, 2, "same"),
                                         (8, 5, 4, 6, 1, 2, "valid"), (9, 6, 4, 6, 1, 1, "valid"), (125, 20, 2, 3, 3, 2, "same"),
                                         (16, 3, 2, 2, 3, 1, "same")):
        mk.reset(seed=1)
        lay = mk.Conv2D(Cout, (k, k), strides=(s, s), padding=pad)
        x = rng.randn(2, H, W, Cin)
        got = lay(mk.KT(x)).v
        kern, bias = lay.weights["kernel"].v, lay.weights["bias"].v
        xt = _t(x).permute(0, 3, 1, 2)
        if pad == "same":                                   # TF SAME: total = max((ceil(n/s)-1)*s + k - n, 0), extra at the END
            def pads(n):
                out = -(-n // s)
                tot = max((out - 1) * s + k - n, 0)
                return tot // 2, tot - tot // 2
            (pt, pb), (pl, pr) = pads(H), pads(W)
            xt = F.pad(xt, (pl, pr, pt, pb))
        want = F.conv2d(xt, _t(kern).permute(3, 2, 0, 1), _t(bias), stride=s).permute(0, 2, 3, 1).numpy()
        worst = max(worst, _close(got, want, "Conv2D %s k%d s%d" % (pad, k, s)))
    return worst


def check_bn_pool_dense(rng):
    mk.reset(seed=2)
    x = rng.randn(3, 7, 5, 6)
    bn = mk.BatchNormalization()
    got = bn(mk.KT(x)).v
    w = {k: _t(v.v) for k, v in bn.weights.items()}
    want = F.batch_norm(_t(x).permute(0, 3, 1, 2), w["moving_mean"], w["moving_variance"], w["gamma"], w["beta"],
                        training=False, eps=1e-3).permute(0, 2, 3, 1).numpy()
    e1 = _close(got, want, "BatchNormalization(eps=1e-3, inference)")
    worst = e1
    for (H, W) in ((250, 40), (13, 9), (12, 10)):
        x = rng.randn(2, H, W, 3)
        got = mk.MaxPooling2D((3, 3), strides=(2, 2), padding="same")(mk.KT(x)).v
        Ho, Wo = -(-H // 2), -(-W // 2)
        th, tw = max((Ho - 1) * 2 + 3 - H, 0), max((Wo - 1) * 2 + 3 - W, 0)
        xt = F.pad(_t(x).permute(0, 3, 1, 2), (tw // 2, tw - tw // 2, th // 2, th - th // 2), value=float("-inf"))
        want = F.max_pool2d(xt, 3, 2).permute(0, 2, 3, 1).numpy()
        worst = max(worst, _close(got, want, "MaxPooling2D same %dx%d" % (H, W)))
    for act in (None, "relu", "tanh", "softmax"):
        mk.reset(seed=3)
        lay = mk.Dense(7, activation=act)
        x = rng.randn(4, 5, 6)
        got = lay(mk.KT(x)).v
        y = _t(x) @ _t(lay.weights["kernel"].v) + _t(lay.weights["bias"].v)
        want = {None: y, "relu": torch.relu(y), "tanh": torch.tanh(y), "softmax": torch.softmax(y, -1)}[act].numpy()
        worst = max(worst, _close(got, want, "Dense(%s)" % act))
    return worst


def _torch_gru_from_keras(W, U, b, u, din):
    """Keras CuDNNGRU weights (gate order z|r|h, bias = [input | recurrent]) -> nn.GRU (gate order r|z|n)."""
    g = torch.nn.GRU(din, u, batch_first=True).double()
    perm = np.concatenate([np.arange(u, 2 * u), np.arange(0, u), np.arange(2 * u, 3 * u)])
    with torch.no_grad():
        g.weight_ih_l0.copy_(_t(W.T[perm]))
        g.weight_hh_l0.copy_(_t(U.T[perm]))
        g.bias_ih_l0.copy_(_t(b[:3 * u][perm]))
        g.bias_hh_l0.copy_(_t(b[3 * u:][perm]))
    return g


def check_gru_ln(rng):
    worst = 0.0
    for seq in (True, False):
        mk.reset(seed=4)
        u, din, B, S = 6, 5, 3, 9
        bi = mk.Bidirectional(mk.CuDNNGRU(u, return_sequences=seq), merge_mode="concat", name="g")
        x = rng.randn(B, S, din)
        got = bi(mk.KT(x)).v
        outs = []
        for d, rev in (("forward", False), ("backward", True)):
            W, U, b = (bi.weights["%s/%s" % (d, k)].v for k in ("kernel", "recurrent_kernel", "bias"))
            g = _torch_gru_from_keras(W, U, b, u, din)
            xin = torch.flip(_t(x), [1]) if rev else _t(x)
            with torch.no_grad():
                o, h = g(xin)
            o = torch.flip(o, [1]) if rev else o          # Keras Bidirectional re-reverses the backward sequence
            outs.append(o if seq else h[0])
        want = torch.cat(outs, -1).numpy()
        worst = max(worst, _close(got, want, "Bidirectional(CuDNNGRU, return_sequences=%s)" % seq))
    mk.reset(seed=5)
    ln = mk.LayerNormalization()
    x = rng.randn(4, 7, 16) * 3 + 1
    got = ln(mk.KT(x)).v
    want = F.layer_norm(_t(x), (16,), _t(ln.weights["gamma"].v), _t(ln.weights["beta"].v), eps=1e-14).numpy()
    worst = max(worst, _close(got, want, "LayerNormalization(eps=1e-14)"))
    return worst


def check_ctc_l2n(rng):
    B, S, C, Lmax = 4, 12, 9, 5
    logits = rng.randn(B, S, C)
    p = np.exp(logits) / np.exp(logits).sum(-1, keepdims=True)
    lab_len = np.array([3, 5, 1, 4])
    in_len = np.array([12, 12, 7, 10])
    labels = rng.randint(0, C - 1, size=(B, Lmax))
    labels[1] = [2, 2, 3, 3, 2]                          # repeats: forced blanks
    got = mk._ctc_batch_cost(mk.KT(labels.astype(np.float64)), mk.KT(p), mk.KT(in_len.reshape(B, 1)), mk.KT(lab_len.reshape(B, 1))).v
    lg = torch.log_softmax(torch.log(_t(p) + 1e-7), -1)   # K.ctc_batch_cost: log(p + eps), then tf.nn.ctc_loss softmaxes
    want = F.ctc_loss(lg.permute(1, 0, 2), torch.as_tensor(labels), torch.as_tensor(in_len), torch.as_tensor(lab_len),
                      blank=C - 1, reduction="none", zero_infinity=False).numpy().reshape(B, 1)
    worst = _close(got, want, "K.ctc_batch_cost (blank = C-1)", tol=1e-9)
    x = rng.randn(5, 8)
    worst = max(worst, _close(mk._l2n(mk.KT(x), -1).v, F.normalize(_t(x), dim=-1, eps=1e-6).numpy(), "K.l2_normalize"))
    return worst


def assert_primitives_match_torch(verbose=False):
    rng = np.random.RandomState(2020)
    res = {"conv2d": check_conv2d(rng), "bn/pool/dense": check_bn_pool_dense(rng), "gru/ln": check_gru_ln(rng),
           "ctc/l2n": check_ctc_l2n(rng)}
    mk.reset()
    if verbose:
        print("minikeras primitives vs torch (max normalised error):", {k: "%.1e" % v for k, v in res.items()})
    return res


if __name__ == "__main__":
    assert_primitives_match_torch(verbose=True)
