"""Pins for oracle/train_oracle.py (SURVEY 8f-1 first slice): autograd gradients vs central finite differences for every
head kind, Keras-style Adam vs torch.optim.Adam driven with the same decayed learning rate, BatchNorm training mode vs
F.batch_norm, and one whole step reproducing a hand-rolled float64 computation."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import train_oracle as TO, sarnet_oracle as O

KINDS = [("arcface", 0.3), ("cosface", 0.3), ("sphereface", 1.35), ("softmax", 0.0), ("circleloss", 0.2)]


def _params(kind, D=24, E=16, n=8, seed=0):
    rng = np.random.RandomState(seed)
    p = {"AR_BN1/gamma": rng.uniform(0.7, 1.3, D), "AR_BN1/beta": rng.randn(D) * 0.1, "AR_BN1/moving_mean": rng.randn(D) * 0.1,
         "AR_BN1/moving_variance": rng.uniform(0.6, 1.4, D),
         "AR_EMBEDDING/kernel": rng.randn(D, E) / np.sqrt(D), "AR_EMBEDDING/bias": rng.randn(E) * 0.1,
         "AR_BN2/gamma": rng.uniform(0.7, 1.3, E), "AR_BN2/beta": rng.randn(E) * 0.1, "AR_BN2/moving_mean": rng.randn(E) * 0.1,
         "AR_BN2/moving_variance": rng.uniform(0.6, 1.4, E),
         "AR_CF_DS1/kernel": rng.randn(E, 12) / 4, "AR_CF_DS1/bias": rng.randn(12) * 0.1,
         "AR_CF_DS2/kernel": rng.randn(12, 12) / 3, "AR_CF_DS2/bias": rng.randn(12) * 0.1,
         "y_accent/kernel": rng.randn(12, n) / 3, "y_accent/bias": rng.randn(n) * 0.1}
    W = rng.uniform(-0.4, 0.4, (E, n))
    if kind == "circleloss":
        W = W / np.sqrt((W ** 2).sum(0, keepdims=True))
    p[TO.disc_key(kind)] = W
    return {k: np.asarray(v, np.float64) for k, v in p.items()}


@pytest.mark.parametrize("kind,margin", KINDS)
def test_autograd_gradients_match_finite_differences(kind, margin):
    B, n = 6, 8
    rng = np.random.RandomState(3)
    params = _params(kind)
    integ = rng.randn(B, 24)
    onehot = np.eye(n)[rng.randint(0, n, B)]
    kw = dict(disc_enable=True, metric_loss=kind, margin=margin, w_accent=0.01, w_disc=0.6)
    _, _, _, grads = TO.train_step(params, {}, integ, onehot, lr=0.01, iterations=0, **kw)

    def loss_of(pp):
        t = {k: torch.as_tensor(v) for k, v in pp.items()}
        return float(TO.head_loss(t, torch.as_tensor(integ), torch.as_tensor(onehot), **kw)[0])
    for k in TO.trainable_keys(True, kind):
        g = grads[k]
        idx = [tuple(rng.randint(0, s) for s in g.shape) for _ in range(4)]
        for i in idx:
            h = 1e-6
            pp, pm = {q: v.copy() for q, v in params.items()}, {q: v.copy() for q, v in params.items()}
            pp[k][i] += h; pm[k][i] -= h
            fd = (loss_of(pp) - loss_of(pm)) / (2 * h)
            assert abs(fd - g[i]) <= 1e-6 * max(1.0, abs(fd)) + 2e-8, (kind, k, i, fd, g[i])


def test_adam_matches_torch_optim_with_keras_decay():
    rng = np.random.RandomState(1)
    p0 = rng.randn(5, 3)
    pt = torch.tensor(p0.copy(), requires_grad=True)
    opt = torch.optim.Adam([pt], lr=0.01, betas=(0.9, 0.999), eps=0.0)       # eps handled below
    p, m, v = torch.as_tensor(p0.copy()), torch.zeros(5, 3, dtype=torch.float64), torch.zeros(5, 3, dtype=torch.float64)
    for it in range(6):
        g = torch.as_tensor(rng.randn(5, 3))
        p, m, v = TO.adam_update(p, g, m, v, it, lr=0.01)
        # torch: p -= lr * mhat / (sqrt(vhat) + eps) with mhat = m/(1-b1^t), vhat = v/(1-b2^t): identical to Keras' form
        # when eps = 0; with Keras' eps OUTSIDE the sqrt of the UNcorrected v the two differ at the 1e-7 level
        for grp in opt.param_groups:
            grp["lr"] = 0.01 / (1 + TO.ADAM_DECAY * it)
        pt.grad = g.clone()
        opt.step()
        assert float((p - pt.detach()).abs().max()) < 2e-6
    assert float((p - torch.as_tensor(p0)).abs().max()) > 1e-2                # it moved


def test_bn_training_mode_matches_torch():
    rng = np.random.RandomState(2)
    x = torch.as_tensor(rng.randn(9, 7) * 2 + 1)
    g, b = torch.as_tensor(rng.uniform(0.5, 1.5, 7)), torch.as_tensor(rng.randn(7))
    y, mean, var = TO.bn_train(x, g, b)
    rm, rv = torch.zeros(7, dtype=torch.float64), torch.ones(7, dtype=torch.float64)
    want = F.batch_norm(x, rm, rv, g, b, training=True, momentum=0.01, eps=TO.BN_EPS)
    assert float((y - want).abs().max()) < 1e-12
    assert float((rm - 0.01 * mean).abs().max()) < 1e-12       # torch's running mean moved by (1 - 0.99) * batch mean
    # (torch updates the running variance with the UNBIASED estimate; Keras' non-fused path uses the biased one)
    assert float((rv - (0.99 + 0.01 * var * 9 / 8)).abs().max()) < 1e-12


@pytest.mark.parametrize("kind,margin", [("arcface", 0.3), ("circleloss", 0.2)])
def test_train_step_reduces_the_loss_and_updates_state(kind, margin):
    rng = np.random.RandomState(5)
    params = _params(kind, seed=4)
    B, n = 16, 8
    lab = rng.randint(0, n, B)
    integ = rng.randn(B, 24) + np.eye(n)[lab] @ rng.randn(n, 24)           # separable classes
    onehot = np.eye(n)[lab]
    kw = dict(disc_enable=True, metric_loss=kind, margin=margin, w_accent=0.01, w_disc=1.0)
    state, losses = {}, []
    p = params
    for it in range(40):
        p, state, l, _ = TO.train_step(p, state, integ, onehot, lr=0.02, iterations=it, **kw)
        losses.append(l["total"])
    assert losses[-1] < 0.6 * losses[0]
    assert set(state) == {s + k for k in TO.trainable_keys(True, kind) for s in ("m/", "v/")}
    assert not np.allclose(p["AR_BN1/moving_mean"], params["AR_BN1/moving_mean"])
    if kind == "circleloss":
        assert np.allclose(np.sqrt((p["y_disc/kernel"] ** 2).sum(0)), 1.0, atol=1e-6)    # unit_norm constraint


@pytest.mark.parametrize("mto,G", [("gvlad", 2), ("vlad", 0)])
def test_pooled_gradients_match_finite_differences(mto, G):
    """Second slice: vlad() (model.py:82-109, VLAD.py:26-49) in front of the head -- autograd gradients of the assignment
    Conv2D, its bias and the centers vs central finite differences; ghost centers get exactly zero."""
    B, S, Dd, K, n = 5, 7, 6, 4, 8
    rng = np.random.RandomState(12)
    params = _params("arcface", D=K * Dd)
    params[mto + "_center_assignment/kernel"] = rng.randn(1, 1, Dd, K + G) * 0.5
    params[mto + "_center_assignment/bias"] = rng.randn(K + G) * 0.1
    params[mto + "_pool/centers"] = rng.randn(K + G, Dd) * 0.5
    feat = rng.randn(B, S, Dd)
    onehot = np.eye(n)[rng.randint(0, n, B)]
    kw = dict(disc_enable=True, metric_loss="arcface", margin=0.3, w_accent=0.01, w_disc=0.6)
    pool = dict(mto=mto, vlad_clusters=K, ghost_clusters=G)
    _, state, losses, grads = TO.train_step(params, {}, feat, onehot, lr=0.01, iterations=0, pool=pool, **kw)
    assert set(TO.pool_keys(mto)) <= set(grads) and "m/" + mto + "_pool/centers" in state

    def loss_of(pp):
        t = {k: torch.as_tensor(v) for k, v in pp.items()}
        return float(TO.pooled_head_loss(t, torch.as_tensor(feat), torch.as_tensor(onehot), **pool, **kw)[0])
    assert abs(loss_of(params) - losses["total"]) < 1e-12
    for k in TO.pool_keys(mto) + ["AR_EMBEDDING/kernel", "AR_BN1/gamma"]:
        g = grads[k]
        for _ in range(5):
            i = tuple(rng.randint(0, s_) for s_ in g.shape)
            h = 1e-6
            pp, pm = {q: v.copy() for q, v in params.items()}, {q: v.copy() for q, v in params.items()}
            pp[k][i] += h; pm[k][i] -= h
            fd = (loss_of(pp) - loss_of(pm)) / (2 * h)
            assert abs(fd - g[i]) <= 2e-6 * max(1.0, abs(fd)) + 2e-8, (k, i, fd, g[i])
    if G:
        assert np.all(grads[mto + "_pool/centers"][K:] == 0.0)                 # VLAD.py:44-45 drops the ghost rows


def test_third_slice_gradients_match_finite_differences():
    """Third slice: AR_DS (Dense + tanh) -> AR_DS_LN -> vlad() -> head on the CRNN_LN output: autograd vs finite differences
    for the Dense kernel / bias and the LayerNormalization gamma / beta."""
    B, S, C0, Dd, K, G, n = 4, 5, 10, 6, 4, 2, 8
    rng = np.random.RandomState(17)
    params = _params("arcface", D=K * Dd)
    params["gvlad_center_assignment/kernel"] = rng.randn(1, 1, Dd, K + G) * 0.5
    params["gvlad_center_assignment/bias"] = rng.randn(K + G) * 0.1
    params["gvlad_pool/centers"] = rng.randn(K + G, Dd) * 0.5
    params["AR_DS/kernel"] = rng.randn(C0, Dd) * 0.4
    params["AR_DS/bias"] = rng.randn(Dd) * 0.1
    params["AR_DS_LN/gamma"] = rng.uniform(0.7, 1.3, Dd)
    params["AR_DS_LN/beta"] = rng.randn(Dd) * 0.1
    x = rng.randn(B, S, C0)
    onehot = np.eye(n)[rng.randint(0, n, B)]
    kw = dict(disc_enable=True, metric_loss="arcface", margin=0.3, w_accent=0.01, w_disc=0.6)
    pool = dict(mto="gvlad", vlad_clusters=K, ghost_clusters=G, train_ds=True)
    _, state, losses, grads = TO.train_step(params, {}, x, onehot, lr=0.01, iterations=0, pool=pool, **kw)
    assert set(TO.DS_KEYS) <= set(grads)

    def loss_of(pp):
        t = {k: torch.as_tensor(v) for k, v in pp.items()}
        return float(TO.pooled_head_loss(t, torch.as_tensor(x), torch.as_tensor(onehot), **pool, **kw)[0])
    for k in TO.DS_KEYS + ["gvlad_pool/centers"]:
        g = grads[k]
        for _ in range(5):
            i = tuple(rng.randint(0, s_) for s_ in g.shape)
            if k == "gvlad_pool/centers":
                i = (rng.randint(0, K),) + i[1:]
            h = 1e-6
            pp, pm = {q: v.copy() for q, v in params.items()}, {q: v.copy() for q, v in params.items()}
            pp[k][i] += h; pm[k][i] -= h
            fd = (loss_of(pp) - loss_of(pm)) / (2 * h)
            assert abs(fd - g[i]) <= 2e-6 * max(1.0, abs(fd)) + 2e-8, (k, i, fd, g[i])


def _crnn_params(rng, params, Cc, u):
    params["CNN_LIN/kernel"] = rng.randn(Cc, u) * 0.4
    params["CNN_LIN/bias"] = rng.randn(u) * 0.1
    params["CNN_LIN_LN/gamma"] = rng.uniform(0.7, 1.3, u)
    params["CNN_LIN_LN/beta"] = rng.randn(u) * 0.1
    for d in ("forward", "backward"):
        params["CRNN/%s/kernel" % d] = rng.randn(u, 3 * u) * 0.4
        params["CRNN/%s/recurrent_kernel" % d] = rng.randn(u, 3 * u) * 0.4
        params["CRNN/%s/bias" % d] = rng.randn(6 * u) * 0.1
    params["CRNN_LN/gamma"] = rng.uniform(0.7, 1.3, 2 * u)
    params["CRNN_LN/beta"] = rng.randn(2 * u) * 0.1
    return params


def test_fourth_slice_gradients_match_finite_differences():
    """Fourth slice: CNN_LIN (Dense + tanh) -> CNN_LIN_LN -> CRNN (Bidirectional CuDNNGRU: back-propagation through time in
    both directions) -> CRNN_LN -> the accent branch, on the frozen ResNet's sequence: autograd vs central differences for
    every new trainable tensor (kernel, recurrent kernel and the 6u bias of both directions, the LayerNormalizations)."""
    B, S, Cc, u, Dd, K, G, n = 3, 6, 7, 5, 6, 4, 2, 8
    rng = np.random.RandomState(23)
    params = _params("arcface", D=K * Dd)
    params["gvlad_center_assignment/kernel"] = rng.randn(1, 1, Dd, K + G) * 0.5
    params["gvlad_center_assignment/bias"] = rng.randn(K + G) * 0.1
    params["gvlad_pool/centers"] = rng.randn(K + G, Dd) * 0.5
    params["AR_DS/kernel"] = rng.randn(2 * u, Dd) * 0.4
    params["AR_DS/bias"] = rng.randn(Dd) * 0.1
    params["AR_DS_LN/gamma"] = rng.uniform(0.7, 1.3, Dd)
    params["AR_DS_LN/beta"] = rng.randn(Dd) * 0.1
    _crnn_params(rng, params, Cc, u)
    x = rng.randn(B, S, Cc)
    onehot = np.eye(n)[rng.randint(0, n, B)]
    kw = dict(disc_enable=True, metric_loss="arcface", margin=0.3, w_accent=0.01, w_disc=0.6)
    pool = dict(mto="gvlad", vlad_clusters=K, ghost_clusters=G, train_crnn=True)
    _, state, losses, grads = TO.train_step(params, {}, x, onehot, lr=0.01, iterations=0, pool=pool, **kw)
    assert set(TO.CRNN_KEYS) | set(TO.DS_KEYS) <= set(grads)

    def loss_of(pp):
        t = {k: torch.as_tensor(v) for k, v in pp.items()}
        return float(TO.pooled_head_loss(t, torch.as_tensor(x), torch.as_tensor(onehot), **pool, **kw)[0])
    for k in TO.CRNN_KEYS + ["AR_DS/kernel"]:
        g = grads[k]
        for _ in range(4):
            i = tuple(rng.randint(0, s_) for s_ in g.shape)
            h = 1e-6
            pp, pm = {q: v.copy() for q, v in params.items()}, {q: v.copy() for q, v in params.items()}
            pp[k][i] += h; pm[k][i] -= h
            fd = (loss_of(pp) - loss_of(pm)) / (2 * h)
            assert abs(fd - g[i]) <= 2e-6 * max(1.0, abs(fd)) + 2e-8, (k, i, fd, g[i])
    # the regularisers: l2 on the Dense and on the GRU's input kernel / bias, none on the recurrent kernels
    assert all(k in TO.CRNN_L2_KEYS for k in ("CRNN/forward/kernel", "CRNN/backward/bias", "CNN_LIN/bias"))
    assert not any(k.endswith("recurrent_kernel") for k in TO.CRNN_L2_KEYS)


def test_fifth_slice_ctc_branch_gradients_match_finite_differences():
    """Fifth slice: the CTC branch (CTC_BIGRU -> LN -> CTC_DS -> LN -> ctc_pred -> K.ctc_batch_cost) trained with the accent
    branch above the frozen ResNet.  The differentiable CTC (torch's own on log q) equals the oracle's explicit lattice;
    autograd of the multi-task total vs central differences for every CTC-branch tensor and for the shared CRNN below."""
    from oracle import sarnet_oracle as O
    B, S, Cc, u, Dd, K, G, n, C, Lmax = 3, 7, 6, 4, 6, 4, 2, 8, 9, 3
    rng = np.random.RandomState(29)
    params = _params("arcface", D=K * Dd)
    params["gvlad_center_assignment/kernel"] = rng.randn(1, 1, Dd, K + G) * 0.5
    params["gvlad_center_assignment/bias"] = rng.randn(K + G) * 0.1
    params["gvlad_pool/centers"] = rng.randn(K + G, Dd) * 0.5
    params["AR_DS/kernel"] = rng.randn(2 * u, Dd) * 0.4
    params["AR_DS/bias"] = rng.randn(Dd) * 0.1
    params["AR_DS_LN/gamma"] = rng.uniform(0.7, 1.3, Dd)
    params["AR_DS_LN/beta"] = rng.randn(Dd) * 0.1
    _crnn_params(rng, params, Cc, u)
    for d in ("forward", "backward"):
        params["CTC_BIGRU/%s/kernel" % d] = rng.randn(2 * u, 3 * u) * 0.4
        params["CTC_BIGRU/%s/recurrent_kernel" % d] = rng.randn(u, 3 * u) * 0.4
        params["CTC_BIGRU/%s/bias" % d] = rng.randn(6 * u) * 0.1
    params["CTC_BIGRU_LN/gamma"] = rng.uniform(0.7, 1.3, 2 * u)
    params["CTC_BIGRU_LN/beta"] = rng.randn(2 * u) * 0.1
    params["CTC_DS/kernel"] = rng.randn(2 * u, u) * 0.4
    params["CTC_DS/bias"] = rng.randn(u) * 0.1
    params["CTC_DS_LN/gamma"] = rng.uniform(0.7, 1.3, u)
    params["CTC_DS_LN/beta"] = rng.randn(u) * 0.1
    params["ctc_pred/kernel"] = rng.randn(u, C) * 0.6
    params["ctc_pred/bias"] = rng.randn(C) * 0.1
    x = rng.randn(B, S, Cc)
    onehot = np.eye(n)[rng.randint(0, n, B)]
    labels = np.array([[1, 1, 4], [0, 7, 0], [5, 0, 0]], np.float64)          # a repeated label, ragged lengths
    in_len, lab_len = np.array([[7], [6], [4]]), np.array([[3], [2], [1]])
    # the differentiable CTC is the oracle's CTC
    pr = torch.softmax(torch.as_tensor(rng.randn(B, S, C)), -1)
    a = TO.ctc_loss_autograd(pr, labels, in_len, lab_len)
    b = O.ctc_batch_cost(torch.as_tensor(labels), pr, torch.as_tensor(in_len), torch.as_tensor(lab_len)).reshape(-1)
    assert torch.allclose(a, b, rtol=1e-10, atol=1e-10)
    kw = dict(disc_enable=True, metric_loss="arcface", margin=0.3, w_accent=0.01, w_disc=0.6)
    pool = dict(mto="gvlad", vlad_clusters=K, ghost_clusters=G, train_ctc=True, ctc=(labels, in_len, lab_len), w_ctc=0.3)
    _, state, losses, grads = TO.train_step(params, {}, x, onehot, lr=0.01, iterations=0, pool=pool, **kw)
    assert set(TO.CTC_KEYS) | set(TO.CRNN_KEYS) <= set(grads) and losses["loss_ctc"] > 0

    def loss_of(pp):
        t = {k: torch.as_tensor(v) for k, v in pp.items()}
        return float(TO.pooled_head_loss(t, torch.as_tensor(x), torch.as_tensor(onehot), **pool, **kw)[0])
    for k in TO.CTC_KEYS + ["CRNN/forward/recurrent_kernel", "CRNN_LN/gamma", "CNN_LIN/kernel"]:
        g = grads[k]
        for _ in range(3):
            i = tuple(rng.randint(0, s_) for s_ in g.shape)
            h = 1e-6
            pp, pm = {q: v.copy() for q, v in params.items()}, {q: v.copy() for q, v in params.items()}
            pp[k][i] += h; pm[k][i] -= h
            fd = (loss_of(pp) - loss_of(pm)) / (2 * h)
            assert abs(fd - g[i]) <= 2e-6 * max(1.0, abs(fd)) + 2e-8, (k, i, fd, g[i])


def test_sixth_slice_resnet_training_gradients_match_finite_differences():
    """Sixth slice: the ResNet in training mode (batch-statistic BN) under the rest of the model: autograd of the total vs
    central differences for conv kernels / biases / BN parameters of the stem, a plain block, a strided block with its
    projection shortcut, and the final BN."""
    from aesrc2020_b200.config import SARConfig
    from aesrc2020_b200 import weights as W
    from aesrc2020_b200.training_resnet import ResNetTrainer
    K, G, n = 4, 2, 8
    cfg = SARConfig(input_shape=(40, 80, 1), ctc_enable=False, ar_enable=True, disc_enable=True, res_type="res18", res_filters=4,
                    hidden_dim=6, mto="gvlad", vlad_clusters=K, ghost_clusters=G, metric_loss="arcface", margin=0.3)
    w = W.init_weights(cfg, 3)
    rng = np.random.RandomState(31)
    params = {k: np.asarray(v, np.float64) for k, v in w.items()}
    for k in params:                                  # biases / betas away from zero so that every term is exercised
        if k.endswith("/bias") or k.endswith("/beta"):
            params[k] = params[k] + rng.randn(*params[k].shape) * 0.1
    rkeys, rl2, _ = ResNetTrainer.param_keys(cfg)
    B = 3
    x = rng.rand(B, 40, 80, 1)
    onehot = np.eye(n)[rng.randint(0, n, B)]
    kw = dict(disc_enable=True, metric_loss="arcface", margin=0.3, w_accent=0.01, w_disc=0.6)
    pool = dict(mto="gvlad", vlad_clusters=K, ghost_clusters=G, train_resnet=dict(res_type="res18", filters=4),
                extra_keys=tuple(rkeys), extra_l2=tuple(rl2))
    _, state, losses, grads = TO.train_step(params, {}, x, onehot, lr=0.01, iterations=0, pool=pool, **kw)
    assert set(rkeys) <= set(grads)

    def loss_of(pp):
        t = {k: torch.as_tensor(v) for k, v in pp.items()}
        return float(TO.pooled_head_loss(t, torch.as_tensor(x), torch.as_tensor(onehot), **pool, **kw)[0])
    probe = ["resnet/stem/kernel", "resnet/stem_bn/gamma", "resnet/s1b2/conv1/kernel", "resnet/s1b2/bn1/beta", "resnet/s2b1/conv1/kernel",
             "resnet/s2b1/short/kernel", "resnet/s2b1/short/bias", "resnet/s4b2/conv2/bias", "resnet/final_bn/gamma", "CNN_LIN/kernel"]
    for k in probe:
        g = grads[k]
        for _ in range(3):
            i = tuple(rng.randint(0, s_) for s_ in g.shape)
            h = 1e-6
            pp, pm = {q: v.copy() for q, v in params.items()}, {q: v.copy() for q, v in params.items()}
            pp[k][i] += h; pm[k][i] -= h
            fd = (loss_of(pp) - loss_of(pm)) / (2 * h)
            assert abs(fd - g[i]) <= 5e-6 * max(1.0, abs(fd)) + 5e-8, (k, i, fd, g[i])


@pytest.mark.parametrize("mto", ["bigru", "avg"])
def test_bigru_and_avg_integration_gradients_match_finite_differences(mto):
    """integration() with mto = 'bigru' (AR_MERGE: Bi-GRU with return_sequences=False -- train.py's default MANY_TO_ONE) and
    'avg' (GlobalAveragePooling1D) under AR_DS -> AR_DS_LN -> head: autograd vs central differences."""
    B, S, C0, Dd, n = 4, 5, 7, 6, 8
    um = 3
    rng = np.random.RandomState(37)
    params = _params("softmax", D=(2 * um if mto == "bigru" else Dd))
    params["AR_DS/kernel"] = rng.randn(C0, Dd) * 0.4
    params["AR_DS/bias"] = rng.randn(Dd) * 0.1
    params["AR_DS_LN/gamma"] = rng.uniform(0.7, 1.3, Dd)
    params["AR_DS_LN/beta"] = rng.randn(Dd) * 0.1
    for d in ("forward", "backward"):
        params["AR_MERGE/%s/kernel" % d] = rng.randn(Dd, 3 * um) * 0.4
        params["AR_MERGE/%s/recurrent_kernel" % d] = rng.randn(um, 3 * um) * 0.4
        params["AR_MERGE/%s/bias" % d] = rng.randn(6 * um) * 0.1
    x = rng.randn(B, S, C0)
    onehot = np.eye(n)[rng.randint(0, n, B)]
    kw = dict(disc_enable=True, metric_loss="softmax", margin=0.0, w_accent=0.01, w_disc=0.6)
    pool = dict(mto=mto, vlad_clusters=0, ghost_clusters=0, train_ds=True)
    _, state, losses, grads = TO.train_step(params, {}, x, onehot, lr=0.01, iterations=0, pool=pool, **kw)
    assert set(TO.pool_keys(mto)) | set(TO.DS_KEYS) <= set(grads)

    def loss_of(pp):
        t = {k: torch.as_tensor(v) for k, v in pp.items()}
        return float(TO.pooled_head_loss(t, torch.as_tensor(x), torch.as_tensor(onehot), **pool, **kw)[0])
    for k in TO.pool_keys(mto) + TO.DS_KEYS:
        g = grads[k]
        for _ in range(4):
            i = tuple(rng.randint(0, s_) for s_ in g.shape)
            h = 1e-6
            pp, pm = {q: v.copy() for q, v in params.items()}, {q: v.copy() for q, v in params.items()}
            pp[k][i] += h; pm[k][i] -= h
            fd = (loss_of(pp) - loss_of(pm)) / (2 * h)
            assert abs(fd - g[i]) <= 2e-6 * max(1.0, abs(fd)) + 2e-8, (k, i, fd, g[i])
