"""Training mode, first slice: the csrc/train.cu building blocks and one HeadTrainer step against the float64
autograd oracle (oracle/train_oracle.py, itself pinned by finite differences in tests/test_oracle_train.py)."""
import numpy as np
import pytest
import torch

from helpers import norm_err, dev
from oracle import train_oracle as TO, sarnet_oracle as O

pytestmark = pytest.mark.gpu

KINDS = [("arcface", 0.3), ("cosface", 0.3), ("sphereface", 1.35), ("softmax", 0.0), ("circleloss", 0.2)]


@pytest.mark.parametrize("M,N,K,ta,tb", [(64, 256, 16384, False, False), (16384, 256, 48, True, False), (48, 16384, 256, False, True),
                                         (7, 5, 3, False, False), (65, 130, 17, True, True), (1, 8, 64, False, False),
                                         (16, 768, 256, False, False), (16, 256, 768, False, True), (12, 250, 16385, False, False),
                                         (32, 77, 100, False, True), (17, 40, 64, False, False), (64, 768, 256, False, False),
                                         (64, 256, 768, False, True), (33, 70, 129, False, True)])       # skinny kernel (M <= 64, A not transposed)
def test_gemm_all_transposes(cuda_device, M, N, K, ta, tb):
    from aesrc2020_b200 import training as T
    rng = np.random.RandomState(M + N)
    a = rng.randn(K, M) if ta else rng.randn(M, K)
    b = rng.randn(N, K) if tb else rng.randn(K, N)
    want = (a.T if ta else a) @ (b.T if tb else b)
    got = T.gemm(dev(a), dev(b), ta=ta, tb=tb)
    assert norm_err(got, want) < 3e-5            # fp32 accumulation over K up to 16384
    c0 = rng.randn(M, N)
    out = dev(c0)
    T.gemm(dev(a), dev(b), ta=ta, tb=tb, alpha=0.5, beta=2.0, out=out)
    assert norm_err(out, 0.5 * want + 2.0 * c0) < 3e-5


@pytest.mark.parametrize("rows,C", [(37, 300), (2049, 33), (40000, 32), (7777, 256)])
def test_bn_train_forward_backward(cuda_device, rows, C):
    """rows <= 2048: one thread per channel (the head); more rows: the row-parallel kernels (the ResNet's maps), incl.
    ragged chunks, a channel count that is not a multiple of 32, and the column-sum kernel they share."""
    from aesrc2020_b200 import training as T
    rng = np.random.RandomState(3)
    x = rng.randn(rows, C) * 2 + 0.5
    g, b = rng.uniform(0.5, 1.5, C), rng.randn(C)
    mm, mv = rng.randn(C), rng.uniform(0.5, 2, C)
    dy = rng.randn(rows, C)
    xt = torch.tensor(x, requires_grad=True)
    gt, bt = torch.tensor(g, requires_grad=True), torch.tensor(b, requires_grad=True)
    y, mean, var = TO.bn_train(xt, gt, bt)
    (y * torch.as_tensor(dy)).sum().backward()
    mmd, mvd = dev(mm), dev(mv)
    yd, mean_d, inv_d = T.bn_train_fwd(dev(x), dev(g), dev(b), mmd, mvd)
    assert norm_err(yd, y.detach()) < 2e-6 and norm_err(mean_d, mean.detach()) < 2e-6
    assert norm_err(mmd, 0.99 * mm + 0.01 * mean.detach().numpy()) < 2e-6
    assert norm_err(mvd, 0.99 * mv + 0.01 * var.detach().numpy()) < 2e-6
    dx, dg, db = T.bn_train_bwd(dev(x), dev(dy), dev(g), mean_d, inv_d)
    assert norm_err(dx, xt.grad) < 1e-5 and norm_err(dg, gt.grad) < 1e-5 and norm_err(db, bt.grad) < 1e-5
    assert norm_err(T.colsum(dev(dy)), dy.sum(0)) < 1e-5


@pytest.mark.parametrize("axis", [0, 1])
def test_l2norm_forward_backward(cuda_device, axis):
    from aesrc2020_b200 import training as T
    rng = np.random.RandomState(axis)
    v = rng.randn(9, 13)
    u = rng.randn(9, 13)
    vt = torch.tensor(v, requires_grad=True)
    vh = vt / torch.sqrt(torch.clamp((vt * vt).sum(1 if axis else 0, keepdim=True), min=1e-12))
    (vh * torch.as_tensor(u)).sum().backward()
    vhd, inv = T.l2norm_fwd(dev(v), axis)
    assert norm_err(vhd, vh.detach()) < 2e-6
    got = T.l2norm_bwd(vhd, inv, dev(u), axis)
    assert norm_err(got, vt.grad) < 1e-5
    acc = dev(np.ones((9, 13)))
    T.l2norm_bwd(vhd, inv, dev(u), axis, out=acc, beta=1.0)
    assert norm_err(acc, vt.grad + 1.0) < 1e-5


def test_adam_and_unit_norm(cuda_device):
    from aesrc2020_b200 import training as T
    rng = np.random.RandomState(8)
    p0 = rng.randn(40, 8)
    p, m, v = torch.as_tensor(p0.copy()), torch.zeros(40, 8, dtype=torch.float64), torch.zeros(40, 8, dtype=torch.float64)
    pd, md, vd = dev(p0), dev(np.zeros((40, 8))), dev(np.zeros((40, 8)))
    for it in range(5):
        g = rng.randn(40, 8)
        p, m, v = TO.adam_update(p, torch.as_tensor(g) + 2 * 1e-4 * p, m, v, it, lr=0.01)
        T.adam_step(pd, dev(g), md, vd, T.adam_lr_t(0.01, it), l2=1e-4)
        assert norm_err(pd, p) < 5e-6 and norm_err(md, m) < 5e-6 and norm_err(vd, v) < 5e-5
    T.unit_norm(pd)
    assert np.allclose(np.sqrt((pd.cpu().numpy() ** 2).sum(0)), 1.0, atol=1e-6)


def _params(kind, D, E=256, n=8, seed=0):
    rng = np.random.RandomState(seed)
    p = {"AR_BN1/gamma": rng.uniform(0.7, 1.3, D), "AR_BN1/beta": rng.randn(D) * 0.1, "AR_BN1/moving_mean": rng.randn(D) * 0.1,
         "AR_BN1/moving_variance": rng.uniform(0.6, 1.4, D),
         "AR_EMBEDDING/kernel": rng.randn(D, E) / np.sqrt(D), "AR_EMBEDDING/bias": rng.randn(E) * 0.1,
         "AR_BN2/gamma": rng.uniform(0.7, 1.3, E), "AR_BN2/beta": rng.randn(E) * 0.1, "AR_BN2/moving_mean": rng.randn(E) * 0.1,
         "AR_BN2/moving_variance": rng.uniform(0.6, 1.4, E),
         "AR_CF_DS1/kernel": rng.randn(E, 64) / 16, "AR_CF_DS1/bias": rng.randn(64) * 0.1,
         "AR_CF_DS2/kernel": rng.randn(64, 64) / 8, "AR_CF_DS2/bias": rng.randn(64) * 0.1,
         "y_accent/kernel": rng.randn(64, n) / 8, "y_accent/bias": rng.randn(n) * 0.1}
    W = rng.uniform(-0.15, 0.15, (E, n))
    if kind == "circleloss":
        W = W / np.sqrt((W ** 2).sum(0, keepdims=True))
    p[TO.disc_key(kind)] = W
    return {k: np.asarray(v, np.float32).astype(np.float64) for k, v in p.items()}


@pytest.mark.parametrize("kind,margin", KINDS)
def test_head_trainer_steps_match_the_autograd_oracle(cuda_device, kind, margin):
    """Three consecutive optimisation steps (gradients, Adam state, BN moving statistics, unit_norm constraint) of
    HeadTrainer.step_on_features vs oracle.train_oracle.train_step on the same features."""
    from aesrc2020_b200 import model as mdl, training as T
    model, _ = mdl.SAR_Net((200, 80, 1), ctc_enable=True, disc_enable=True, res_type="res34", res_filters=32, mto="gvlad",
                           vlad_clusters=8, ghost_clusters=2, metric_loss=kind, margin=margin)
    D = 8 * 256
    params = _params(kind, D, seed=len(kind))
    for k, v in params.items():
        model.weights[k] = v.astype(np.float32)
    tr = T.HeadTrainer(model, lr=0.01)
    assert abs(tr.w_acc - 0.01) < 1e-12 and abs(tr.w_disc - 0.6) < 1e-12          # model.py:344-367 with CTC + disc
    rng = np.random.RandomState(11)
    B = 24
    lab = rng.randint(0, 8, B)
    state = {}
    p_or = dict(params)
    p_prev = dict(params)
    for it in range(3):
        integ = (rng.randn(B, D) * 0.05 + np.eye(8)[lab] @ rng.randn(8, D) * 0.02).astype(np.float32)
        onehot = np.eye(8, dtype=np.float32)[lab]
        p_or, state, l_or, g_or = TO.train_step(p_or, state, integ, onehot, lr=0.01, iterations=it, disc_enable=True,
                                                metric_loss=kind, margin=margin, w_accent=tr.w_acc, w_disc=tr.w_disc)
        got = tr.step_on_features(dev(integ), dev(onehot))
        assert abs(got["loss_accent"] - l_or["loss_accent"]) < 1e-4 * max(1, abs(l_or["loss_accent"]))
        assert abs(got["loss_disc"] - l_or["loss_disc"]) < 2e-4 * max(1, abs(l_or["loss_disc"]))
        for k in tr.keys:                                  # raw gradients (before the l2 term, which Adam's kernel adds)
            reg = 2 * TO.L2_REG * p_prev[k] if k in TO.l2_keys(True, kind) else 0.0
            # (AR_BN1/beta has an exactly zero gradient -- AR_BN2 removes any constant shift of its input -- so the
            # error is measured against a floor, not against the reference's own 1e-18 rounding noise)
            want = g_or[k] - reg
            got_k = tr.last_grads[k].cpu().numpy().astype(np.float64)
            if k in ("AR_BN1/beta", "AR_EMBEDDING/bias"):
                # a constant shift ahead of a batch-statistics BatchNorm has an exactly-zero data gradient
                # (AR_BN2 removes it); the fp32 path leaves rounding noise there, so bound it by the scale of the
                # neighbouring gradient instead of a relative error against ~0
                scale = float(tr.last_grads[k.split("/")[0] + ("/gamma" if "BN1" in k else "/kernel")].abs().max())
                assert np.max(np.abs(got_k)) <= 1e-3 * scale + 1e-7 and np.max(np.abs(want)) < 1e-9, (it, k)
                continue
            err = float(np.max(np.abs(got_k - want)) / max(np.max(np.abs(want)), 1e-6))
            assert err < 3e-4, (it, k, err)
        p_prev = {k: v.copy() for k, v in p_or.items()}
        for k in tr.keys + ["AR_BN1/moving_mean", "AR_BN1/moving_variance", "AR_BN2/moving_mean", "AR_BN2/moving_variance"]:
            if k == "AR_BN1/beta":
                # Adam normalises the gradient: fp32 rounding noise of ~1e-8 on an exactly-zero gradient moves the
                # parameter by up to lr * |g| / (|g| + 1e-7) per step (Keras' own fp32 graph does the same)
                assert float(np.max(np.abs(tr.p[k].cpu().numpy() - p_or[k]))) <= 0.01 * (it + 1), (it, k)
                continue
            # (Adam's first steps are sign-like, lr * g / (|g| + 1e-7): elements whose gradient is near 1e-7 turn a
            # 1e-8 gradient rounding difference into a few % of one lr step -- hence 2e-3 of max|p|, not 1e-5)
            assert norm_err(tr.p[k], p_or[k]) < 2e-3, (it, k)
    tr.sync_to_model()
    assert norm_err(torch.as_tensor(model.weights["AR_EMBEDDING/kernel"]), p_or["AR_EMBEDDING/kernel"]) < 2e-3


def test_train_on_batch_lowers_the_loss_through_the_frozen_encoder(cuda_device):
    """Keras-like surface: train_on_batch(x) = frozen encoder forward -> head step; 25 steps on one batch reduce the loss,
    and the updated weights change model.predict."""
    from aesrc2020_b200 import model as mdl, training as T, utils as us
    model, _ = mdl.SAR_Net((200, 80, 1), disc_enable=True, res_type="res34", res_filters=32, mto="gvlad", vlad_clusters=8,
                           ghost_clusters=2, metric_loss="arcface", margin=0.3)
    x, y = us.synthetic_batch(model.config, 16, seed=3)
    before = model.predict(x, batch_size=16)[1]
    tr = T.HeadTrainer(model, lr=0.02)
    hist = [tr.train_on_batch(x, y)["loss"] for _ in range(25)]
    assert hist[-1] < 0.7 * hist[0], hist
    tr.sync_to_model()
    after = model.predict(x, batch_size=16)[1]
    assert not np.allclose(before, after)
    lab = np.argmax(x["x_accent"], -1)
    assert np.mean(np.argmax(after, -1) == lab) >= np.mean(np.argmax(before, -1) == lab)


@pytest.mark.parametrize("mto,G", [("gvlad", 2), ("vlad", 0)])
def test_vlad_training_kernels_match_autograd(cuda_device, mto, G):
    """sar_vlad_train_fwd / sar_vlad_train_bwd (+ the per-cluster l2norm) vs torch autograd on oracle.vlad (float64)."""
    from aesrc2020_b200 import training as T
    from oracle import sarnet_oracle as O
    rng = np.random.RandomState(5 + G)
    B, S, D, K = 6, 13, 256, 8
    x = rng.randn(B, S, D) * 0.7
    wa = rng.randn(D, K + G) * 0.1
    ba = rng.randn(K + G) * 0.1
    c = rng.randn(K + G, D) * 0.5
    u = rng.randn(B, K * D)
    p = {mto + "_center_assignment/kernel": torch.tensor(wa.reshape(1, 1, D, K + G), requires_grad=True),
         mto + "_center_assignment/bias": torch.tensor(ba, requires_grad=True),
         mto + "_pool/centers": torch.tensor(c, requires_grad=True)}
    want = O.integration(torch.as_tensor(x), p, D, mto, K, G)
    (want * torch.as_tensor(u)).sum().backward()
    xd, cd = dev(x), dev(c)
    A, R, asum = T.vlad_train_fwd(xd, dev(wa), dev(ba), cd, K, G)
    assert abs(float(A.sum()) - B * S) < 1e-3 and norm_err(asum, A[:, :, :K].sum(1)) < 1e-6
    V, rinv = T.l2norm_fwd(R.view(B * K, D), 1)
    assert norm_err(V.view(B, K * D), want.detach()) < 2e-6
    gR = T.l2norm_bwd(V, rinv, dev(u).view(B * K, D), 1)
    g_scores, gc_part = T.vlad_train_bwd(xd, A, cd, gR, asum, K, G)
    g_wa = T.gemm(xd.view(B * S, D), g_scores.view(B * S, K + G), ta=True)
    g_ba = T.colsum(g_scores.view(B * S, K + G))
    g_c = T.colsum(gc_part.view(B, K * D)).view(K, D)
    assert norm_err(g_wa, p[mto + "_center_assignment/kernel"].grad.reshape(D, K + G)) < 2e-5
    assert norm_err(g_ba, p[mto + "_center_assignment/bias"].grad) < 2e-5
    assert norm_err(g_c, p[mto + "_pool/centers"].grad[:K]) < 2e-5
    assert float(p[mto + "_pool/centers"].grad[K:].abs().max()) == 0.0 if G else True


def test_head_trainer_with_pooling_layer_matches_the_oracle(cuda_device):
    """Second slice: HeadTrainer(train_pool=True).step_on_features on AR_DS_LN descriptors vs train_oracle.train_step(pool=...):
    losses, every gradient (head + assignment Conv2D + centers) and the parameters after two Adam steps."""
    from aesrc2020_b200 import model as mdl, training as T
    K, G, Dh = 8, 2, 256
    model, _ = mdl.SAR_Net((200, 80, 1), ctc_enable=True, disc_enable=True, res_type="res34", res_filters=32, mto="gvlad",
                           vlad_clusters=K, ghost_clusters=G, metric_loss="arcface", margin=0.3)
    params = _params("arcface", K * Dh, seed=9)
    rng = np.random.RandomState(21)
    params["gvlad_center_assignment/kernel"] = (rng.randn(1, 1, Dh, K + G) * 0.1).astype(np.float32).astype(np.float64)
    params["gvlad_center_assignment/bias"] = (rng.randn(K + G) * 0.1).astype(np.float32).astype(np.float64)
    params["gvlad_pool/centers"] = (rng.randn(K + G, Dh) * 0.3).astype(np.float32).astype(np.float64)
    for k, v in params.items():
        model.weights[k] = v.astype(np.float32)
    tr = T.HeadTrainer(model, lr=0.01, train_pool=True)
    assert set(TO.pool_keys("gvlad")) <= set(tr.keys)
    B, S = 16, 12
    lab = rng.randint(0, 8, B)
    pool = dict(mto="gvlad", vlad_clusters=K, ghost_clusters=G)
    state, p_or, p_prev = {}, dict(params), dict(params)
    l2k = set(TO.l2_keys(True, "arcface")) | set(TO.pool_l2_keys("gvlad"))
    for it in range(2):
        feat = (rng.randn(B, S, Dh) * 0.5 + (np.eye(8)[lab] @ rng.randn(8, Dh))[:, None, :] * 0.3).astype(np.float32)
        onehot = np.eye(8, dtype=np.float32)[lab]
        p_or, state, l_or, g_or = TO.train_step(p_or, state, feat, onehot, lr=0.01, iterations=it, disc_enable=True,
                                                metric_loss="arcface", margin=0.3, w_accent=tr.w_acc, w_disc=tr.w_disc, pool=pool)
        got = tr.step_on_features(dev(feat), dev(onehot))
        assert abs(got["loss_accent"] - l_or["loss_accent"]) < 1e-4 * max(1, abs(l_or["loss_accent"]))
        assert abs(got["loss_disc"] - l_or["loss_disc"]) < 2e-4 * max(1, abs(l_or["loss_disc"]))
        for k in tr.keys:
            if k in ("AR_BN1/beta", "AR_EMBEDDING/bias"):
                continue                                   # exactly-zero data gradients (see the test above)
            want = g_or[k] - (2 * TO.L2_REG * p_prev[k] if k in l2k else 0.0)
            got_k = tr.last_grads[k].cpu().numpy().astype(np.float64)
            err = float(np.max(np.abs(got_k - want)) / max(np.max(np.abs(want)), 1e-6))
            assert err < 5e-4, (it, k, err)
        p_prev = {k: v.copy() for k, v in p_or.items()}
        for k in TO.pool_keys("gvlad") + ["AR_EMBEDDING/kernel", "y_disc/W"]:
            assert norm_err(tr.p[k], p_or[k]) < 2e-3, (it, k)
    tr.sync_to_model()
    assert model.weights["gvlad_center_assignment/kernel"].shape == (1, 1, Dh, K + G)


def test_train_on_batch_with_pooling_layer_lowers_the_loss(cuda_device):
    from aesrc2020_b200 import model as mdl, training as T, utils as us
    model, _ = mdl.SAR_Net((200, 80, 1), disc_enable=True, res_type="res34", res_filters=32, mto="gvlad", vlad_clusters=8,
                           ghost_clusters=2, metric_loss="arcface", margin=0.3)
    x, y = us.synthetic_batch(model.config, 16, seed=3)
    c0 = model.weights["gvlad_pool/centers"].copy()
    tr = T.HeadTrainer(model, lr=0.02, train_pool=True)
    hist = [tr.train_on_batch(x, y)["loss"] for _ in range(25)]
    assert hist[-1] < 0.7 * hist[0], hist
    tr.sync_to_model()
    assert not np.allclose(c0[:8], model.weights["gvlad_pool/centers"][:8])
    assert np.array_equal(c0[8:], model.weights["gvlad_pool/centers"][8:])       # ghost centers never move


def test_ln_tanh_backward_matches_autograd(cuda_device):
    from aesrc2020_b200 import training as T
    rng = np.random.RandomState(4)
    rows, C = 37, 256
    pre = rng.randn(rows, C) * 0.8
    gamma, beta = rng.uniform(0.5, 1.5, C), rng.randn(C) * 0.1
    gz = rng.randn(rows, C)
    for tanh_in in (True, False):
        pt = torch.tensor(pre, requires_grad=True)
        gt, bt = torch.tensor(gamma, requires_grad=True), torch.tensor(beta, requires_grad=True)
        y = torch.tanh(pt) if tanh_in else pt
        mu = y.mean(-1, keepdim=True)
        z = (y - mu) / torch.sqrt(((y - mu) ** 2).mean(-1, keepdim=True) + 1e-14) * gt + bt
        (z * torch.as_tensor(gz)).sum().backward()
        g_pre, gzx = T.ln_train_bwd(dev(y.detach().numpy()), dev(gamma), dev(gz), tanh_in)
        assert norm_err(g_pre, pt.grad) < 2e-5
        assert norm_err(T.colsum(gzx), gt.grad) < 2e-5 and norm_err(T.colsum(dev(gz)), bt.grad) < 2e-6


def test_head_trainer_third_slice_matches_the_oracle(cuda_device):
    """HeadTrainer(train_ds=True).step_on_features on CRNN_LN features vs train_oracle.train_step(pool=dict(train_ds=True)):
    AR_DS / AR_DS_LN / pooling / head gradients and the parameters after two Adam steps."""
    from aesrc2020_b200 import model as mdl, training as T
    K, G, Dh = 8, 2, 256
    model, _ = mdl.SAR_Net((200, 80, 1), ctc_enable=True, disc_enable=True, res_type="res34", res_filters=32, mto="gvlad",
                           vlad_clusters=K, ghost_clusters=G, metric_loss="cosface", margin=0.3)
    params = _params("cosface", K * Dh, seed=13)
    rng = np.random.RandomState(31)
    f32 = lambda a: np.asarray(a, np.float32).astype(np.float64)
    params["gvlad_center_assignment/kernel"] = f32(rng.randn(1, 1, Dh, K + G) * 0.1)
    params["gvlad_center_assignment/bias"] = f32(rng.randn(K + G) * 0.1)
    params["gvlad_pool/centers"] = f32(rng.randn(K + G, Dh) * 0.3)
    params["AR_DS/kernel"] = f32(rng.randn(2 * Dh, Dh) / np.sqrt(2 * Dh))
    params["AR_DS/bias"] = f32(rng.randn(Dh) * 0.1)
    params["AR_DS_LN/gamma"] = f32(rng.uniform(0.7, 1.3, Dh))
    params["AR_DS_LN/beta"] = f32(rng.randn(Dh) * 0.1)
    for k, v in params.items():
        model.weights[k] = v.astype(np.float32)
    tr = T.HeadTrainer(model, lr=0.01, train_ds=True)
    assert tr.train_pool and set(TO.DS_KEYS) <= set(tr.keys)
    B, S = 12, 10
    lab = rng.randint(0, 8, B)
    pool = dict(mto="gvlad", vlad_clusters=K, ghost_clusters=G, train_ds=True)
    state, p_or, p_prev = {}, dict(params), dict(params)
    l2k = set(TO.l2_keys(True, "cosface")) | set(TO.pool_l2_keys("gvlad")) | {"AR_DS/kernel", "AR_DS/bias"}
    for it in range(2):
        crnn = (rng.randn(B, S, 2 * Dh) + (np.eye(8)[lab] @ rng.randn(8, 2 * Dh))[:, None, :] * 0.5).astype(np.float32)
        onehot = np.eye(8, dtype=np.float32)[lab]
        p_or, state, l_or, g_or = TO.train_step(p_or, state, crnn, onehot, lr=0.01, iterations=it, disc_enable=True,
                                                metric_loss="cosface", margin=0.3, w_accent=tr.w_acc, w_disc=tr.w_disc, pool=pool)
        got = tr.step_on_features(dev(crnn), dev(onehot))
        assert abs(got["loss_disc"] - l_or["loss_disc"]) < 2e-4 * max(1, abs(l_or["loss_disc"]))
        for k in tr.keys:
            if k in ("AR_BN1/beta", "AR_EMBEDDING/bias"):
                continue
            want = g_or[k] - (2 * TO.L2_REG * p_prev[k] if k in l2k else 0.0)
            got_k = tr.last_grads[k].cpu().numpy().astype(np.float64)
            err = float(np.max(np.abs(got_k - want)) / max(np.max(np.abs(want)), 1e-6))
            assert err < 1e-3, (it, k, err)
        p_prev = {k: v.copy() for k, v in p_or.items()}
        for k in TO.DS_KEYS + TO.pool_keys("gvlad"):
            assert norm_err(tr.p[k], p_or[k]) < 2e-3, (it, k)


def test_train_on_batch_third_slice_lowers_the_loss(cuda_device):
    from aesrc2020_b200 import model as mdl, training as T, utils as us
    model, _ = mdl.SAR_Net((200, 80, 1), disc_enable=True, res_type="res34", res_filters=32, mto="gvlad", vlad_clusters=8,
                           ghost_clusters=2, metric_loss="arcface", margin=0.3)
    x, y = us.synthetic_batch(model.config, 16, seed=3)
    w0 = model.weights["AR_DS/kernel"].copy()
    tr = T.HeadTrainer(model, lr=0.02, train_ds=True)
    hist = [tr.train_on_batch(x, y)["loss"] for _ in range(25)]
    assert hist[-1] < 0.7 * hist[0], hist
    tr.sync_to_model()
    assert not np.allclose(w0, model.weights["AR_DS/kernel"])
    out = model.predict(x, batch_size=16)                   # the inference engine rebuilds with the trained weights
    assert np.all(np.isfinite(out[0]))


def test_gru_training_step_kernels_match_autograd(cuda_device):
    """sar_gru_gate_fwd / sar_gru_gate_bwd through training.gru_dir_fwd / gru_dir_bwd (one GEMM + one gate kernel per time
    step, BPTT) vs float64 autograd through the oracle's CuDNNGRU restatement, both directions, ragged sizes."""
    from aesrc2020_b200 import training as T
    rng = np.random.RandomState(5)
    B, S, Din, u = 5, 7, 24, 16
    f32 = lambda a: np.asarray(a, np.float32)
    x = f32(rng.randn(B, S, Din))
    gout = f32(rng.randn(B, S, 2 * u))
    for di, (dname, reverse) in enumerate((("forward", False), ("backward", True))):
        W, U, b = f32(rng.randn(Din, 3 * u) * 0.3), f32(rng.randn(u, 3 * u) * 0.3), f32(rng.randn(6 * u) * 0.2)
        tx, tW, tU, tb = (torch.tensor(a.astype(np.float64), requires_grad=True) for a in (x, W, U, b))
        o, _ = O.gru_direction(tx, tW, tU, tb, reverse)
        (o * torch.tensor(gout[:, :, di * u:(di + 1) * u].astype(np.float64))).sum().backward()
        out = torch.zeros((B, S, 2 * u), device="cuda")
        sv = T.gru_dir_fwd(dev(x).view(B * S, Din), B, S, dev(W), dev(U), dev(b), reverse, out, di * u)
        assert norm_err(out[:, :, di * u:(di + 1) * u], o.detach()) < 1e-5
        gW, gU, gb, gx = T.gru_dir_bwd(dev(gout), sv, B, S, dev(W), dev(U))
        for name, got, want in (("kernel", gW, tW.grad), ("recurrent", gU, tU.grad), ("bias", gb, tb.grad),
                                ("x", gx.view(B, S, Din), tx.grad)):
            assert norm_err(got, want) < 2e-5, (dname, name, norm_err(got, want))


def test_head_trainer_fourth_slice_matches_the_oracle(cuda_device):
    """HeadTrainer(train_crnn=True).step_on_features on the frozen ResNet's sequence vs
    train_oracle.train_step(pool=dict(train_crnn=True)): gradients of CNN_LIN / CNN_LIN_LN / both CRNN directions / CRNN_LN
    (and of everything above) and the parameters after two Adam steps."""
    from aesrc2020_b200 import model as mdl, training as T
    K, G, Dh = 8, 2, 256
    model, _ = mdl.SAR_Net((200, 80, 1), ctc_enable=False, disc_enable=True, res_type="res34", res_filters=32, mto="gvlad",
                           vlad_clusters=K, ghost_clusters=G, metric_loss="arcface", margin=0.3)
    Cc = model.config.plan().cout
    params = _params("arcface", K * Dh, seed=17)
    rng = np.random.RandomState(37)
    f32 = lambda a: np.asarray(a, np.float32).astype(np.float64)
    params["gvlad_center_assignment/kernel"] = f32(rng.randn(1, 1, Dh, K + G) * 0.1)
    params["gvlad_center_assignment/bias"] = f32(rng.randn(K + G) * 0.1)
    params["gvlad_pool/centers"] = f32(rng.randn(K + G, Dh) * 0.3)
    for k in TO.DS_KEYS + TO.CRNN_KEYS:
        params[k] = f32(model.weights[k])
    for k, v in params.items():
        model.weights[k] = v.astype(np.float32)
    tr = T.HeadTrainer(model, lr=0.01, train_crnn=True)
    assert tr.train_ds and tr.train_pool and set(TO.CRNN_KEYS) <= set(tr.keys)
    B, S = 6, 9
    lab = rng.randint(0, 8, B)
    pool = dict(mto="gvlad", vlad_clusters=K, ghost_clusters=G, train_crnn=True)
    state, p_or, p_prev = {}, dict(params), dict(params)
    l2k = set(TO.l2_keys(True, "arcface")) | set(TO.pool_l2_keys("gvlad")) | {"AR_DS/kernel", "AR_DS/bias"} | set(TO.CRNN_L2_KEYS)
    for it in range(2):
        seq = np.maximum(rng.randn(B, S, Cc) + (np.eye(8)[lab] @ rng.randn(8, Cc))[:, None, :] * 0.5, 0).astype(np.float32)
        onehot = np.eye(8, dtype=np.float32)[lab]
        p_or, state, l_or, g_or = TO.train_step(p_or, state, seq, onehot, lr=0.01, iterations=it, disc_enable=True,
                                                metric_loss="arcface", margin=0.3, w_accent=tr.w_acc, w_disc=tr.w_disc, pool=pool)
        got = tr.step_on_features(dev(seq), dev(onehot))
        assert abs(got["loss_disc"] - l_or["loss_disc"]) < 2e-4 * max(1, abs(l_or["loss_disc"]))
        for k in tr.keys:
            if k in ("AR_BN1/beta", "AR_EMBEDDING/bias"):
                continue
            want = g_or[k] - (2 * TO.L2_REG * p_prev[k] if k in l2k else 0.0)
            got_k = tr.last_grads[k].cpu().numpy().astype(np.float64)
            err = float(np.max(np.abs(got_k - want)) / max(np.max(np.abs(want)), 1e-6))
            assert err < 2e-3, (it, k, err)
        p_prev = {k: v.copy() for k, v in p_or.items()}
        # Adam's first steps move every element by ~lr * sign(g): the few elements whose fp32 gradient is within rounding of
        # zero can take the other sign (a 2 * lr difference each), so the parameters get a looser bound than the gradients
        for k in TO.CRNN_KEYS:
            assert norm_err(tr.p[k], p_or[k]) < 8e-3, (it, k)


def test_train_on_batch_fourth_slice_lowers_the_loss(cuda_device):
    from aesrc2020_b200 import model as mdl, training as T, utils as us
    model, _ = mdl.SAR_Net((200, 80, 1), disc_enable=True, res_type="res34", res_filters=32, mto="gvlad", vlad_clusters=8,
                           ghost_clusters=2, metric_loss="arcface", margin=0.3)
    x, y = us.synthetic_batch(model.config, 16, seed=3)
    w0 = model.weights["CRNN/forward/recurrent_kernel"].copy()
    tr = T.HeadTrainer(model, lr=0.01, train_crnn=True)
    hist = [tr.train_on_batch(x, y)["loss"] for _ in range(20)]
    assert hist[-1] < 0.7 * hist[0], hist
    tr.sync_to_model()
    assert not np.allclose(w0, model.weights["CRNN/forward/recurrent_kernel"])
    out = model.predict(x, batch_size=16)
    assert np.all(np.isfinite(out[0]))


@pytest.mark.parametrize("B,S,C,Lmax,ld", [(5, 12, 9, 4, 9), (3, 75, 1000, 72, 1024), (4, 21, 33, 6, 64)])
def test_ctc_gradient_kernel_matches_autograd(cuda_device, B, S, C, Lmax, ld):
    """sar_ctc_grad_fwd (alpha-beta recursion, gradient through q = softmax(log(softmax(a) + 1e-7))) vs float64 autograd
    of the oracle's CTC: losses and d loss / d logits, ragged input / label lengths, repeated labels, padded rows
    (ld > C: the tensor-core ctc_pred layout), frames past in_len get zero gradient."""
    from aesrc2020_b200 import training as T
    rng = np.random.RandomState(B * 100 + S)
    logits = (rng.randn(B, S, ld) * 2).astype(np.float32)
    lab_len = rng.randint(1, Lmax + 1, B)
    in_len = np.array([rng.randint(min(S, 2 * l + 1), S + 1) for l in lab_len])
    labels = np.zeros((B, Lmax), np.float32)
    for b in range(B):
        labels[b, :lab_len[b]] = rng.randint(0, C - 1, lab_len[b])
        if lab_len[b] >= 2:
            labels[b, 1] = labels[b, 0]                      # a repeat: needs the blank between
    scale = 0.37
    ta = torch.tensor(logits[:, :, :C].astype(np.float64), requires_grad=True)
    lo = TO.ctc_loss_autograd(torch.softmax(ta, -1), labels, in_len, lab_len)
    (scale * lo.sum()).backward()
    loss, grad, status = T.ctc_grad(dev(logits), dev(labels), torch.from_numpy(in_len.astype(np.int32)).cuda(),
                                    torch.from_numpy(lab_len.astype(np.int32)).cuda(), scale, classes=C)
    assert int(status.abs().max()) == 0
    assert norm_err(loss, lo.detach()) < 1e-5
    # fp32 log-space alpha / beta over up to 75 frames x 145 lattice states: ~2e-4 of the largest gradient at the full size
    assert norm_err(grad, ta.grad) < (5e-5 if S * Lmax < 200 else 5e-4), norm_err(grad, ta.grad)
    for b in range(B):
        assert float(grad[b, in_len[b]:].abs().sum()) == 0.0


def test_head_trainer_fifth_slice_multitask_matches_the_oracle(cuda_device):
    """HeadTrainer(train_ctc=True): CTC branch + accent branch above the frozen ResNet, one step vs
    train_oracle.train_step(pool=dict(train_ctc=True, ctc=..., w_ctc=...)): every gradient, the three losses."""
    from aesrc2020_b200 import model as mdl, training as T
    K, G, Dh = 8, 2, 256
    model, _ = mdl.SAR_Net((200, 80, 1), ctc_enable=True, disc_enable=True, res_type="res34", res_filters=32, mto="gvlad",
                           vlad_clusters=K, ghost_clusters=G, metric_loss="arcface", margin=0.3, bpe_classes=50, max_ctc_len=6)
    Cc = model.config.plan().cout
    params = _params("arcface", K * Dh, seed=19)
    rng = np.random.RandomState(41)
    f32 = lambda a: np.asarray(a, np.float32).astype(np.float64)
    params["gvlad_center_assignment/kernel"] = f32(rng.randn(1, 1, Dh, K + G) * 0.1)
    params["gvlad_center_assignment/bias"] = f32(rng.randn(K + G) * 0.1)
    params["gvlad_pool/centers"] = f32(rng.randn(K + G, Dh) * 0.3)
    for k in TO.DS_KEYS + TO.CRNN_KEYS + TO.CTC_KEYS:
        params[k] = f32(model.weights[k])
    for k, v in params.items():
        model.weights[k] = v.astype(np.float32)
    tr = T.HeadTrainer(model, lr=0.01, train_ctc=True)
    assert tr.train_crnn and set(TO.CTC_KEYS) <= set(tr.keys) and tr.w_ctc > 0
    B, S = 5, 13
    lab = rng.randint(0, 8, B)
    labels = rng.randint(0, 49, (B, 6)).astype(np.float32)
    lab_len = rng.randint(1, 7, (B, 1)).astype(np.int32)
    in_len = np.full((B, 1), S, np.int32)
    pool = dict(mto="gvlad", vlad_clusters=K, ghost_clusters=G, train_ctc=True, ctc=(labels, in_len, lab_len), w_ctc=tr.w_ctc)
    l2k = set(TO.l2_keys(True, "arcface")) | set(TO.pool_l2_keys("gvlad")) | {"AR_DS/kernel", "AR_DS/bias"} | set(TO.CRNN_L2_KEYS) | set(TO.CTC_L2_KEYS)
    seq = np.maximum(rng.randn(B, S, Cc) + (np.eye(8)[lab] @ rng.randn(8, Cc))[:, None, :] * 0.5, 0).astype(np.float32)
    onehot = np.eye(8, dtype=np.float32)[lab]
    p_or, state, l_or, g_or = TO.train_step(dict(params), {}, seq, onehot, lr=0.01, iterations=0, disc_enable=True,
                                            metric_loss="arcface", margin=0.3, w_accent=tr.w_acc, w_disc=tr.w_disc, pool=pool)
    got = tr.step_on_features(dev(seq), dev(onehot), (dev(labels), torch.from_numpy(in_len).cuda(), torch.from_numpy(lab_len).cuda()))
    assert abs(got["loss_ctc"] - l_or["loss_ctc"]) < 2e-4 * max(1, abs(l_or["loss_ctc"]))
    assert abs(got["loss_disc"] - l_or["loss_disc"]) < 2e-4 * max(1, abs(l_or["loss_disc"]))
    for k in tr.keys:
        if k in ("AR_BN1/beta", "AR_EMBEDDING/bias"):
            continue
        want = g_or[k] - (2 * TO.L2_REG * params[k] if k in l2k else 0.0)
        got_k = tr.last_grads[k].cpu().numpy().astype(np.float64)
        err = float(np.max(np.abs(got_k - want)) / max(np.max(np.abs(want)), 1e-6))
        assert err < 2e-3, (k, err)


def test_train_on_batch_fifth_slice_lowers_both_losses(cuda_device):
    from aesrc2020_b200 import model as mdl, training as T, utils as us
    model, _ = mdl.SAR_Net((200, 80, 1), ctc_enable=True, disc_enable=True, res_type="res34", res_filters=32, mto="gvlad",
                           vlad_clusters=8, ghost_clusters=2, metric_loss="arcface", margin=0.3)
    x, y = us.synthetic_batch(model.config, 12, seed=5)
    tr = T.HeadTrainer(model, lr=0.005, train_ctc=True)
    hist = [tr.train_on_batch(x, y) for _ in range(30)]
    # Adam on 10 M parameters with a batch of 12: the curves are not monotone step by step -- compare the end of the run
    assert min(h["loss_ctc"] for h in hist[-5:]) < 0.8 * hist[0]["loss_ctc"], [round(h["loss_ctc"], 2) for h in hist]
    assert min(h["loss_disc"] for h in hist[-5:]) < hist[0]["loss_disc"], [round(h["loss_disc"], 2) for h in hist]
    assert min(h["loss"] for h in hist[-5:]) < 0.8 * hist[0]["loss"], [round(h["loss"], 2) for h in hist]
    tr.sync_to_model()
    outs = model.predict(x, batch_size=12)
    assert all(np.all(np.isfinite(o)) for o in outs)


@pytest.mark.parametrize("B,H,W,Cin,Cout,k,stride", [(2, 9, 8, 5, 7, 3, 1), (2, 9, 8, 5, 7, 3, 2), (3, 10, 7, 4, 6, 3, 2), (2, 8, 6, 8, 16, 1, 2),
                                                     (2, 20, 16, 1, 6, 7, 2), (1, 5, 4, 32, 32, 3, 1), (3, 9, 5, 64, 128, 3, 1),
                                                     (2, 11, 6, 70, 65, 3, 2), (2, 8, 6, 128, 256, 1, 2)])
def test_conv_backward_kernels_match_autograd(cuda_device, B, H, W, Cin, Cout, k, stride):
    """sar_conv2d_bwd_data / sar_conv2d_bwd_weight vs float64 autograd of the oracle's TF-SAME (k > 1) / VALID (1x1 shortcut)
    Conv2D: strides 1 and 2, even and odd sizes (asymmetric SAME pads), the 7x7 stem, accumulation into dx (beta = 1)."""
    from aesrc2020_b200 import training_resnet as TR
    from aesrc2020_b200.config import ConvSpec, same_pad
    rng = np.random.RandomState(B + H * 3 + k)
    x = rng.randn(B, H, W, Cin).astype(np.float32)
    w = (rng.randn(k, k, Cin, Cout) * 0.3).astype(np.float32)
    pad = "same" if k > 1 else "valid"
    if pad == "same":
        Ho, pt, _ = same_pad(H, k, stride)
        Wo, pl, _ = same_pad(W, k, stride)
    else:
        Ho, Wo, pt, pl = (H - 1) // stride + 1, (W - 1) // stride + 1, 0, 0
    tx, tw = torch.tensor(x.astype(np.float64), requires_grad=True), torch.tensor(w.astype(np.float64), requires_grad=True)
    y = O.conv2d(tx, tw, None, stride, pad)
    assert tuple(y.shape) == (B, Ho, Wo, Cout)
    gy = rng.randn(B, Ho, Wo, Cout).astype(np.float32)
    (y * torch.tensor(gy.astype(np.float64))).sum().backward()
    spec = ConvSpec(name="t", kh=k, kw=k, stride=stride, cin=Cin, cout=Cout, hin=H, win=W, hout=Ho, wout=Wo, pad_t=pt, pad_l=pl,
                    pre_bn=None, post_bn=None)
    dx = TR.conv_bwd_data(spec, dev(gy), dev(w), B)
    assert norm_err(dx, tx.grad) < 1e-5
    base = rng.randn(B, H, W, Cin).astype(np.float32)
    dx2 = TR.conv_bwd_data(spec, dev(gy), dev(w), B, dx=dev(base), beta=1.0)
    assert norm_err(dx2, tx.grad + torch.tensor(base.astype(np.float64))) < 1e-5
    dw = TR.conv_bwd_weight(spec, dev(x), dev(gy), B)
    assert norm_err(dw, tw.grad) < 1e-5


@pytest.mark.parametrize("B,H,W,C", [(2, 9, 8, 5), (2, 10, 7, 3), (1, 50, 40, 8)])
def test_maxpool_backward_matches_autograd(cuda_device, B, H, W, C):
    from aesrc2020_b200 import training_resnet as TR, ops
    from aesrc2020_b200.config import same_pad
    rng = np.random.RandomState(H)
    x = rng.randn(B, H, W, C).astype(np.float32)
    Ho, pt, _ = same_pad(H, 3, 2)
    Wo, pl, _ = same_pad(W, 3, 2)
    tx = torch.tensor(x.astype(np.float64), requires_grad=True)
    y = O.maxpool_same(tx)
    gy = rng.randn(B, Ho, Wo, C).astype(np.float32)
    (y * torch.tensor(gy.astype(np.float64))).sum().backward()
    dx = TR.maxpool_bwd(dev(x), dev(gy), 3, 2, pt, pl)
    assert norm_err(dx, tx.grad) < 1e-6


def test_whole_model_training_step_matches_the_oracle(cuda_device):
    """HeadTrainer(train_resnet=True): NOTHING frozen -- ResNet (training-mode BN) + CNN_LIN + CRNN + accent branch, one
    optimisation step vs the float64 autograd oracle: every gradient of the model and the accent / margin losses."""
    import warnings
    from aesrc2020_b200 import model as mdl, training as T
    from aesrc2020_b200.training_resnet import ResNetTrainer
    K, G = 8, 2
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model, _ = mdl.SAR_Net((100, 80, 1), ctc_enable=False, disc_enable=True, res_type="res18", res_filters=8, mto="gvlad",
                               vlad_clusters=K, ghost_clusters=G, metric_loss="arcface", margin=0.3)
    cfg = model.config
    rng = np.random.RandomState(43)
    for k in list(model.weights):                     # biases / betas away from zero
        if k.endswith("/bias") or k.endswith("/beta"):
            model.weights[k] = (model.weights[k] + rng.randn(*model.weights[k].shape) * 0.05).astype(np.float32)
    params = {k: np.asarray(v, np.float32).astype(np.float64) for k, v in model.weights.items()}
    rkeys, rl2, _ = ResNetTrainer.param_keys(cfg)
    tr = T.HeadTrainer(model, lr=0.01, train_resnet=True)
    assert set(rkeys) <= set(tr.keys) and tr.train_crnn
    B = 4
    lab = rng.randint(0, 8, B)
    x = rng.rand(B, 100, 80, 1).astype(np.float32)
    onehot = np.eye(8, dtype=np.float32)[lab]
    pool = dict(mto="gvlad", vlad_clusters=K, ghost_clusters=G, train_resnet=dict(res_type="res18", filters=8),
                extra_keys=tuple(rkeys), extra_l2=tuple(rl2))
    l2k = set(TO.l2_keys(True, "arcface")) | set(TO.pool_l2_keys("gvlad")) | {"AR_DS/kernel", "AR_DS/bias"} | set(TO.CRNN_L2_KEYS) | set(rl2)
    p_or, state, l_or, g_or = TO.train_step(dict(params), {}, x, onehot, lr=0.01, iterations=0, disc_enable=True,
                                            metric_loss="arcface", margin=0.3, w_accent=tr.w_acc, w_disc=tr.w_disc, pool=pool)
    got = tr.step_on_features(dev(x), dev(onehot))
    assert abs(got["loss_disc"] - l_or["loss_disc"]) < 5e-4 * max(1, abs(l_or["loss_disc"]))
    worst = {}
    for k in tr.keys:
        if k in ("AR_BN1/beta", "AR_EMBEDDING/bias"):
            continue
        want = g_or[k] - (2 * TO.L2_REG * params[k] if k in l2k else 0.0)
        got_k = tr.last_grads[k].cpu().numpy().astype(np.float64)
        if k.startswith("resnet/") and k.endswith("/bias"):
            # a constant added in front of a batch-statistic BN is removed by its mean: the true gradient of EVERY conv bias
            # of the ResNet is zero in training mode (float64: ~1e-17).  fp32 leaves rounding noise: bound it by the scale of
            # the same convolution's kernel gradient
            scale = float(np.max(np.abs(g_or[k[:-5] + "/kernel"])))
            assert float(np.max(np.abs(want))) < 1e-9 * max(scale, 1.0), (k, float(np.max(np.abs(want))))
            worst[k] = float(np.max(np.abs(got_k))) / max(scale, 1e-6) * 5.0          # passes below 1e-3 of the kernel gradient
            continue
        worst[k] = float(np.max(np.abs(got_k - want)) / max(np.max(np.abs(want)), 1e-6))
    bad = {k: round(v, 5) for k, v in worst.items() if v > 5e-3}
    assert not bad, sorted(bad.items(), key=lambda kv: -kv[1])[:12]


def test_whole_model_training_lowers_the_loss(cuda_device):
    import warnings
    from aesrc2020_b200 import model as mdl, training as T, utils as us
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model, _ = mdl.SAR_Net((100, 80, 1), ctc_enable=True, disc_enable=True, res_type="res18", res_filters=8, mto="gvlad",
                               vlad_clusters=8, ghost_clusters=2, metric_loss="arcface", margin=0.3, bpe_classes=40, max_ctc_len=4)
    x, y = us.synthetic_batch(model.config, 8, seed=7)
    w0 = model.weights["resnet/s2b1/conv1/kernel"].copy()
    m0 = model.weights["resnet/stem_bn/moving_mean"].copy()
    tr = T.HeadTrainer(model, lr=0.005, train_resnet=True, train_ctc=True)      # the reference's full multi-task fit
    hist = [tr.train_on_batch(x, y) for _ in range(15)]
    assert hist[-1]["loss"] < 0.8 * hist[0]["loss"], [h["loss"] for h in hist]
    tr.sync_to_model()
    assert not np.allclose(w0, model.weights["resnet/s2b1/conv1/kernel"])
    assert not np.allclose(m0, model.weights["resnet/stem_bn/moving_mean"])      # BN moving averages follow the batches
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        outs = model.predict(x, batch_size=8)
    assert all(np.all(np.isfinite(o)) for o in outs)


def test_fit_generator_mirrors_train_py(cuda_device, tmp_path):
    """train.py:25-44 end to end on the mirrored surface: utils.data_generator -> train_model.fit_generator(steps_per_epoch,
    epochs, callbacks=[a Callback that saves '%03d.h5' every epoch], validation_data) -> the saved checkpoint, loaded back through
    SAR_Net(raw_model=...), predicts what the trained model predicts.  Whole model, multi-task (CTC + accent), nothing frozen."""
    import warnings
    from aesrc2020_b200 import model as mdl, utils as us
    kw = dict(ctc_enable=True, ar_enable=True, disc_enable=True, res_type="res18", res_filters=8, mto="gvlad", vlad_clusters=8,
              ghost_clusters=2, metric_loss="arcface", margin=0.3, bpe_classes=40, max_ctc_len=4)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model, train_model = mdl.SAR_Net((100, 80, 1), lr=0.004, **kw)
    cfg = model.config
    rng = np.random.RandomState(11)
    lst = ["u%d" % i for i in range(16)]
    data = {u: rng.rand(int(n), 80).astype(np.float32) for u, n in zip(lst, rng.randint(60, 100, size=16))}
    acc = {u: int(rng.randint(0, 8)) for u in lst}
    trans = {u: [int(v) for v in rng.randint(3, 38, size=rng.randint(1, 4))] for u in lst}
    gkw = dict(ctc_enable=True, ar_enable=True, disc_enable=True, batch_size=8, data_dct=data, accent_dct=acc, trans_dct=trans,
               max_input_len=100, max_ctc_len=4, encoder_len=cfg.plan().seq_len, accent_classes=8)
    dev_x, dev_y = next(us.data_generator(lst, seed=3, **gkw))

    saved = []

    class evaluation:                                    # train.py:31-35
        def on_epoch_end(self, epoch, logs=None):
            p = str(tmp_path / ("%03d.h5" % epoch))
            model.save(p)
            saved.append((p, dict(logs)))

    hist = train_model.fit_generator(generator=us.data_generator(lst, seed=1, **gkw), steps_per_epoch=4, epochs=3, callbacks=[evaluation()],
                                     initial_epoch=0, validation_data=(dev_x, dev_y), max_queue_size=20, verbose=0)
    assert len(hist) == 3 and len(saved) == 3
    assert hist[-1]["loss"] < hist[0]["loss"], [h["loss"] for h in hist]
    assert any(k.startswith("val_") for k in hist[0]) and all(np.isfinite(v) for h in hist for v in h.values())
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m2, _ = mdl.SAR_Net((100, 80, 1), seed=77, raw_model=saved[-1][0], **kw)
        want = model.predict(dev_x, batch_size=8)
        got = m2.predict(dev_x, batch_size=8)
    for g, w in zip(got, want):
        g, w = (g.cpu().numpy() if hasattr(g, "cpu") else g), (w.cpu().numpy() if hasattr(w, "cpu") else w)
        assert np.array_equal(g, w)


@pytest.mark.parametrize("flags", [dict(train_ds=True), dict(train_resnet=True, train_ctc=True)])
def test_graph_replay_of_the_training_step_is_bitwise_the_eager_step(cuda_device, flags):
    """train_on_batch captures the step (forward, backward, Adam with lr_t from device memory) as a CUDA graph after two
    eager steps and replays it: over 6 steps on DIFFERENT batches the losses and every parameter / Adam moment / BN moving
    average are bitwise those of a trainer that never leaves eager mode."""
    import warnings
    from aesrc2020_b200 import model as mdl, training as T, utils as us
    kw = dict(ctc_enable=True, disc_enable=True, res_type="res18", res_filters=8, mto="gvlad", vlad_clusters=8, ghost_clusters=2,
              metric_loss="arcface", margin=0.3, bpe_classes=40, max_ctc_len=4)
    runs = []
    for use_graph in (False, True):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            model, _ = mdl.SAR_Net((100, 80, 1), seed=21, **kw)
            tr = T.HeadTrainer(model, lr=0.004, **flags)
            tr.use_graph = use_graph
            hist = []
            for i in range(6):
                x, y = us.synthetic_batch(model.config, 6, seed=400 + i)
                hist.append(tr.train_on_batch(x, y))
        assert (len(tr._graphs) == 1) == use_graph and tr.iterations == 6
        runs.append((hist, {k: v.clone() for k, v in tr.p.items()}, {k: v.clone() for k, v in tr.m.items()}))
    (h0, p0, m0), (h1, p1, m1) = runs
    assert h0 == h1, (h0, h1)
    for k in p0:
        assert torch.equal(p0[k], p1[k]), k
    for k in m0:
        assert torch.equal(m0[k], m1[k]), k


@pytest.mark.parametrize("mto", ["bigru", "avg"])
def test_reference_default_integration_trains(cuda_device, mto):
    """train.py's own defaults: MANY_TO_ONE = 'bigru' (AR_MERGE Bi-GRU, final states), METRIC_LOSS = 'softmax', CTC + accent +
    disc -- and 'avg'.  One step of HeadTrainer(train_ds=True) on CRNN_LN features vs the float64 autograd oracle (every gradient),
    then the whole model (train_resnet + train_ctc) through train_on_batch incl. the graph replay: the loss falls."""
    import warnings
    from aesrc2020_b200 import model as mdl, training as T, utils as us
    kw = dict(ctc_enable=True, disc_enable=True, res_type="res18", res_filters=8, mto=mto, metric_loss="softmax", margin=0.3,
              bpe_classes=40, max_ctc_len=4)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model, _ = mdl.SAR_Net((100, 80, 1), seed=5, **kw)
    rng = np.random.RandomState(3)
    for k in list(model.weights):
        if k.endswith("/bias") or k.endswith("/beta"):
            model.weights[k] = (model.weights[k] + rng.randn(*model.weights[k].shape) * 0.05).astype(np.float32)
    params = {k: np.asarray(v, np.float32).astype(np.float64) for k, v in model.weights.items()}
    tr = T.HeadTrainer(model, lr=0.01, train_ds=True)
    assert set(TO.pool_keys(mto)) <= set(tr.keys)
    B, S = 6, 9
    lab = rng.randint(0, 8, B)
    crnn = (rng.randn(B, S, 512) + (np.eye(8)[lab] @ rng.randn(8, 512))[:, None, :] * 0.5).astype(np.float32)
    onehot = np.eye(8, dtype=np.float32)[lab]
    pool = dict(mto=mto, vlad_clusters=0, ghost_clusters=0, train_ds=True)
    l2k = set(TO.l2_keys(True, "softmax")) | set(TO.pool_l2_keys(mto)) | {"AR_DS/kernel", "AR_DS/bias"}
    _, _, l_or, g_or = TO.train_step(dict(params), {}, crnn, onehot, lr=0.01, iterations=0, disc_enable=True, metric_loss="softmax",
                                     margin=0.3, w_accent=tr.w_acc, w_disc=tr.w_disc, pool=pool)
    got = tr.step_on_features(dev(crnn), dev(onehot))
    assert abs(got["loss_disc"] - l_or["loss_disc"]) < 2e-4 * max(1, abs(l_or["loss_disc"]))
    for k in tr.keys:
        # zero-gradient keys (a per-channel constant in front of a batch-statistic BN): AR_BN1/beta, AR_EMBEDDING/bias, and with
        # average pooling AR_DS_LN/beta too (it shifts every frame, hence the mean, by the same vector -- AR_BN1 removes it)
        if k in ("AR_BN1/beta", "AR_EMBEDDING/bias") or (mto == "avg" and k == "AR_DS_LN/beta"):
            continue
        want = g_or[k] - (2 * TO.L2_REG * params[k] if k in l2k else 0.0)
        got_k = tr.last_grads[k].cpu().numpy().astype(np.float64)
        err = float(np.max(np.abs(got_k - want)) / max(np.max(np.abs(want)), 1e-6))
        assert err < 2e-3, (k, err)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model2, train_model = mdl.SAR_Net((100, 80, 1), seed=5, lr=0.004, **kw)
        x, y = us.synthetic_batch(model2.config, 8, seed=9)
        hist = [train_model.train_on_batch(x, y) for _ in range(12)]
    assert min(h["loss"] for h in hist[-3:]) < 0.85 * hist[0]["loss"], [round(h["loss"], 3) for h in hist]
    assert len(model2.trainer()._graphs) == 1
