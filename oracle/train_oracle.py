"""TEST INFRASTRUCTURE -- float64 torch restatement of ONE TRAINING STEP of the reference's accent head
(first slice): the layers after `integration(...)` of model.py:286-296 / 142-167 in TRAINING mode

    integ -> AR_BN1 (batch statistics) -> AR_EMBEDDING Dense(+l2) -> AR_BN2 (batch statistics)
          -> AR_CF_DS1 relu -> AR_CF_DS2 relu -> y_accent softmax            (categorical cross-entropy)
          -> y_disc: SphereFace / CosFace / ArcFace / softmax Dense / l2norm + Dense (Circle-Loss)

with the loss wiring of model.py:344-367 (loss_weights), the l2(1e-4) kernel / bias regularisers of DS (model.py:35-42),
and `Adam(lr, decay=2e-4)` (model.py:197).  Gradients come from torch autograd on this float64 graph; the graph itself is
the oracle's own forward functions (sarnet_oracle.face_logits / circle_loss / categorical_crossentropy), and the tests
pin the gradients with central finite differences and Adam / BatchNorm against torch.optim.Adam / F.batch_norm.

Only tests/ may import this module.  [KERAS-SEMANTICS] constants (Keras 2.2.4 on TF 1.13, third-party, restated):
  * BatchNormalization(training): biased batch variance, eps 1e-3, moving = 0.99 * moving + 0.01 * batch (2-D inputs take
    the non-fused path: the moving variance is updated with the BIASED batch variance);
  * Adam: lr_t = lr / (1 + decay * iterations) * sqrt(1 - b2^t) / (1 - b1^t), t = iterations + 1, update
    p -= lr_t * m / (sqrt(v) + 1e-7) (epsilon OUTSIDE the square root, K.epsilon());
  * regulariser l2(c): c * sum(w^2) added to the loss (kernel AND bias of every DS layer);
  * unit_norm constraint (Circle-Loss head, model.py:163): W <- W / (1e-7 + ||W[:, j]||) after every update.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch

from . import sarnet_oracle as O

BN_EPS, BN_MOMENTUM = 1e-3, 0.99
L2_REG = 1e-4
ADAM_B1, ADAM_B2, ADAM_EPS, ADAM_DECAY = 0.9, 0.999, 1e-7, 2e-4

TRAINABLE = ("AR_BN1/gamma", "AR_BN1/beta", "AR_EMBEDDING/kernel", "AR_EMBEDDING/bias", "AR_BN2/gamma", "AR_BN2/beta",
             "AR_CF_DS1/kernel", "AR_CF_DS1/bias", "AR_CF_DS2/kernel", "AR_CF_DS2/bias", "y_accent/kernel", "y_accent/bias")
L2_KEYS = ("AR_EMBEDDING/kernel", "AR_EMBEDDING/bias", "AR_CF_DS1/kernel", "AR_CF_DS1/bias", "AR_CF_DS2/kernel",
           "AR_CF_DS2/bias", "y_accent/kernel", "y_accent/bias")


def disc_key(metric_loss: str) -> str:
    return "y_disc/W" if metric_loss in ("sphereface", "cosface", "arcface") else "y_disc/kernel"


def trainable_keys(disc_enable: bool, metric_loss: str):
    return list(TRAINABLE) + ([disc_key(metric_loss)] if disc_enable else [])


def l2_keys(disc_enable: bool, metric_loss: str):
    # DS(accent_classes, 'softmax', use_bias=False) carries the l2 kernel regulariser; the Face layers and the
    # Circle-Loss Dense do not (losses.py:13 regularizer=None; model.py:163 plain Dense)
    return list(L2_KEYS) + (["y_disc/kernel"] if (disc_enable and metric_loss == "softmax") else [])


def bn_train(x, gamma, beta):
    mean = x.mean(0)
    var = ((x - mean) ** 2).mean(0)
    return (x - mean) / torch.sqrt(var + BN_EPS) * gamma + beta, mean, var


def head_loss(p: Dict[str, torch.Tensor], integ: torch.Tensor, onehot: torch.Tensor, *, disc_enable: bool, metric_loss: str,
              margin: float, w_accent: float, w_disc: float):
    """-> (total loss incl. regularisers, dict of parts / batch statistics)."""
    x1, m1, v1 = bn_train(integ, p["AR_BN1/gamma"], p["AR_BN1/beta"])
    e0 = x1 @ p["AR_EMBEDDING/kernel"] + p["AR_EMBEDDING/bias"]
    e, m2, v2 = bn_train(e0, p["AR_BN2/gamma"], p["AR_BN2/beta"])
    h1 = torch.relu(e @ p["AR_CF_DS1/kernel"] + p["AR_CF_DS1/bias"])
    h2 = torch.relu(h1 @ p["AR_CF_DS2/kernel"] + p["AR_CF_DS2/bias"])
    pa = torch.softmax(h2 @ p["y_accent/kernel"] + p["y_accent/bias"], -1)
    l_acc = O.categorical_crossentropy(onehot, pa).mean()
    total = w_accent * l_acc
    parts = {"loss_accent": l_acc, "bn1_mean": m1, "bn1_var": v1, "bn2_mean": m2, "bn2_var": v2, "embedding": e}
    if disc_enable:
        out, _ = O.disc_head(e, p, onehot, metric_loss, margin)
        if metric_loss == "circleloss":
            l_d = O.circle_loss(onehot, out, gamma=O.CIRCLE_GAMMA, margin=margin).mean()
        else:
            l_d = O.categorical_crossentropy(onehot, out).mean()
        parts["loss_disc"] = l_d
        total = total + w_disc * l_d
    reg = sum(L2_REG * (p[k] ** 2).sum() for k in l2_keys(disc_enable, metric_loss))
    parts["reg"] = reg
    return total + reg, parts


# ---- second slice: the NetVLAD / GhostVLAD pooling layer is trained together with the head (frozen encoder below it)
MERGE_KEYS = ["AR_MERGE/%s/%s" % (d, w) for d in ("forward", "backward") for w in ("kernel", "recurrent_kernel", "bias")]


def pool_keys(mto: str):
    """Trainable weights of integration() (model.py:118-139): vlad / gvlad -- the 1x1 assignment Conv2D and VladPooling's centers
    (model.py:82-109, VLAD.py:17-20); bigru -- the AR_MERGE Bi-GRU (return_sequences=False); avg -- none."""
    if mto == "bigru":
        return list(MERGE_KEYS)
    if mto == "avg":
        return []
    return [mto + "_center_assignment/kernel", mto + "_center_assignment/bias", mto + "_pool/centers"]


def pool_l2_keys(mto: str):
    # vlad: kernel_regularizer = bias_regularizer = l2(1e-4) on the assignment Conv2D (model.py:87-95), the centers carry none;
    # bigru: BIGRU's kernel and bias regularisers (model.py:44-50), none on the recurrent kernel
    if mto == "bigru":
        return [k for k in MERGE_KEYS if not k.endswith("recurrent_kernel")]
    if mto == "avg":
        return []
    return [mto + "_center_assignment/kernel", mto + "_center_assignment/bias"]


DS_KEYS = ["AR_DS/kernel", "AR_DS/bias", "AR_DS_LN/gamma", "AR_DS_LN/beta"]       # third slice (model.py:275-276)
# fourth slice (model.py:252-256): CNN_LIN -> CNN_LIN_LN -> CRNN (Bidirectional CuDNNGRU) -> CRNN_LN
GRU_KEYS = ["CRNN/%s/%s" % (d, w) for d in ("forward", "backward") for w in ("kernel", "recurrent_kernel", "bias")]
CRNN_KEYS = ["CNN_LIN/kernel", "CNN_LIN/bias", "CNN_LIN_LN/gamma", "CNN_LIN_LN/beta"] + GRU_KEYS + ["CRNN_LN/gamma", "CRNN_LN/beta"]
# DS: l2(1e-4) on kernel and bias; BIGRU: kernel_regularizer and bias_regularizer, none on the recurrent kernel (model.py:35-50)
CRNN_L2_KEYS = ["CNN_LIN/kernel", "CNN_LIN/bias"] + [k for k in GRU_KEYS if not k.endswith("recurrent_kernel")]
# fifth slice (model.py:261-269): CTC_BIGRU -> CTC_BIGRU_LN -> CTC_DS -> CTC_DS_LN -> ctc_pred -> K.ctc_batch_cost
CTC_GRU_KEYS = ["CTC_BIGRU/%s/%s" % (d, w) for d in ("forward", "backward") for w in ("kernel", "recurrent_kernel", "bias")]
CTC_KEYS = CTC_GRU_KEYS + ["CTC_BIGRU_LN/gamma", "CTC_BIGRU_LN/beta", "CTC_DS/kernel", "CTC_DS/bias", "CTC_DS_LN/gamma", "CTC_DS_LN/beta",
                           "ctc_pred/kernel", "ctc_pred/bias"]
CTC_L2_KEYS = ["CTC_DS/kernel", "CTC_DS/bias", "ctc_pred/kernel", "ctc_pred/bias"] + [k for k in CTC_GRU_KEYS if not k.endswith("recurrent_kernel")]


def resnet_train(x, p, res_type: str, filters: int):
    """O.resnet (resnet.py:170-201) in TRAINING mode: every BatchNormalization of the ResNet normalises with the batch
    statistics of its input (biased variance over batch and positions, eps 1e-3) -- the only difference to inference."""
    def bn_batch(t, pp, prefix):
        mean = t.mean(dim=(0, 1, 2))
        var = ((t - mean) ** 2).mean(dim=(0, 1, 2))
        return (t - mean) / torch.sqrt(var + BN_EPS) * pp[prefix + "/gamma"] + pp[prefix + "/beta"]
    keep = O.batchnorm
    O.batchnorm = bn_batch
    try:
        return O.resnet(x, p, res_type, filters)
    finally:
        O.batchnorm = keep


def ctc_loss_autograd(probs, labels, in_len, lab_len):
    """K.ctc_batch_cost (O.ctc_batch_cost: q = softmax(log(p + 1e-7)), blank = C-1) in a form autograd can differentiate:
    torch's own CTC on log q (O.ctc_batch_cost's explicit lattice takes logsumexp over all -inf states, whose gradient is
    NaN; the two forwards agree -- tests/test_oracle_kats.py).  Returns the (B,) losses."""
    B, S, C = probs.shape
    logq = torch.log_softmax(torch.log(probs + O.K_EPS), dim=-1).transpose(0, 1)           # (S,B,C)
    il = torch.as_tensor(np.asarray(in_len).reshape(-1), dtype=torch.long)
    ll = torch.as_tensor(np.asarray(lab_len).reshape(-1), dtype=torch.long)
    lab = torch.as_tensor(np.asarray(labels), dtype=torch.long)
    return torch.nn.functional.ctc_loss(logq, lab, il, ll, blank=C - 1, reduction="none", zero_infinity=False)


def pooled_head_loss(p: Dict[str, torch.Tensor], feat: torch.Tensor, onehot: torch.Tensor, *, mto: str, vlad_clusters: int,
                     ghost_clusters: int, train_ds: bool = False, train_crnn: bool = False, train_ctc: bool = False, ctc=None,
                     w_ctc: float = 0.0, train_resnet=None, extra_keys=(), extra_l2=(), **kw):
    """feat (B,S,D) = AR_DS_LN output (frozen encoder) -> vlad() -> the head of head_loss, with the pooling layer's
    regularisers added.  train_ds: feat is the CRNN_LN output (B,S,2u) instead and AR_DS (Dense + tanh, l2 regularisers on
    kernel and bias) -> AR_DS_LN run in front of vlad() (model.py:275-276)."""
    reg_ds = 0.0
    reg_res = 0.0
    if train_resnet:         # dict(res_type=, filters=): feat is x_data (B,T,80,1); the ResNet trains too (sixth slice), then CNN2SEQ
        fmap = resnet_train(feat, p, train_resnet["res_type"], train_resnet["filters"])
        feat = fmap.reshape(fmap.shape[0], fmap.shape[1] * fmap.shape[2], fmap.shape[3])
        reg_res = sum(L2_REG * (p[k] ** 2).sum() for k in extra_l2)       # l2(1e-4) on every conv kernel (resnet.py:36,56,86,116)
        train_crnn = True
    train_crnn = train_crnn or train_ctc
    if train_crnn:           # feat is the frozen ResNet's sequence (B,S,Cc): model.py:252-256 in front of the accent branch
        feat = O.layernorm(O.dense(feat, p, "CNN_LIN", "tanh"), p, "CNN_LIN_LN")
        feat = O.layernorm(O.bigru(feat, p, "CRNN"), p, "CRNN_LN")
        reg_ds = reg_ds + sum(L2_REG * (p[k] ** 2).sum() for k in CRNN_L2_KEYS)
        train_ds = True
    ctc_term = ctc_losses = None
    if train_ctc:            # ctc = (labels (B,Lmax), in_len, lab_len): the ASR branch on the CRNN_LN output (model.py:261-269)
        asr = O.layernorm(O.bigru(feat, p, "CTC_BIGRU"), p, "CTC_BIGRU_LN")
        asr = O.layernorm(O.dense(asr, p, "CTC_DS", "tanh"), p, "CTC_DS_LN")
        ctc_losses = ctc_loss_autograd(O.dense(asr, p, "ctc_pred", "softmax"), *ctc)
        ctc_term = w_ctc * ctc_losses.mean()
        reg_ds = reg_ds + sum(L2_REG * (p[k] ** 2).sum() for k in CTC_L2_KEYS)
    if train_ds:
        feat = O.layernorm(O.dense(feat, p, "AR_DS", "tanh"), p, "AR_DS_LN")
        reg_ds = reg_ds + L2_REG * ((p["AR_DS/kernel"] ** 2).sum() + (p["AR_DS/bias"] ** 2).sum())
    integ = O.integration(feat, p, feat.shape[-1], mto, vlad_clusters, ghost_clusters)       # model.py:118-139
    total, parts = head_loss(p, integ, onehot, **kw)
    reg = sum(L2_REG * (p[k] ** 2).sum() for k in pool_l2_keys(mto)) + reg_ds + reg_res
    parts["reg"] = parts["reg"] + reg
    parts["integration"] = integ
    if ctc_term is not None:
        parts["loss_ctc"] = ctc_losses.mean()
        total = total + ctc_term
    return total + reg, parts


def adam_update(p, g, m, v, iterations: int, lr: float):
    """One Keras-2.2.4 Adam update of a tensor; returns (p, m, v)."""
    lr_t = lr * (1.0 / (1.0 + ADAM_DECAY * iterations))
    t = iterations + 1
    lr_t = lr_t * np.sqrt(1.0 - ADAM_B2 ** t) / (1.0 - ADAM_B1 ** t)
    m = ADAM_B1 * m + (1 - ADAM_B1) * g
    v = ADAM_B2 * v + (1 - ADAM_B2) * g * g
    return p - lr_t * m / (torch.sqrt(v) + ADAM_EPS), m, v


def train_step(params: Dict[str, np.ndarray], state: Dict[str, np.ndarray], integ, onehot, *, lr: float, iterations: int,
               disc_enable: bool, metric_loss: str, margin: float, w_accent: float, w_disc: float, pool: Dict = None
               ) -> Tuple[Dict[str, np.ndarray], Dict[str, np.ndarray], Dict[str, float], Dict[str, np.ndarray]]:
    """params: canonical weights (the trainable ones are updated); state: Adam moments `m/<key>`, `v/<key>`.
    `pool` = dict(mto=, vlad_clusters=, ghost_clusters=): `integ` is then the (B,S,D) descriptor tensor in front of
    vlad() and the pooling layer's weights are trained too (pooled_head_loss).
    Returns (new params incl. the BN moving statistics, new state, losses, gradients)."""
    keys = trainable_keys(disc_enable, metric_loss) + (pool_keys(pool["mto"]) if pool else []) + \
        (DS_KEYS if (pool and (pool.get("train_ds") or pool.get("train_crnn") or pool.get("train_ctc") or pool.get("train_resnet"))) else []) + \
        (CRNN_KEYS if (pool and (pool.get("train_crnn") or pool.get("train_ctc") or pool.get("train_resnet"))) else []) + list(pool.get("extra_keys", ()) if pool else ()) + (CTC_KEYS if (pool and pool.get("train_ctc")) else [])
    p = {k: torch.tensor(np.asarray(v, np.float64), requires_grad=(k in keys)) for k, v in params.items()}
    kw = dict(disc_enable=disc_enable, metric_loss=metric_loss, margin=margin, w_accent=w_accent, w_disc=w_disc)
    x_in, y_in = torch.as_tensor(np.asarray(integ, np.float64)), torch.as_tensor(np.asarray(onehot, np.float64))
    total, parts = pooled_head_loss(p, x_in, y_in, **pool, **kw) if pool else head_loss(p, x_in, y_in, **kw)
    grads = torch.autograd.grad(total, [p[k] for k in keys])
    new_p = {k: np.asarray(v, np.float64).copy() for k, v in params.items()}
    new_s = dict(state)
    g_out = {}
    for k, g in zip(keys, grads):
        m = torch.as_tensor(np.asarray(state.get("m/" + k, np.zeros_like(new_p[k])), np.float64))
        v = torch.as_tensor(np.asarray(state.get("v/" + k, np.zeros_like(new_p[k])), np.float64))
        pk, m, v = adam_update(p[k].detach(), g, m, v, iterations, lr)
        if k == "y_disc/kernel" and metric_loss == "circleloss":          # unit_norm constraint, axis 0
            pk = pk / (1e-7 + torch.sqrt((pk ** 2).sum(0, keepdim=True)))
        new_p[k], new_s["m/" + k], new_s["v/" + k] = pk.numpy(), m.numpy(), v.numpy()
        g_out[k] = g.numpy()
    for bn, mk, vk in (("AR_BN1", "bn1_mean", "bn1_var"), ("AR_BN2", "bn2_mean", "bn2_var")):
        new_p[bn + "/moving_mean"] = BN_MOMENTUM * new_p[bn + "/moving_mean"] + (1 - BN_MOMENTUM) * parts[mk].detach().numpy()
        new_p[bn + "/moving_variance"] = BN_MOMENTUM * new_p[bn + "/moving_variance"] + (1 - BN_MOMENTUM) * parts[vk].detach().numpy()
    losses = {k: float(v.detach()) for k, v in parts.items() if k.startswith("loss") or k == "reg"}
    losses["total"] = float(total.detach())
    return new_p, new_s, losses, g_out
