"""Float64 CPU restatement of the SAR-Net forward (TEST INFRASTRUCTURE ONLY).

Every function cites the reference file:line (under /root/reference) it
restates.  Items tagged [KERAS-SEMANTICS] restate documented behaviour of
Keras 2.2.x / TF 1.13 / keras_layer_normalization, whose sources are not part
of the reference tree (see oracle/__init__.py for the parity status).

Tensors are torch CPU tensors, channels-last exactly like the reference
(resnet.py:12-14).  `dtype` defaults to float64 (the parity oracle); bench.py's
cpu_baseline leg calls the same code in float32 with all host threads.

Weights are a flat dict  canonical-name -> array  in Keras layouts:
  conv kernel HWIO, dense kernel (in, out), GRU kernel (Din, 3u) / recurrent
  (u, 3u) / bias (6u,) with gate order z|r|h, BN gamma/beta/moving_mean/
  moving_variance, LN gamma/beta, VLAD centers (K+G, D), margin head W (D, n).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3            # [KERAS-SEMANTICS] BatchNormalization default epsilon
LN_EPS = 1e-14           # [KERAS-SEMANTICS] keras_layer_normalization: K.epsilon()**2
L2_EPS = 1e-12           # [KERAS-SEMANTICS] K.l2_normalize / tf.nn.l2_normalize epsilon
K_EPS = 1e-7             # [KERAS-SEMANTICS] K.epsilon()
FACE_S = 30.0            # losses.py:13,59,106 (model.py never overrides s)
CIRCLE_GAMMA = 256.0     # model.py:355

RES_REPS = {"res18": [2, 2, 2, 2], "res34": [3, 4, 6, 3]}   # resnet.py:175,193


# --------------------------------------------------------------------------
# shape rules
# --------------------------------------------------------------------------
def same_pad(n_in: int, k: int, s: int) -> Tuple[int, int, int]:
    """[KERAS-SEMANTICS] TF 'SAME': out=ceil(in/s); extra padding goes AFTER."""
    n_out = -(-n_in // s)
    total = max((n_out - 1) * s + k - n_in, 0)
    return n_out, total // 2, total - total // 2


def cal_descriptors(T: int, D: int) -> int:
    """utils.py:156-159 -- the reference's own known-answer for S (=114 @1200x80)."""
    def pool(x):
        return math.ceil(x / 2)
    t, d = T, D
    for _ in range(5):
        t, d = pool(t), pool(d)
    return int(t * d)


def _t(w, dtype):
    if isinstance(w, torch.Tensor):
        return w.to(dtype=dtype, device="cpu")
    return torch.as_tensor(np.asarray(w), dtype=dtype)


# --------------------------------------------------------------------------
# ResNet primitives (resnet.py)
# --------------------------------------------------------------------------
def conv2d(x, kernel, bias, stride: int, padding: str):
    """Keras Conv2D on NHWC input with HWIO kernel, use_bias=True
    (resnet.py:39-42,60-63,82-87 -- no use_bias argument => bias on).
    [KERAS-SEMANTICS] cross-correlation, TF SAME/VALID padding."""
    kh, kw, cin, cout = kernel.shape
    B, H, W, C = x.shape
    assert C == cin
    xc = x.permute(0, 3, 1, 2)                       # NCHW for torch
    if padding == "same":
        _, pt, pb = same_pad(H, kh, stride)
        _, pl, pr = same_pad(W, kw, stride)
        xc = F.pad(xc, (pl, pr, pt, pb))
    else:
        assert padding == "valid"
    w = kernel.permute(3, 2, 0, 1).contiguous()      # OIHW
    y = F.conv2d(xc, w, bias, stride=stride)
    return y.permute(0, 2, 3, 1).contiguous()


def batchnorm(x, p, prefix):
    """[KERAS-SEMANTICS] BatchNormalization inference on the last axis:
    gamma*(x-moving_mean)/sqrt(moving_var+1e-3)+beta (resnet.py:25, model.py:29-30)."""
    g, b = p[prefix + "/gamma"], p[prefix + "/beta"]
    m, v = p[prefix + "/moving_mean"], p[prefix + "/moving_variance"]
    return (x - m) / torch.sqrt(v + BN_EPS) * g + b


def bn_relu(x, p, prefix):
    """resnet.py:22-26."""
    return torch.relu(batchnorm(x, p, prefix))


def maxpool_same(x, k: int = 3, s: int = 2):
    """MaxPooling2D(3x3, strides 2, 'same') (resnet.py:174,192).
    [KERAS-SEMANTICS] padded cells never win (-inf padding)."""
    B, H, W, C = x.shape
    _, pt, pb = same_pad(H, k, s)
    _, pl, pr = same_pad(W, k, s)
    xc = F.pad(x.permute(0, 3, 1, 2), (pl, pr, pt, pb), value=float("-inf"))
    return F.max_pool2d(xc, k, s).permute(0, 2, 3, 1).contiguous()


def basic_block(x, p, name, filters, stride, first_of_first):
    """resnet.py:105-125 (basic_block) + :67-89 (_shortcut)."""
    if first_of_first:                                # resnet.py:111-117
        c1 = conv2d(x, p[name + "/conv1/kernel"], p[name + "/conv1/bias"], stride, "same")
    else:                                             # resnet.py:119-120
        c1 = conv2d(bn_relu(x, p, name + "/bn1"), p[name + "/conv1/kernel"],
                    p[name + "/conv1/bias"], stride, "same")
    res = conv2d(bn_relu(c1, p, name + "/bn2"), p[name + "/conv2/kernel"],
                 p[name + "/conv2/bias"], 1, "same")  # resnet.py:122
    # _shortcut (resnet.py:73-89): strides from rounded shape ratios, 1x1 'valid' conv on
    # the RAW block input when shape or channels differ, else identity.
    sh = int(round(x.shape[1] / res.shape[1]))
    sw = int(round(x.shape[2] / res.shape[2]))
    if sh > 1 or sw > 1 or x.shape[3] != res.shape[3]:
        assert sh == sw
        short = conv2d(x, p[name + "/short/kernel"], p[name + "/short/bias"], sh, "valid")
    else:
        short = x
    return short + res


def resnet(x, p, res_type: str, filters: int):
    """resnet18_ (resnet.py:170-182) / resnet34_ (resnet.py:188-201).
    x: (B, T, 80, 1) -> (B, H', W', 8*filters)."""
    if res_type not in RES_REPS:
        raise NotImplementedError("res50/101/152 return Model objects and cannot be wired "
                                  "into SAR_Net (resnet.py:217,233,249 vs model.py:252)")
    f0 = filters if res_type == "res18" else 64      # resnet.py:173 vs :191
    assert p["resnet/stem/kernel"].shape[-1] == f0
    x = conv2d(x, p["resnet/stem/kernel"], p["resnet/stem/bias"], 2, "same")
    x = bn_relu(x, p, "resnet/stem_bn")               # _conv_bn_relu, resnet.py:28-45
    x = maxpool_same(x)
    f = filters
    for i, r in enumerate(RES_REPS[res_type]):        # resnet.py:175-177 / 193-195
        for j in range(r):                            # _residual_block, resnet.py:91-103
            stride = 2 if (j == 0 and i != 0) else 1
            x = basic_block(x, p, "resnet/s%db%d" % (i + 1, j + 1), f, stride,
                            first_of_first=(i == 0 and j == 0))
        f *= 2
    return bn_relu(x, p, "resnet/final_bn")           # resnet.py:178,196


# --------------------------------------------------------------------------
# encoder tail primitives (model.py)
# --------------------------------------------------------------------------
def dense(x, p, name, activation=None, use_bias=True):
    """DS = Dense(kernel (in,out), bias) on the last axis (model.py:35-42)."""
    y = x @ p[name + "/kernel"]
    if use_bias:
        y = y + p[name + "/bias"]
    if activation == "tanh":
        y = torch.tanh(y)
    elif activation == "relu":
        y = torch.relu(y)
    elif activation == "softmax":
        y = torch.softmax(y, dim=-1)
    else:
        assert activation is None
    return y


def layernorm(x, p, name, eps: float = LN_EPS):
    """LN = keras_layer_normalization.LayerNormalization() (model.py:32-33).
    [KERAS-SEMANTICS] mean/biased variance over the last axis, eps=K.epsilon()^2."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * p[name + "/gamma"] + p[name + "/beta"]


def gru_direction(x, kernel, rec, bias, reverse: bool):
    """One direction of CuDNNGRU (model.py:44-50).  [KERAS-SEMANTICS] reset_after
    GRU, gate order z|r|h, bias (6u,) = [bz_i, br_i, bh_i, bz_r, br_r, bh_r]:
        z = s(xWz+bz_i+hUz+bz_r)  r = s(xWr+br_i+hUr+br_r)
        hh = tanh(xWh+bh_i + r*(hUh+bh_r))   h' = z*h + (1-z)*hh,  h0 = 0.
    Returns the per-step outputs in TIME order (the backward direction is re-reversed,
    as Bidirectional does) and the final state."""
    B, S, _ = x.shape
    u = rec.shape[0]
    bi, br = bias[:3 * u], bias[3 * u:]
    xp = x @ kernel + bi                               # (B,S,3u)
    h = torch.zeros(B, u, dtype=x.dtype)
    outs = [None] * S
    order = range(S - 1, -1, -1) if reverse else range(S)
    for t in order:
        hp = h @ rec + br
        z = torch.sigmoid(xp[:, t, :u] + hp[:, :u])
        r = torch.sigmoid(xp[:, t, u:2 * u] + hp[:, u:2 * u])
        hh = torch.tanh(xp[:, t, 2 * u:] + r * hp[:, 2 * u:])
        h = z * h + (1.0 - z) * hh
        outs[t] = h
    return torch.stack(outs, dim=1), h


def bigru(x, p, name, seq: bool = True):
    """BIGRU = Bidirectional(CuDNNGRU(return_sequences=seq), merge_mode='concat')
    (model.py:44-50).  seq=False -> concat(fwd final state, bwd final state)."""
    of, hf = gru_direction(x, p[name + "/forward/kernel"], p[name + "/forward/recurrent_kernel"],
                           p[name + "/forward/bias"], reverse=False)
    ob, hb = gru_direction(x, p[name + "/backward/kernel"], p[name + "/backward/recurrent_kernel"],
                           p[name + "/backward/bias"], reverse=True)
    if seq:
        return torch.cat([of, ob], dim=-1)
    return torch.cat([hf, hb], dim=-1)


# --------------------------------------------------------------------------
# NetVLAD / GhostVLAD (VLAD.py:26-49, model.py:82-109)
# --------------------------------------------------------------------------
def l2_normalize(x, dim):
    """[KERAS-SEMANTICS] x * rsqrt(max(sum(x^2), 1e-12))."""
    ss = (x * x).sum(dim=dim, keepdim=True)
    return x / torch.sqrt(torch.clamp(ss, min=L2_EPS))


def vlad_pooling(feat, score, centers, mode: str, k_centers: int):
    """VladPooling.call written literally as VLAD.py:26-49, including the two
    (B,W,H,K+G,D) temporaries.  feat (B,1,S,D), score (B,1,S,K+G)."""
    mx = score.max(dim=-1, keepdim=True).values                    # VLAD.py:33
    e = torch.exp(score - mx)                                      # VLAD.py:34
    A = e / e.sum(dim=-1, keepdim=True)                            # VLAD.py:35
    A = A.unsqueeze(-1)                                            # VLAD.py:38
    feat_res = feat.unsqueeze(-2) - centers                        # VLAD.py:39-40
    weighted = A * feat_res                                        # VLAD.py:41
    cluster_res = weighted.sum(dim=(1, 2))                         # VLAD.py:42
    if mode == "gvlad":
        cluster_res = cluster_res[:, :k_centers, :]                # VLAD.py:44-45
    out = l2_normalize(cluster_res, -1)                            # VLAD.py:47
    return out.reshape(out.shape[0], k_centers * feat.shape[-1])   # VLAD.py:48


def vlad(x, p, aggregation: str, vlad_clusters: int, ghost_clusters: int):
    """model.py:82-109: 1x1 Conv2D (+bias) to K(+G) scores, then VladPooling."""
    pre = "vlad" if aggregation == "vlad" else "gvlad"
    score = conv2d(x, p[pre + "_center_assignment/kernel"], p[pre + "_center_assignment/bias"], 1, "same")
    return vlad_pooling(x, score, p[pre + "_pool/centers"], aggregation, vlad_clusters)


def integration(x, p, hidden_dim, mto, vlad_clusters, ghost_clusters):
    """model.py:118-139."""
    if mto == "avg":
        return x.mean(dim=1)                                       # GlobalAveragePooling1D
    if mto == "bigru":
        return bigru(x, p, "AR_MERGE", seq=False)
    if mto in ("vlad", "gvlad"):
        return vlad(x.unsqueeze(1), p, mto, vlad_clusters, ghost_clusters)   # EXPAND(axis=1)
    raise SystemExit("Please specify avg/bigru/vlad/gvlad ..")    # model.py:136-138


# --------------------------------------------------------------------------
# margin heads and losses (losses.py, model.py:142-167,344-367)
# --------------------------------------------------------------------------
def face_logits(x, W, y, kind: str, m: float, s: float = FACE_S):
    """Pre-softmax logits of SphereFace/CosFace/ArcFace (losses.py:28-44,74-88,121-141)."""
    xn = l2_normalize(x, 1)
    Wn = l2_normalize(W, 0)
    cos = xn @ Wn
    if kind == "cosface":
        target = cos - m                                           # losses.py:84
    else:
        theta = torch.acos(torch.clamp(cos, -1.0 + K_EPS, 1.0 - K_EPS))
        target = torch.cos(m * theta) if kind == "sphereface" else torch.cos(theta + m)
    logits = cos * (1 - y) + target * y
    return logits * s


def disc_head(x, p, y_onehot, loss: str, margin: float, name: str = "y_disc"):
    """disc_loss (model.py:142-167): returns (layer output, pre-softmax logits)."""
    if loss == "softmax":
        logits = x @ p[name + "/kernel"]                           # Dense, use_bias=False
        return torch.softmax(logits, dim=-1), logits
    if loss in ("sphereface", "cosface", "arcface"):
        logits = face_logits(x, p[name + "/W"], y_onehot, loss, margin)
        return torch.softmax(logits, dim=-1), logits               # losses.py:45,89,142
    if loss == "circleloss":
        # model.py:162-163: l2_normalize(x,1) then Dense(no bias, no in-graph W norm)
        cos = l2_normalize(x, 1) @ p[name + "/kernel"]
        return cos, cos
    raise ValueError(loss)


def circle_loss(y_true, y_pred, gamma: float = CIRCLE_GAMMA, margin: float = 0.25):
    """losses.py:157-172."""
    alpha_p = torch.relu(1 + margin - y_pred)
    alpha_n = torch.relu(y_pred + margin)
    logit = (y_true * (alpha_p * (y_pred - (1 - margin)))
             + (1 - y_true) * (alpha_n * (y_pred - margin))) * gamma
    return -(y_true * torch.log_softmax(logit, dim=-1)).sum(dim=-1)


def categorical_crossentropy(y_true, p):
    """[KERAS-SEMANTICS] p/=sum(p); clip(1e-7,1-1e-7); -sum(y*log p) (model.py:350,354)."""
    p = p / p.sum(dim=-1, keepdim=True)
    p = torch.clamp(p, K_EPS, 1 - K_EPS)
    return -(y_true * torch.log(p)).sum(dim=-1)


# --------------------------------------------------------------------------
# CTC (model.py:62-73)
# --------------------------------------------------------------------------
def ctc_batch_cost(labels, probs, in_len, lab_len):
    """[KERAS-SEMANTICS] K.ctc_batch_cost: log(p+1e-7) fed to tf.nn.ctc_loss which
    re-applies softmax => per-frame distribution q=(p+1e-7)/sum(p+1e-7); blank = C-1;
    merge_repeated.  labels (B,Lmax) float, probs (B,S,C), in_len/lab_len (B,1) int.
    Returns (B,1) negative log-likelihoods.  Infeasible labels raise (TF raises)."""
    B, S, C = probs.shape
    blank = C - 1
    logq = torch.log_softmax(torch.log(probs + K_EPS), dim=-1)
    out = torch.zeros(B, 1, dtype=probs.dtype)
    for b in range(B):
        T = int(in_len[b].item() if hasattr(in_len[b], "item") else in_len[b])
        L = int(lab_len[b].item() if hasattr(lab_len[b], "item") else lab_len[b])
        lab = [int(v) for v in labels[b, :L].tolist()]
        assert all(0 <= v < blank for v in lab), "labels must be < num_classes-1"
        ext = [blank]
        for v in lab:
            ext += [v, blank]
        n = len(ext)
        lq = logq[b, :T][:, ext]                                    # (T, n)
        ninf = float("-inf")
        alpha = torch.full((n,), ninf, dtype=probs.dtype)
        alpha[0] = lq[0, 0]
        if n > 1:
            alpha[1] = lq[0, 1]
        skip = torch.zeros(n, dtype=torch.bool)
        for s in range(2, n):
            skip[s] = ext[s] != blank and ext[s] != ext[s - 2]
        for t in range(1, T):
            a1 = torch.cat([torch.full((1,), ninf, dtype=probs.dtype), alpha[:-1]])
            a2 = torch.cat([torch.full((2,), ninf, dtype=probs.dtype), alpha[:-2]])
            a2 = torch.where(skip, a2, torch.full_like(a2, ninf))
            alpha = torch.logsumexp(torch.stack([alpha, a1, a2]), dim=0) + lq[t]
        tail = alpha[-2:] if n > 1 else alpha[-1:]
        ll = torch.logsumexp(tail, dim=0)
        if not torch.isfinite(ll):
            raise ValueError("Not enough time for target transition sequence (infeasible CTC label)")
        out[b, 0] = -ll
    return out


# --------------------------------------------------------------------------
# the whole forward (model.py:204-371)
# --------------------------------------------------------------------------
def ctc_greedy_decode(probs, input_len: int):
    """ctc_pred(), model.py:385-389: K.ctc_decode(pred, [input_len]*n, greedy=True)[0][0] -- tf.nn.ctc_greedy_decoder
    with merge_repeated=True on log(pred): per frame the first maximum over the classes, consecutive repeats merged,
    then blanks (C-1) dropped; returned dense (n, Lmax) int64 padded with -1 (sparse_to_dense default_value=-1)."""
    pr = probs.detach().cpu().numpy() if hasattr(probs, "detach") else np.asarray(probs)
    n, S, C = pr.shape
    T = max(0, min(int(input_len), S))
    rows = []
    for b in range(n):
        best = pr[b, :T].argmax(-1)                      # np.argmax: first maximum
        prev, seq = -1, []
        for v in best:
            if v != prev and v != C - 1:
                seq.append(int(v))
            prev = v
        rows.append(seq)
    L = max([len(r) for r in rows] + [1])
    out = -np.ones((n, L), dtype=np.int64)
    for i, r in enumerate(rows):
        out[i, :len(r)] = r
    return out


def sar_net_forward(params: Dict[str, np.ndarray], inputs: Dict[str, np.ndarray], *,
                    ctc_enable=False, ar_enable=True, disc_enable=False, res_type="res18",
                    res_filters=64, hidden_dim=256, bn_dim=0, bpe_classes=1000, accent_classes=8,
                    max_ctc_len=72, mto=None, vlad_clusters=8, ghost_clusters=2,
                    metric_loss="cosface", margin=0.3, dtype=torch.float64,
                    return_intermediates=False) -> Dict[str, torch.Tensor]:
    """SAR_Net forward in inference mode (model.predict): model.py:229-339.
    Returns a dict with the model outputs (y_accent, y_disc, y_ctc_loss, y_disc_bn)
    plus, for parity diagnostics, 'embedding', 'y_accent_logits', 'y_disc_logits'."""
    p = {k: _t(v, dtype) for k, v in params.items()}
    x = _t(inputs["x_data"], dtype)
    out: Dict[str, torch.Tensor] = {}
    cnn = resnet(x, p, res_type, res_filters)                              # model.py:240-251
    B = cnn.shape[0]
    if return_intermediates:
        out["resnet"] = cnn
    cnn = cnn.reshape(B, -1, cnn.shape[-1])                                # CNN2SEQ, model.py:252
    cnn = layernorm(dense(cnn, p, "CNN_LIN", "tanh"), p, "CNN_LIN_LN")     # model.py:253-254
    crnn = layernorm(bigru(cnn, p, "CRNN"), p, "CRNN_LN")                  # model.py:255-256
    if return_intermediates:
        out["cnn_lin"] = cnn
        out["crnn"] = crnn
    if ctc_enable:                                                         # model.py:261-269
        asr = layernorm(bigru(crnn, p, "CTC_BIGRU"), p, "CTC_BIGRU_LN")
        asr = layernorm(dense(asr, p, "CTC_DS", "tanh"), p, "CTC_DS_LN")
        probs = dense(asr, p, "ctc_pred", "softmax")
        out["ctc_pred"] = probs
        out["y_ctc_loss"] = ctc_batch_cost(_t(inputs["x_ctc_label"], dtype), probs,
                                           np.asarray(inputs["x_ctc_in_len"]).reshape(-1),
                                           np.asarray(inputs["x_ctc_out_len"]).reshape(-1))
    if ar_enable:                                                          # model.py:275-296
        ar = layernorm(dense(crnn, p, "AR_DS", "tanh"), p, "AR_DS_LN")
        if return_intermediates:
            out["ar_ds"] = ar
        ar = integration(ar, p, hidden_dim, mto, vlad_clusters, ghost_clusters)
        if return_intermediates:
            out["integration"] = ar
        ar = batchnorm(ar, p, "AR_BN1")
        ar = dense(ar, p, "AR_EMBEDDING", None)
        ar = batchnorm(ar, p, "AR_BN2")
        out["embedding"] = ar
        h = dense(dense(ar, p, "AR_CF_DS1", "relu"), p, "AR_CF_DS2", "relu")
        logits = dense(h, p, "y_accent", None)
        out["y_accent_logits"] = logits
        out["y_accent"] = torch.softmax(logits, dim=-1)
        if disc_enable:                                                    # model.py:301-307
            y = _t(inputs["x_accent"], dtype)
            out["y_disc"], out["y_disc_logits"] = disc_head(ar, p, y, metric_loss, margin, "y_disc")
        if disc_enable and bn_dim:                                         # model.py:312-322
            bn = dense(ar, p, "AR_BN_DS", "relu")
            bn = batchnorm(bn, p, "AR_BN3")
            bn = dense(bn, p, "bottleneck", None)
            bn = batchnorm(bn, p, "AR_BN4")
            out["y_disc_bn"], _ = disc_head(bn, p, _t(inputs["x_accent"], dtype), metric_loss,
                                            margin, "y_disc_bn")
    return out


def loss_weights(ctc_enable, ar_enable, disc_enable, bn_dim):
    """model.py:344-367 incl. the double assignment of y_ctc_loss (second wins)."""
    alpha, beta = 0.4, 0.01
    w = {}
    if ar_enable:
        w["y_accent"] = beta if disc_enable else 1.0
        if disc_enable:
            w["y_disc"] = 1 - alpha if ctc_enable else 1.0
    if ctc_enable:
        w["y_ctc_loss"] = 1 - alpha if not disc_enable else beta
    if bn_dim:
        w["y_disc_bn"] = 0.1
    return w


def sar_net_losses(outputs, targets_onehot, *, ctc_enable, ar_enable, disc_enable, bn_dim,
                   metric_loss, margin):
    """Keras compile()'d losses/metrics evaluated on a batch (model.py:344-367):
    per-output batch means, accuracies and the weighted total (no regularisers)."""
    res = {}
    w = loss_weights(ctc_enable, ar_enable, disc_enable, bn_dim)
    y = targets_onehot
    if ar_enable:
        res["y_accent_loss"] = categorical_crossentropy(y, outputs["y_accent"]).mean()
        res["y_accent_acc"] = (outputs["y_accent"].argmax(-1) == y.argmax(-1)).double().mean()
        if disc_enable:
            if metric_loss == "circleloss":
                l = circle_loss(y, outputs["y_disc"], CIRCLE_GAMMA, margin)
            else:
                l = categorical_crossentropy(y, outputs["y_disc"])
            res["y_disc_loss"] = l.mean()
            res["y_disc_acc"] = (outputs["y_disc"].argmax(-1) == y.argmax(-1)).double().mean()
    if ctc_enable:
        res["y_ctc_loss_loss"] = outputs["y_ctc_loss"].mean()
    if bn_dim and "y_disc_bn" in outputs:
        res["y_disc_bn_loss"] = categorical_crossentropy(y, outputs["y_disc_bn"]).mean()   # Q6
    total = 0.0
    for k, wk in w.items():
        total = total + wk * res[k + "_loss"]
    res["loss"] = total
    return res
