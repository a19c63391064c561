"""Training mode: optimisation steps of the reference's model on the device, built slice by slice from the
accent head down to the ResNet -- `HeadTrainer(model, train_resnet=True, train_ctc=True)` trains the WHOLE model, the other
flags freeze everything below the chosen layer (fine-tuning above the frozen inference engine).

What the reference does with the path is `train_model.fit_generator(...)` (train.py:38-44) on a model compiled with
`Adam(lr, decay=2e-4)` (model.py:187-201).  This module builds the first slice of that: the layers AFTER
`integration(...)` -- AR_BN1 -> AR_EMBEDDING -> AR_BN2 -> AR_CF_DS1 -> AR_CF_DS2 -> y_accent and the y_disc margin head
(model.py:286-296, 142-167) -- run in TRAINING mode (batch-statistic BatchNormalization per replica, the l2(1e-4)
regularisers of DS, the loss weights of model.py:344-367), are differentiated by hand-written backward kernels
(csrc/train.cu through the C ABI) and updated with Keras' Adam; with `torch.distributed` initialised the gradients are
all-reduced (mean) over the replicas -- the path's first bandwidth-relevant collective (4.2 M parameters at K = 64).
The encoder in front (ResNet, Bi-GRU, AR_DS, VLAD) is the frozen inference engine and supplies `integ`.

`HeadTrainer(model).train_on_batch(x)` mirrors Keras' `train_on_batch`: returns the weighted total and the per-output
losses; `fit_generator(generator, steps_per_epoch, epochs)` loops it over utils.data_generator batches.

Second slice, `HeadTrainer(model, train_pool=True)`: the NetVLAD / GhostVLAD pooling layer (model.py:82-109, VLAD.py:26-49)
is trained too -- the frozen encoder ends at AR_DS_LN, `sar_vlad_train_fwd` keeps the soft assignments, and the backward
(`sar_l2norm_bwd` per cluster, `sar_vlad_train_bwd`, two contractions) yields the gradients of the centers and of the
assignment Conv2D (kernel, bias; l2(1e-4) regularisers).

Third slice, `HeadTrainer(model, train_ds=True)`: AR_DS (Dense + tanh) and AR_DS_LN are trained as well -- the whole accent
branch above the shared CRNN encoder (model.py:275-296).  The frozen encoder ends at CRNN_LN; `sar_vlad_train_bwd` also returns
d loss / d descriptors, `sar_ln_train_bwd` differentiates LayerNormalization together with the tanh in front of it.
Fourth slice, `HeadTrainer(model, train_crnn=True)`: CNN_LIN (Dense + tanh) -> CNN_LIN_LN -> CRNN (Bidirectional CuDNNGRU) ->
CRNN_LN are trained too (model.py:252-256) -- everything above the ResNet on the accent path.  The Bi-GRU runs in training
mode as one GEMM + `sar_gru_gate_fwd` per time step (gates kept) and is differentiated by back-propagation through time
(`sar_gru_gate_bwd` + one GEMM per step, the weight gradients as three GEMMs over all steps).
Fifth slice, `HeadTrainer(model, train_ctc=True)`: the CTC branch -- CTC_BIGRU -> CTC_BIGRU_LN -> CTC_DS -> CTC_DS_LN -> ctc_pred ->
K.ctc_batch_cost (model.py:261-269, 62-71; `sar_ctc_grad_fwd`: alpha-beta recursion, gradient through the double normalisation) --
is trained together with the accent branch, their gradients joining at CRNN_LN: multi-task training of everything above the ResNet.
Sixth slice, `HeadTrainer(model, train_resnet=True[, train_ctc=True])`: the ResNet as well (training_resnet.py: training-mode
BatchNormalization, convolution / max-pool backward) -- the whole model trains end to end, as `fit_generator` does in the
reference (train.py:38-44).  This last slice is correctness-first fp32 CUDA-core code, not yet the tensor-core backward.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch

from . import _shim, ops
from ._shim import check, ptr, stream_ptr

BN_EPS, BN_MOMENTUM = 1e-3, 0.99            # Keras BatchNormalization defaults
L2_REG = 1e-4                               # DS(..., rgr=l2(1e-4)), model.py:35-42
ADAM_B1, ADAM_B2, ADAM_EPS, ADAM_DECAY = 0.9, 0.999, 1e-7, 2e-4      # Adam(lr, decay=2e-4), model.py:197

TRAINABLE = ("AR_BN1/gamma", "AR_BN1/beta", "AR_EMBEDDING/kernel", "AR_EMBEDDING/bias", "AR_BN2/gamma", "AR_BN2/beta",
             "AR_CF_DS1/kernel", "AR_CF_DS1/bias", "AR_CF_DS2/kernel", "AR_CF_DS2/bias", "y_accent/kernel", "y_accent/bias")
L2_KEYS = {"AR_EMBEDDING/kernel", "AR_EMBEDDING/bias", "AR_CF_DS1/kernel", "AR_CF_DS1/bias", "AR_CF_DS2/kernel",
           "AR_CF_DS2/bias", "y_accent/kernel", "y_accent/bias"}


# ---------------------------------------------------------------------------------------- thin wrappers (one per entry point)
def gemm(a, b, *, ta=False, tb=False, alpha=1.0, beta=0.0, out=None):
    M = a.shape[1] if ta else a.shape[0]
    K = a.shape[0] if ta else a.shape[1]
    N = b.shape[0] if tb else b.shape[1]
    assert (b.shape[1] if tb else b.shape[0]) == K, (a.shape, b.shape, ta, tb)
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=torch.float32)
    check(_shim.lib().sar_gemm_fwd(ptr(a), ptr(b), ptr(out), M, N, K, int(ta), int(tb), float(alpha), float(beta), stream_ptr()),
          "sar_gemm_fwd")
    ops._count(1)
    return out


ROWS_PARALLEL = 2048        # more rows than this: the row-parallel kernels (the ResNet's maps: rows = B*H*W)


def _row_chunks(rows: int, C: int, dev):
    nch = max(1, min(256, rows // 512))
    return nch, torch.empty(((2 * nch + 2) * C,), device=dev, dtype=torch.float32)


def bn_train_fwd(x, gamma, beta, mov_mean=None, mov_var=None):
    rows, C = x.shape
    y = torch.empty_like(x)
    mean = torch.empty((C,), device=x.device, dtype=torch.float32)
    inv = torch.empty((C,), device=x.device, dtype=torch.float32)
    if rows > ROWS_PARALLEL:
        nch, ws = _row_chunks(rows, C, x.device)
        check(_shim.lib().sar_bn_train_rows_fwd(ptr(x), ptr(gamma), ptr(beta), ptr(mov_mean), ptr(mov_var), ptr(y), ptr(mean), ptr(inv),
                                                rows, C, BN_EPS, BN_MOMENTUM, nch, ptr(ws), stream_ptr()), "sar_bn_train_rows_fwd")
        ops._count(6)
        return y, mean, inv
    check(_shim.lib().sar_bn_train_fwd(ptr(x), ptr(gamma), ptr(beta), ptr(mov_mean), ptr(mov_var), ptr(y), ptr(mean), ptr(inv),
                                       rows, C, BN_EPS, BN_MOMENTUM, stream_ptr()), "sar_bn_train_fwd")
    ops._count(1)
    return y, mean, inv


def bn_train_bwd(x, dy, gamma, mean, inv, want_dx=True):
    rows, C = x.shape
    dx = torch.empty_like(x) if want_dx else None
    dg = torch.empty((C,), device=x.device, dtype=torch.float32)
    db = torch.empty((C,), device=x.device, dtype=torch.float32)
    if rows > ROWS_PARALLEL:
        nch, ws = _row_chunks(rows, C, x.device)
        check(_shim.lib().sar_bn_train_rows_bwd(ptr(x), ptr(dy), ptr(gamma), ptr(mean), ptr(inv), ptr(dx), ptr(dg), ptr(db), rows, C, nch,
                                                ptr(ws), stream_ptr()), "sar_bn_train_rows_bwd")
        ops._count(4)
        return dx, dg, db
    check(_shim.lib().sar_bn_train_bwd(ptr(x), ptr(dy), ptr(gamma), ptr(mean), ptr(inv), ptr(dx), ptr(dg), ptr(db), rows, C,
                                       stream_ptr()), "sar_bn_train_bwd")
    ops._count(1)
    return dx, dg, db


def bias_act(x, bias, relu=False, tanh=False):
    y = torch.empty_like(x)
    check(_shim.lib().sar_bias_act_fwd(ptr(x), ptr(bias), ptr(y), x.shape[0], x.shape[1], 1 if relu else (2 if tanh else 0), stream_ptr()),
          "sar_bias_act_fwd")
    ops._count(1)
    return y


def relu_bwd(g, h):
    out = torch.empty_like(g)
    check(_shim.lib().sar_relu_bwd(ptr(g), ptr(h), ptr(out), g.numel(), stream_ptr()), "sar_relu_bwd")
    ops._count(1)
    return out


def colsum(g):
    out = torch.empty((g.shape[1],), device=g.device, dtype=torch.float32)
    if g.shape[0] > ROWS_PARALLEL:
        nch, ws = _row_chunks(g.shape[0], g.shape[1], g.device)
        check(_shim.lib().sar_colsum_rows_fwd(ptr(g), ptr(out), g.shape[0], g.shape[1], nch, ptr(ws), stream_ptr()), "sar_colsum_rows_fwd")
        ops._count(2)
        return out
    check(_shim.lib().sar_colsum_fwd(ptr(g), ptr(out), g.shape[0], g.shape[1], stream_ptr()), "sar_colsum_fwd")
    ops._count(1)
    return out


def l2norm_fwd(v, axis):
    out = torch.empty_like(v)
    inv = torch.empty((v.shape[0] if axis else v.shape[1],), device=v.device, dtype=torch.float32)
    check(_shim.lib().sar_l2norm_fwd(ptr(v), ptr(out), ptr(inv), v.shape[0], v.shape[1], int(axis), stream_ptr()), "sar_l2norm_fwd")
    ops._count(1)
    return out, inv


def l2norm_bwd(vhat, inv, u, axis, out=None, beta=0.0):
    if out is None:
        out = torch.empty_like(vhat)
    check(_shim.lib().sar_l2norm_bwd(ptr(vhat), ptr(inv), ptr(u), ptr(out), vhat.shape[0], vhat.shape[1], int(axis), float(beta),
                                     stream_ptr()), "sar_l2norm_bwd")
    ops._count(1)
    return out


def head_grad(z_acc, c_disc, onehot, head_kind, margin, w_acc, w_disc, s=ops.FACE_S, gamma=ops.CIRCLE_GAMMA):
    B, n = onehot.shape
    g_acc = torch.empty((B, n), device=onehot.device, dtype=torch.float32) if z_acc is not None else None
    g_disc = torch.empty((B, n), device=onehot.device, dtype=torch.float32) if c_disc is not None else None
    losses = torch.zeros((B, 2), device=onehot.device, dtype=torch.float32)
    check(_shim.lib().sar_head_grad_fwd(ptr(z_acc), ptr(c_disc), ptr(onehot), n, ops.HEAD[head_kind], float(margin), float(s),
                                        float(gamma), float(w_acc), float(w_disc), ptr(g_acc), ptr(g_disc), ptr(losses), B,
                                        stream_ptr()), "sar_head_grad_fwd")
    ops._count(1)
    return g_acc, g_disc, losses


def vlad_train_fwd(x, w_assign, b_assign, centers, K: int, G: int):
    """x (B,S,D) -> soft assignments A (B,S,K+G), un-normalised residual sums R (B,K,D), asum (B,K)."""
    B, S, D = x.shape
    A = torch.empty((B, S, K + G), device=x.device, dtype=torch.float32)
    R = torch.empty((B, K, D), device=x.device, dtype=torch.float32)
    asum = torch.empty((B, K), device=x.device, dtype=torch.float32)
    check(_shim.lib().sar_vlad_train_fwd(ptr(x), ptr(w_assign), ptr(b_assign), ptr(centers), ptr(A), ptr(R), ptr(asum), B, S, D, K, G,
                                         stream_ptr()), "sar_vlad_train_fwd")
    ops._count(1)
    return A, R, asum


def vlad_train_bwd(x, A, centers, gR, asum, K: int, G: int, want_gx: bool = False):
    """-> g_scores (B,S,K+G) = d loss / d assignment scores, gc_part (B,K,D) (sum over B = gradient of the K real centers)
    [, g_x (B,S,D): the residual-sum path of d loss / d x; the caller adds g_scores Wa^T]."""
    B, S, D = x.shape
    g_scores = torch.empty((B, S, K + G), device=x.device, dtype=torch.float32)
    gc_part = torch.empty((B, K, D), device=x.device, dtype=torch.float32)
    g_x = torch.empty((B, S, D), device=x.device, dtype=torch.float32) if want_gx else None
    check(_shim.lib().sar_vlad_train_bwd(ptr(x), ptr(A), ptr(centers), ptr(gR), ptr(asum), ptr(g_scores), ptr(gc_part), ptr(g_x),
                                         B, S, D, K, G, stream_ptr()), "sar_vlad_train_bwd")
    ops._count(1)
    return (g_scores, gc_part, g_x) if want_gx else (g_scores, gc_part)


LN_EPS = 1e-14                              # keras_layer_normalization: K.epsilon() ** 2


def ln_train_bwd(y, gamma, g_z, tanh_in: bool):
    """Backward of LayerNormalization (optionally with the tanh that produced its input y): -> (g_pre, g_z * xhat)."""
    rows, C = y.shape
    g_pre, gzx = torch.empty_like(y), torch.empty_like(y)
    check(_shim.lib().sar_ln_train_bwd(ptr(y), ptr(gamma), ptr(g_z), ptr(g_pre), ptr(gzx), rows, C, LN_EPS, int(tanh_in), stream_ptr()),
          "sar_ln_train_bwd")
    ops._count(1)
    return g_pre, gzx


def gru_dir_fwd(x_rows, B: int, S: int, W, U, bias, reverse: bool, out, off: int):
    """One direction of CuDNNGRU in training mode (model.py:44-50): the input projection is one GEMM, every time step one
    GEMM (h_prev U) + `sar_gru_gate_fwd`.  Writes h_t into out[:, t, off:off+u] (`out` None: return_sequences=False, the final
    state is saved["hseq"][S]); returns what the backward needs."""
    u = U.shape[0]
    dev = x_rows.device
    xp = bias_act(gemm(x_rows, W), bias[:3 * u])                              # (B*S, 3u) = x W + b_i
    hseq = torch.zeros((S + 1, B, u), device=dev, dtype=torch.float32)        # hseq[k]: state before processing step k
    z, r, hh, hph = (torch.empty((S, B, u), device=dev, dtype=torch.float32) for _ in range(4))
    b_r = bias[3 * u:]
    lib = _shim.lib()
    for k in range(S):
        t = S - 1 - k if reverse else k
        hu = gemm(hseq[k], U)
        check(lib.sar_gru_gate_fwd(ptr(xp), ptr(hu), ptr(b_r), ptr(hseq[k]), ptr(z[k]), ptr(r[k]), ptr(hh[k]), ptr(hph[k]),
                                   ptr(hseq[k + 1]), ptr(out), B, S, u, t, out.shape[-1] if out is not None else 0, off, stream_ptr()),
              "sar_gru_gate_fwd")
    ops._count(S)
    return dict(hseq=hseq, z=z, r=r, hh=hh, hph=hph, x_rows=x_rows, reverse=reverse, off=off)


def gru_dir_bwd(g_out, saved, B: int, S: int, W, U, g_x=None, dh_last=None):
    """Back-propagation through time of gru_dir_fwd: g_out (B,S,2u) = d loss / d layer output (None for
    return_sequences=False), dh_last (B,u) = d loss / d final state.  Returns the gradients of (kernel, recurrent_kernel,
    bias (6u)) and accumulates d loss / d x into g_x (B*S, Din)."""
    u = U.shape[0]
    dev = saved["hseq"].device
    hseq, reverse, off = saved["hseq"], saved["reverse"], saved["off"]
    d_xp = torch.empty((B * S, 3 * u), device=dev, dtype=torch.float32)
    d_hu = torch.empty((S, B, 3 * u), device=dev, dtype=torch.float32)
    lib = _shim.lib()
    dh = dh_last
    for k in range(S - 1, -1, -1):
        t = S - 1 - k if reverse else k
        dh_prev = torch.empty((B, u), device=dev, dtype=torch.float32)
        check(lib.sar_gru_gate_bwd(ptr(g_out), ptr(dh) if dh is not None else None, ptr(saved["z"][k]), ptr(saved["r"][k]),
                                   ptr(saved["hh"][k]), ptr(saved["hph"][k]), ptr(hseq[k]), ptr(d_xp), ptr(d_hu[k]), ptr(dh_prev),
                                   B, S, u, t, g_out.shape[-1] if g_out is not None else 0, off, stream_ptr()), "sar_gru_gate_bwd")
        if k > 0:
            gemm(d_hu[k], U, tb=True, out=dh_prev, beta=1.0)                    # + d_hu U^T
        dh = dh_prev
    ops._count(S)
    x_rows = saved["x_rows"]
    gW = gemm(x_rows, d_xp, ta=True)
    gU = gemm(hseq[:S].view(S * B, u), d_hu.view(S * B, 3 * u), ta=True)
    gb = torch.cat([colsum(d_xp), colsum(d_hu.view(S * B, 3 * u))])
    if g_x is None:
        g_x = gemm(d_xp, W, tb=True)
    else:
        gemm(d_xp, W, tb=True, out=g_x, beta=1.0)
    return gW, gU, gb, g_x


def ctc_grad(logits, labels, in_len, lab_len, scale: float, classes=None):
    """sar_ctc_grad_fwd: K.ctc_batch_cost per utterance and scale * d loss_b / d (pre-softmax ctc_pred logits) (B,S,C)."""
    B, S, ld = logits.shape
    Cc = int(classes) if classes else ld
    labels = labels.to(torch.float32).contiguous()
    in_len = in_len.reshape(-1).to(torch.int32).contiguous()
    lab_len = lab_len.reshape(-1).to(torch.int32).contiguous()
    loss = torch.empty((B,), device=logits.device, dtype=torch.float32)
    status = torch.empty((B,), device=logits.device, dtype=torch.int32)
    grad = torch.empty((B, S, Cc), device=logits.device, dtype=torch.float32)
    check(_shim.lib().sar_ctc_grad_fwd(ptr(logits), ld, ptr(labels), ptr(in_len), ptr(lab_len), ptr(loss), ptr(grad), ptr(status),
                                       B, S, Cc, labels.shape[1], float(scale), stream_ptr()), "sar_ctc_grad_fwd")
    ops._count(1)
    return loss, grad, status


def adam_step(p, g, m, v, lr_t, l2=0.0):
    check(_shim.lib().sar_adam_fwd(ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), float(lr_t), ADAM_B1, ADAM_B2, ADAM_EPS, float(l2),
                                   stream_ptr()), "sar_adam_fwd")
    ops._count(1)


def adam_step_dev(p, g, m, v, lr_dev, l2=0.0):
    """adam_step with lr_t in a one-element device tensor (graph replay, see HeadTrainer.train_on_batch)."""
    check(_shim.lib().sar_adam_dev_fwd(ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), ptr(lr_dev), ADAM_B1, ADAM_B2, ADAM_EPS, float(l2),
                                       stream_ptr()), "sar_adam_dev_fwd")
    ops._count(1)


def unit_norm(w):
    check(_shim.lib().sar_unit_norm_fwd(ptr(w), w.shape[0], w.shape[1], stream_ptr()), "sar_unit_norm_fwd")
    ops._count(1)


def adam_lr_t(lr: float, iterations: int) -> float:
    """Keras 2.2.4 Adam.get_updates: lr / (1 + decay * iterations) * sqrt(1 - b2^t) / (1 - b1^t), t = iterations + 1."""
    t = iterations + 1
    return lr / (1.0 + ADAM_DECAY * iterations) * float(np.sqrt(1.0 - ADAM_B2 ** t) / (1.0 - ADAM_B1 ** t))


# ---------------------------------------------------------------------------------------- the trainer
class HeadTrainer:
    """Fine-tunes the accent head of a SARModel on the device (module docstring)."""

    def __init__(self, model, lr: float = 0.01, group=None, train_pool: bool = False, train_ds: bool = False,
                 train_crnn: bool = False, train_ctc: bool = False, train_resnet: bool = False):
        """train_pool: also train the NetVLAD / GhostVLAD pooling layer (assignment Conv2D + centers, model.py:82-109):
        the frozen encoder then ends at AR_DS_LN and the step differentiates vlad() as well (second slice).
        train_ds (implies train_pool): also train AR_DS (Dense + tanh, l2 regularisers) and AR_DS_LN (model.py:275-276):
        the whole accent branch above the shared CRNN encoder; the frozen encoder ends at CRNN_LN (third slice).
        train_crnn (implies train_ds): also train CNN_LIN (Dense + tanh) -> CNN_LIN_LN -> CRNN (Bidirectional CuDNNGRU, back-
        propagation through time) -> CRNN_LN (model.py:252-256): everything above the ResNet on the accent path; the
        frozen encoder is the ResNet alone (fourth slice).
        train_ctc (implies train_crnn; needs ctc_enable): the CTC branch as well -- CTC_BIGRU -> CTC_BIGRU_LN -> CTC_DS (Dense +
        tanh) -> CTC_DS_LN -> ctc_pred -> K.ctc_batch_cost (model.py:261-269, 62-71), its gradient joining the accent
        branch's at CRNN_LN: multi-task training of everything above the ResNet (fifth slice).
        train_resnet (implies train_crnn): the ResNet too, in training mode (batch-statistic BatchNormalization) with its
        backward (training_resnet.ResNetTrainer) -- nothing is frozen: the whole model trains end to end (sixth slice;
        with train_ctc: the reference's full multi-task fit)."""
        cfg = model.config
        if not cfg.ar_enable:
            raise ValueError("HeadTrainer needs ar_enable=True")
        if train_ctc and not cfg.ctc_enable:
            raise ValueError("train_ctc needs a model built with ctc_enable=True")
        train_crnn = bool(train_crnn or train_ctc or train_resnet)
        train_ds = bool(train_ds or train_crnn)
        train_pool = bool(train_pool or train_ds)
        if train_pool and cfg.mto not in ("avg", "bigru", "vlad", "gvlad"):
            raise ValueError("train_pool needs mto in avg | bigru | vlad | gvlad (got %r)" % cfg.mto)
        self.model, self.cfg, self.lr, self.group = model, cfg, float(lr), group
        self.train_pool = bool(train_pool)
        self.train_ds = bool(train_ds)
        self.train_crnn = bool(train_crnn)
        self.train_ctc = bool(train_ctc)
        self.train_resnet = bool(train_resnet)
        self.iterations = 0
        self.head_kind = cfg.metric_loss if cfg.disc_enable else None
        self.disc_key = None
        if cfg.disc_enable:
            self.disc_key = "y_disc/W" if cfg.metric_loss in ("sphereface", "cosface", "arcface") else "y_disc/kernel"
        self.keys: List[str] = list(TRAINABLE) + ([self.disc_key] if self.disc_key else [])
        self.l2 = set(L2_KEYS) | ({"y_disc/kernel"} if (cfg.disc_enable and cfg.metric_loss == "softmax") else set())
        self.pool_keys: List[str] = []
        if self.train_pool:
            pre = cfg.mto
            if pre in ("vlad", "gvlad"):
                self.pool_keys = [pre + "_center_assignment/kernel", pre + "_center_assignment/bias", pre + "_pool/centers"]
                self.l2 |= set(self.pool_keys[:2])       # l2(1e-4) on the assignment kernel and bias; none on the centers
            elif pre == "bigru":                         # integration(): BIGRU(hidden_dim, seq=False, name="AR_MERGE"), model.py:118-123
                self.pool_keys = ["AR_MERGE/%s/%s" % (d, w) for d in ("forward", "backward") for w in ("kernel", "recurrent_kernel", "bias")]
                self.l2 |= {k for k in self.pool_keys if not k.endswith("recurrent_kernel")}
            self.keys += self.pool_keys                  # (avg: GlobalAveragePooling1D has no weights)
        self.ds_keys: List[str] = []
        if self.train_ds:
            self.ds_keys = ["AR_DS/kernel", "AR_DS/bias", "AR_DS_LN/gamma", "AR_DS_LN/beta"]
            self.keys += self.ds_keys
            self.l2 |= set(self.ds_keys[:2])             # DS(hidden_dim, 'tanh'): l2(1e-4) on kernel and bias (model.py:35-42)
        self.crnn_keys: List[str] = []
        if self.train_crnn:
            gk = ["CRNN/%s/%s" % (d, w) for d in ("forward", "backward") for w in ("kernel", "recurrent_kernel", "bias")]
            self.crnn_keys = ["CNN_LIN/kernel", "CNN_LIN/bias", "CNN_LIN_LN/gamma", "CNN_LIN_LN/beta"] + gk + ["CRNN_LN/gamma", "CRNN_LN/beta"]
            self.keys += self.crnn_keys
            # DS: l2(1e-4) on kernel and bias; BIGRU: kernel_regularizer + bias_regularizer, none on the recurrent kernel (model.py:35-50)
            self.l2 |= {"CNN_LIN/kernel", "CNN_LIN/bias"} | {k for k in gk if not k.endswith("recurrent_kernel")}
        self.resnet_keys: List[str] = []
        res_stats: List[str] = []
        if self.train_resnet:
            from .training_resnet import ResNetTrainer
            self.resnet_keys, res_l2, res_stats = ResNetTrainer.param_keys(cfg)
            self.keys += self.resnet_keys
            self.l2 |= set(res_l2)
        self.ctc_keys: List[str] = []
        if self.train_ctc:
            gk = ["CTC_BIGRU/%s/%s" % (d, w) for d in ("forward", "backward") for w in ("kernel", "recurrent_kernel", "bias")]
            self.ctc_keys = gk + ["CTC_BIGRU_LN/gamma", "CTC_BIGRU_LN/beta", "CTC_DS/kernel", "CTC_DS/bias", "CTC_DS_LN/gamma",
                                  "CTC_DS_LN/beta", "ctc_pred/kernel", "ctc_pred/bias"]
            self.keys += self.ctc_keys
            self.l2 |= {"CTC_DS/kernel", "CTC_DS/bias", "ctc_pred/kernel", "ctc_pred/bias"} | {k for k in gk if not k.endswith("recurrent_kernel")}
        lw = cfg.loss_weights()
        self.w_acc, self.w_disc = float(lw.get("y_accent", 0.0)), float(lw.get("y_disc", 0.0))
        self.w_ctc = float(lw.get("y_ctc_loss", 0.0))
        dev = torch.device(model.device)
        put = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
        bn_stats = [b + s for b in ("AR_BN1", "AR_BN2") for s in ("/moving_mean", "/moving_variance")]
        self.p: Dict[str, torch.Tensor] = {k: put(model.weights[k]) for k in self.keys + bn_stats + res_stats}
        self.stat_keys: List[str] = bn_stats + res_stats       # BN moving averages: per-replica updates, averaged over the ranks
        self.resnet_tr = None
        if self.train_resnet:
            from .training_resnet import ResNetTrainer
            self.resnet_tr = ResNetTrainer(cfg, self.p)
        self.m = {k: torch.zeros_like(self.p[k]) for k in self.keys}
        self.v = {k: torch.zeros_like(self.p[k]) for k in self.keys}
        self.last_grads: Dict[str, torch.Tensor] = {}
        # CUDA-graph replay of the step (train_on_batch): after two eager steps of a batch shape the whole step -- forward,
        # backward, Adam -- is captured once and replayed; lr_t travels through a device scalar.  Single process only (the
        # gradient all-reduce stays eager); SAR_TRAIN_GRAPH=0 disables.
        import os as _os
        self.use_graph = _os.environ.get("SAR_TRAIN_GRAPH", "1") != "0"
        self._avg_E: Dict = {}          # mto='avg': the (B*S, B) averaging matrices
        self._graphs: Dict = {}
        self._eager_count: Dict = {}
        self._lr_dev = None            # set while a step is being captured
        self._defer = False            # capture: no host synchronisation inside the step
        self._last = None              # (losses, loss_ctc, status) device tensors of the last step

    # ---- frozen encoder: x_data -> integ (B, K*D | D | 2u), or (train_pool) the descriptors (B,S,D) in front of vlad()
    def encode(self, x) -> torch.Tensor:
        if self.train_resnet:                     # nothing is frozen: the step starts from the features themselves
            xd = self.model._as_dict(x)
            return self.model._to_device("x_data", xd["x_data"]).contiguous()
        out = self.model.forward_device(x, want_intermediates=True, graph=False)
        if self.train_crnn:                       # CNN2SEQ (model.py:252): the ResNet's (B,H,W,C) map as (B, S, Cc) rows
            plan = self.cfg.plan()
            raw = out["resnet_raw"]
            return raw.reshape(raw.shape[0], plan.seq_len, plan.cout).contiguous()
        return out["crnn" if self.train_ds else ("ar_ds" if self.train_pool else "integration")].contiguous()

    # ---- one step on (integ | descriptors, onehot) device tensors
    def step_on_features(self, integ: torch.Tensor, onehot: torch.Tensor, ctc=None) -> Dict[str, float]:
        """`ctc` (train_ctc): (x_ctc_label (B,Lmax), x_ctc_in_len (B,1), x_ctc_out_len (B,1)) device tensors."""
        p, cfg = self.p, self.cfg
        B = integ.shape[0]
        g_ctc: Dict[str, torch.Tensor] = {}
        g_crnn_ctc = loss_ctc = ctc_status = None
        pool = ds = rn = None
        res_shape = None
        if self.train_resnet:                    # ResNet in training mode; CNN2SEQ (model.py:252): (B,H,W,C) -> (B, H*W, C)
            fmap = self.resnet_tr.forward(integ)
            res_shape = fmap.shape
            integ = fmap.view(B, res_shape[1] * res_shape[2], res_shape[3])
        if self.train_crnn:                      # CNN_LIN -> CNN_LIN_LN -> CRNN -> CRNN_LN on the frozen ResNet's sequence (B,S,Cc)
            _, Sr, Cr = integ.shape
            x0 = integ.view(B * Sr, Cr)
            y_lin = bias_act(gemm(x0, p["CNN_LIN/kernel"]), p["CNN_LIN/bias"], tanh=True)
            z_lin = ops.layernorm(y_lin, p["CNN_LIN_LN/gamma"], p["CNN_LIN_LN/beta"])
            u = p["CRNN/forward/recurrent_kernel"].shape[0]
            gru_out = torch.empty((B, Sr, 2 * u), device=integ.device, dtype=torch.float32)
            sv = [gru_dir_fwd(z_lin, B, Sr, p["CRNN/%s/kernel" % d], p["CRNN/%s/recurrent_kernel" % d], p["CRNN/%s/bias" % d],
                              d == "backward", gru_out, i * u) for i, d in enumerate(("forward", "backward"))]
            crnn_rows = ops.layernorm(gru_out.view(B * Sr, 2 * u), p["CRNN_LN/gamma"], p["CRNN_LN/beta"])
            rn = (x0, y_lin, z_lin, gru_out, sv, Sr)
            integ = crnn_rows.view(B, Sr, 2 * u)
            if self.train_ctc:                   # the whole CTC branch, forward and backward, down to d loss / d CRNN_LN output
                if ctc is None:
                    raise ValueError("train_ctc needs the CTC inputs (x_ctc_label, x_ctc_in_len, x_ctc_out_len)")
                uc = p["CTC_BIGRU/forward/recurrent_kernel"].shape[0]
                ago = torch.empty((B, Sr, 2 * uc), device=integ.device, dtype=torch.float32)
                svc = [gru_dir_fwd(crnn_rows, B, Sr, p["CTC_BIGRU/%s/kernel" % d], p["CTC_BIGRU/%s/recurrent_kernel" % d],
                                   p["CTC_BIGRU/%s/bias" % d], d == "backward", ago, i * uc) for i, d in enumerate(("forward", "backward"))]
                a1 = ops.layernorm(ago.view(B * Sr, 2 * uc), p["CTC_BIGRU_LN/gamma"], p["CTC_BIGRU_LN/beta"])
                y2 = bias_act(gemm(a1, p["CTC_DS/kernel"]), p["CTC_DS/bias"], tanh=True)
                a2 = ops.layernorm(y2, p["CTC_DS_LN/gamma"], p["CTC_DS_LN/beta"])
                logits = bias_act(gemm(a2, p["ctc_pred/kernel"]), p["ctc_pred/bias"])
                Cb = logits.shape[-1]
                loss_b, g_log, status = ctc_grad(logits.view(B, Sr, Cb), ctc[0], ctc[1], ctc[2], self.w_ctc / B)
                ctc_status = status
                if not self._defer and bool((status != 0).any()):
                    raise _shim.SarnetError("CTC: infeasible or out-of-range label sequence in batch")
                loss_ctc = loss_b
                g_log = g_log.view(B * Sr, Cb)
                g_ctc["ctc_pred/kernel"], g_ctc["ctc_pred/bias"] = gemm(a2, g_log, ta=True), colsum(g_log)
                g_a2 = gemm(g_log, p["ctc_pred/kernel"], tb=True)
                g_p2, gzx = ln_train_bwd(y2, p["CTC_DS_LN/gamma"], g_a2, tanh_in=True)
                g_ctc["CTC_DS_LN/gamma"], g_ctc["CTC_DS_LN/beta"] = colsum(gzx), colsum(g_a2)
                g_ctc["CTC_DS/kernel"], g_ctc["CTC_DS/bias"] = gemm(a1, g_p2, ta=True), colsum(g_p2)
                g_a1 = gemm(g_p2, p["CTC_DS/kernel"], tb=True)
                g_ago, gzx = ln_train_bwd(ago.view(B * Sr, 2 * uc), p["CTC_BIGRU_LN/gamma"], g_a1, tanh_in=False)
                g_ctc["CTC_BIGRU_LN/gamma"], g_ctc["CTC_BIGRU_LN/beta"] = colsum(gzx), colsum(g_a1)
                for d, s_ in zip(("forward", "backward"), svc):
                    gW, gU, gb, g_crnn_ctc = gru_dir_bwd(g_ago.view(B, Sr, 2 * uc), s_, B, Sr, p["CTC_BIGRU/%s/kernel" % d],
                                                         p["CTC_BIGRU/%s/recurrent_kernel" % d], g_x=g_crnn_ctc)
                    g_ctc["CTC_BIGRU/%s/kernel" % d], g_ctc["CTC_BIGRU/%s/recurrent_kernel" % d], g_ctc["CTC_BIGRU/%s/bias" % d] = gW, gU, gb
        if self.train_ds:                        # AR_DS -> AR_DS_LN on the frozen encoder's CRNN_LN output (B,S,2u)
            crnn = integ
            _, S0, C0 = crnn.shape
            yds = bias_act(gemm(crnn.view(B * S0, C0), p["AR_DS/kernel"]), p["AR_DS/bias"], tanh=True)
            zds = ops.layernorm(yds, p["AR_DS_LN/gamma"], p["AR_DS_LN/beta"])
            ds = (crnn.view(B * S0, C0), yds)
            integ = zds.view(B, S0, -1)
        if self.train_pool and cfg.mto in ("vlad", "gvlad"):   # vlad() in training mode: integ = l2norm_k(A^T x - (sum A) c), flattened
            feat = integ
            _, S, D = feat.shape
            K, G = cfg.vlad_clusters, (cfg.ghost_clusters if cfg.mto == "gvlad" else 0)
            kw_, kb_, kc_ = self.pool_keys
            A, R, asum = vlad_train_fwd(feat, p[kw_].view(D, K + G), p[kb_], p[kc_], K, G)
            V, rinv = l2norm_fwd(R.view(B * K, D), 1)
            integ = V.view(B, K * D)
            pool = ("vlad", feat, A, asum, V, rinv, S, D, K, G)
        elif self.train_pool and cfg.mto == "bigru":           # AR_MERGE: Bi-GRU with return_sequences=False (the reference's default mto)
            feat = integ
            _, S, D = feat.shape
            rows = feat.reshape(B * S, D)
            svm = [gru_dir_fwd(rows, B, S, p["AR_MERGE/%s/kernel" % d], p["AR_MERGE/%s/recurrent_kernel" % d], p["AR_MERGE/%s/bias" % d],
                               d == "backward", None, 0) for d in ("forward", "backward")]
            integ = torch.cat([s_["hseq"][S] for s_ in svm], dim=1)             # concat(final forward state, final backward state)
            pool = ("bigru", svm, S, D, svm[0]["hseq"].shape[-1])
        elif self.train_pool:                                   # avg: GlobalAveragePooling1D as a GEMM with the (B*S, B) averaging matrix
            feat = integ
            _, S, D = feat.shape
            E = self._avg_E.get((B, S))               # built once per shape, in an eager step (never inside a graph capture)
            if E is None:
                E = torch.zeros((B, S, B), device=feat.device, dtype=torch.float32)
                ib = torch.arange(B, device=feat.device)
                E[ib, :, ib] = 1.0 / S
                E = self._avg_E[(B, S)] = E.view(B * S, B)
            integ = gemm(E, feat.reshape(B * S, D), ta=True)
            pool = ("avg", E, S, D)
        # forward, training mode
        x1, m1, i1 = bn_train_fwd(integ, p["AR_BN1/gamma"], p["AR_BN1/beta"], p["AR_BN1/moving_mean"], p["AR_BN1/moving_variance"])
        e0 = bias_act(gemm(x1, p["AR_EMBEDDING/kernel"]), p["AR_EMBEDDING/bias"])
        e, m2, i2 = bn_train_fwd(e0, p["AR_BN2/gamma"], p["AR_BN2/beta"], p["AR_BN2/moving_mean"], p["AR_BN2/moving_variance"])
        h1 = bias_act(gemm(e, p["AR_CF_DS1/kernel"]), p["AR_CF_DS1/bias"], relu=True)
        h2 = bias_act(gemm(h1, p["AR_CF_DS2/kernel"]), p["AR_CF_DS2/bias"], relu=True)
        z = bias_act(gemm(h2, p["y_accent/kernel"]), p["y_accent/bias"])
        c = xh = xinv = wh = winv = None
        kind = self.head_kind
        if kind in ("sphereface", "cosface", "arcface"):
            xh, xinv = l2norm_fwd(e, 1)
            wh, winv = l2norm_fwd(p[self.disc_key], 0)
            c = gemm(xh, wh)
        elif kind == "circleloss":
            xh, xinv = l2norm_fwd(e, 1)
            c = gemm(xh, p[self.disc_key])
        elif kind == "softmax":
            c = gemm(e, p[self.disc_key])
        g_z, g_c, losses = head_grad(z, c, onehot, kind, cfg.margin, self.w_acc, self.w_disc)
        # backward
        g: Dict[str, torch.Tensor] = {}
        g["y_accent/kernel"], g["y_accent/bias"] = gemm(h2, g_z, ta=True), colsum(g_z)
        g_h2 = relu_bwd(gemm(g_z, p["y_accent/kernel"], tb=True), h2)
        g["AR_CF_DS2/kernel"], g["AR_CF_DS2/bias"] = gemm(h1, g_h2, ta=True), colsum(g_h2)
        g_h1 = relu_bwd(gemm(g_h2, p["AR_CF_DS2/kernel"], tb=True), h1)
        g["AR_CF_DS1/kernel"], g["AR_CF_DS1/bias"] = gemm(e, g_h1, ta=True), colsum(g_h1)
        g_e = gemm(g_h1, p["AR_CF_DS1/kernel"], tb=True)
        if kind in ("sphereface", "cosface", "arcface"):
            g[self.disc_key] = l2norm_bwd(wh, winv, gemm(xh, g_c, ta=True), 0)
            l2norm_bwd(xh, xinv, gemm(g_c, wh, tb=True), 1, out=g_e, beta=1.0)
        elif kind == "circleloss":
            g[self.disc_key] = gemm(xh, g_c, ta=True)
            l2norm_bwd(xh, xinv, gemm(g_c, p[self.disc_key], tb=True), 1, out=g_e, beta=1.0)
        elif kind == "softmax":
            g[self.disc_key] = gemm(e, g_c, ta=True)
            gemm(g_c, p[self.disc_key], tb=True, out=g_e, beta=1.0)
        g_e0, g["AR_BN2/gamma"], g["AR_BN2/beta"] = bn_train_bwd(e0, g_e, p["AR_BN2/gamma"], m2, i2)
        g["AR_EMBEDDING/kernel"], g["AR_EMBEDDING/bias"] = gemm(x1, g_e0, ta=True), colsum(g_e0)
        g_x1 = gemm(g_e0, p["AR_EMBEDDING/kernel"], tb=True)
        g_integ, g["AR_BN1/gamma"], g["AR_BN1/beta"] = bn_train_bwd(integ, g_x1, p["AR_BN1/gamma"], m1, i1, want_dx=pool is not None)
        if pool is not None:
            g_feat = None
            if pool[0] == "vlad":
                _, feat, A, asum, V, rinv, S, D, K, G = pool
                kw_, kb_, kc_ = self.pool_keys
                gR = l2norm_bwd(V, rinv, g_integ.view(B * K, D), 1)                   # (B*K, D) = d loss / d R
                if ds is not None:
                    g_scores, gc_part, g_feat = vlad_train_bwd(feat, A, p[kc_], gR, asum, K, G, want_gx=True)
                    gemm(g_scores.view(B * S, K + G), p[kw_].view(D, K + G), tb=True, out=g_feat.view(B * S, D), beta=1.0)
                else:
                    g_scores, gc_part = vlad_train_bwd(feat, A, p[kc_], gR, asum, K, G)
                g[kw_] = gemm(feat.view(B * S, D), g_scores.view(B * S, K + G), ta=True).view_as(p[kw_])
                g[kb_] = colsum(g_scores.view(B * S, K + G))
                gc = torch.zeros_like(p[kc_])                                         # ghost centers: no gradient
                gc[:K] = colsum(gc_part.view(B, K * D)).view(K, D)
                g[kc_] = gc
            elif pool[0] == "bigru":
                _, svm, S, D, um = pool
                for i, (d, s_) in enumerate(zip(("forward", "backward"), svm)):
                    gW, gU, gb, g_feat = gru_dir_bwd(None, s_, B, S, p["AR_MERGE/%s/kernel" % d], p["AR_MERGE/%s/recurrent_kernel" % d],
                                                     g_x=g_feat, dh_last=g_integ[:, i * um:(i + 1) * um].contiguous())
                    g["AR_MERGE/%s/kernel" % d], g["AR_MERGE/%s/recurrent_kernel" % d], g["AR_MERGE/%s/bias" % d] = gW, gU, gb
            else:
                _, E, S, D = pool
                g_feat = gemm(E, g_integ)                                             # every frame gets 1/S of its utterance's gradient
            if ds is not None:
                x_ds, yds = ds
                g_pre, gzx = ln_train_bwd(yds, p["AR_DS_LN/gamma"], g_feat.view(B * S, D), tanh_in=True)
                g["AR_DS_LN/gamma"], g["AR_DS_LN/beta"] = colsum(gzx), colsum(g_feat.view(B * S, D))
                g["AR_DS/kernel"], g["AR_DS/bias"] = gemm(x_ds, g_pre, ta=True), colsum(g_pre)
                if rn is not None:               # ... and on through CRNN_LN, the Bi-GRU (BPTT), CNN_LIN_LN and CNN_LIN
                    x0, y_lin, z_lin, gru_out, sv, Sr = rn
                    u2 = gru_out.shape[-1]
                    if g_crnn_ctc is not None:                                          # + the CTC branch's share
                        g_crnn = gemm(g_pre, p["AR_DS/kernel"], tb=True, out=g_crnn_ctc, beta=1.0)
                    else:
                        g_crnn = gemm(g_pre, p["AR_DS/kernel"], tb=True)                # (B*S, 2u) = d loss / d CRNN_LN output
                    g_gru, gzx2 = ln_train_bwd(gru_out.view(B * Sr, u2), p["CRNN_LN/gamma"], g_crnn, tanh_in=False)
                    g["CRNN_LN/gamma"], g["CRNN_LN/beta"] = colsum(gzx2), colsum(g_crnn)
                    g_zlin = None
                    for d, s_ in zip(("forward", "backward"), sv):
                        gW, gU, gb, g_zlin = gru_dir_bwd(g_gru.view(B, Sr, u2), s_, B, Sr, p["CRNN/%s/kernel" % d],
                                                         p["CRNN/%s/recurrent_kernel" % d], g_x=g_zlin)
                        g["CRNN/%s/kernel" % d], g["CRNN/%s/recurrent_kernel" % d], g["CRNN/%s/bias" % d] = gW, gU, gb
                    g_pl, gzx1 = ln_train_bwd(y_lin, p["CNN_LIN_LN/gamma"], g_zlin, tanh_in=True)
                    g["CNN_LIN_LN/gamma"], g["CNN_LIN_LN/beta"] = colsum(gzx1), colsum(g_zlin)
                    g["CNN_LIN/kernel"], g["CNN_LIN/bias"] = gemm(x0, g_pl, ta=True), colsum(g_pl)
                    if res_shape is not None:    # ... and through the whole ResNet
                        g_x0 = gemm(g_pl, p["CNN_LIN/kernel"], tb=True)
                        g.update(self.resnet_tr.backward(g_x0.view(res_shape)))
        g.update(g_ctc)
        self._all_reduce(g)
        self.last_grads = g
        # Adam (the l2 regulariser's gradient 2 * 1e-4 * w is added inside the kernel), then the kernel constraint
        if self._lr_dev is not None:               # being captured: the step size is read from device memory at replay time
            for k in self.keys:
                adam_step_dev(p[k], g[k], self.m[k], self.v[k], self._lr_dev, l2=L2_REG if k in self.l2 else 0.0)
        else:
            lr_t = adam_lr_t(self.lr, self.iterations)
            for k in self.keys:
                adam_step(p[k], g[k], self.m[k], self.v[k], lr_t, l2=L2_REG if k in self.l2 else 0.0)
        if kind == "circleloss":
            unit_norm(p[self.disc_key])
        self._sync_stats()
        self._last = (losses, loss_ctc, ctc_status)
        if self._defer:
            return {}
        self.iterations += 1
        return self._report()

    def _report(self) -> Dict[str, float]:
        """Host copy of the last step's losses (the step's only synchronisation)."""
        losses, loss_ctc, _ = self._last
        lm = losses.mean(0).tolist()
        out = {"loss_accent": lm[0]}
        total = self.w_acc * lm[0]
        if self.head_kind:
            out["loss_disc"] = lm[1]
            total += self.w_disc * lm[1]
        if loss_ctc is not None:
            out["loss_ctc"] = float(loss_ctc.mean().item())
            total += self.w_ctc * out["loss_ctc"]
        out["loss"] = total                        # data terms (Keras adds the regulariser terms to the reported total)
        return out

    # ---- the step as a replayed CUDA graph
    def _distributed(self) -> bool:
        import torch.distributed as dist
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def step_graphed(self, feats: torch.Tensor, onehot: torch.Tensor, ctc=None) -> Dict[str, float]:
        """step_on_features through a CUDA graph: the first two steps of a batch shape run eagerly (they also raise every
        kernel's shared-memory limit and fill the allocator), the third is captured -- ~2 k launches become one replay.
        Inputs are copied into the graph's static buffers, lr_t into its device scalar.  Bitwise the eager step (tested)."""
        if not self.use_graph or self._distributed():
            return self.step_on_features(feats, onehot, ctc)
        key = (tuple(feats.shape), tuple(onehot.shape), None if ctc is None else tuple(tuple(c.shape) for c in ctc))
        st = self._graphs.get(key)
        if st is None:
            n = self._eager_count.get(key, 0)
            if n < 2:
                self._eager_count[key] = n + 1
                return self.step_on_features(feats, onehot, ctc)
            st = {"feats": feats.clone(), "onehot": onehot.clone(), "ctc": None if ctc is None else tuple(c.clone() for c in ctc),
                  "lr": torch.zeros((1,), device=feats.device, dtype=torch.float32), "graph": torch.cuda.CUDAGraph()}
            torch.cuda.synchronize()
            self._lr_dev, self._defer = st["lr"], True
            try:
                with torch.cuda.graph(st["graph"]):
                    self.step_on_features(st["feats"], st["onehot"], st["ctc"])
            finally:
                self._lr_dev, self._defer = None, False
            st["last"], st["grads"] = self._last, self.last_grads
            self._graphs[key] = st
        st["feats"].copy_(feats)
        st["onehot"].copy_(onehot)
        if ctc is not None:
            for d, c in zip(st["ctc"], ctc):
                d.copy_(c)
        st["lr"].fill_(adam_lr_t(self.lr, self.iterations))
        st["graph"].replay()
        self.iterations += 1
        self._last, self.last_grads = st["last"], st["grads"]
        status = self._last[2]
        if status is not None and bool((status != 0).any()):
            raise _shim.SarnetError("CTC: infeasible or out-of-range label sequence in batch (the step was applied with a zero "
                                    "CTC gradient for the rejected utterances)")
        return self._report()

    def _sync_stats(self):
        """Every replica updates the BN moving averages with ITS shard's statistics; the mean over the replicas keeps the ranks'
        inference weights identical (Keras' multi_gpu_model towers share one set of variables)."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return
        flat = torch.cat([self.p[k].reshape(-1) for k in self.stat_keys])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        flat /= dist.get_world_size(self.group)
        off = 0
        for k in self.stat_keys:
            n = self.p[k].numel()
            self.p[k].copy_(flat[off:off + n].view_as(self.p[k]))
            off += n

    def _all_reduce(self, grads: Dict[str, torch.Tensor]):
        """Mean of the replicas' gradients: ONE flat all-reduce (NCCL over NVLink when initialised)."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return
        flat = torch.cat([grads[k].reshape(-1) for k in self.keys])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        flat /= dist.get_world_size(self.group)
        off = 0
        for k in self.keys:
            n = grads[k].numel()
            grads[k].copy_(flat[off:off + n].view_as(grads[k]))
            off += n

    # ---- Keras-like surface
    def train_on_batch(self, x, y=None) -> Dict[str, float]:
        xd = self.model._as_dict(x)
        onehot = xd.get("x_accent") if self.cfg.disc_enable else (y or {}).get("y_accent")
        if onehot is None:
            raise ValueError("train_on_batch needs the accent labels (x_accent, or y['y_accent'])")
        onehot = self.model._to_device("x_accent", onehot).contiguous()
        ctc = None
        if self.train_ctc:
            ctc = tuple(self.model._to_device(k, xd[k]).contiguous() for k in ("x_ctc_label", "x_ctc_in_len", "x_ctc_out_len"))
        return self.step_graphed(self.encode(xd), onehot, ctc)

    def fit_generator(self, generator, steps_per_epoch: int, epochs: int = 1, verbose: int = 0):
        """train.py:38-44 shape: `epochs` x `steps_per_epoch` batches of (inputs, targets) from the generator."""
        history = []
        it = iter(generator)
        for ep in range(epochs):
            acc = []
            for _ in range(steps_per_epoch):
                item = next(it)
                x, y = item if isinstance(item, tuple) else (item, None)
                acc.append(self.train_on_batch(x, y))
            history.append({k: float(np.mean([a[k] for a in acc])) for k in acc[0]})
            if verbose:
                print("epoch %d: %s" % (ep, history[-1]))
        self.sync_to_model()
        return history

    def sync_to_model(self):
        """Write the trained parameters (and the BN moving statistics) back into model.weights; the inference engine
        is rebuilt on next use."""
        for k, t in self.p.items():
            self.model.weights[k] = t.detach().cpu().numpy().astype(np.float32)
        self.model._engine = None
