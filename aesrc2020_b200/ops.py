"""One thin Python wrapper per C-ABI entry point (include/sarnet.h).  Tensors in, tensors out;
outputs are allocated here with torch (caller-owned device memory at the boundary)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _shim
from ._shim import check, ptr, stream_ptr

ACT = {None: 0, "none": 0, "linear": 0, "relu": 1, "tanh": 2}
HEAD = {None: 0, "none": 0, "softmax": 1, "sphereface": 2, "cosface": 3, "arcface": 4, "circleloss": 5,
        "circle_raw": 6}
FACE_S = 30.0            # losses.py:13,59,106
CIRCLE_GAMMA = 256.0     # model.py:355
LN_EPS = 1e-14           # keras_layer_normalization default (K.epsilon()**2)

# kernels launched through the C ABI since import (bench.py reports the delta as gpu_launches)
LAUNCHES = {"n": 0}


def _count(k: int = 1):
    LAUNCHES["n"] += k


def _f32(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise _shim.SarnetError("expected float32 tensor, got %s" % t.dtype)
    return t.contiguous()


def conv2d(x, w_hwio, bias=None, *, stride=1, pad_t=0, pad_l=0, out_hw: Tuple[int, int],
           pre=None, post=None, residual=None, act=None, out=None):
    """sar_conv2d_fwd.  x (B,H,W,Cin) NHWC, w (kh,kw,Cin,Cout); pre/post = (scale, shift)."""
    x = _f32(x)
    B, H, W, Cin = x.shape
    kh, kw, cin2, Cout = w_hwio.shape
    assert cin2 == Cin, (cin2, Cin)
    Ho, Wo = out_hw
    if out is None:
        out = torch.empty((B, Ho, Wo, Cout), device=x.device, dtype=torch.float32)
    ps, pt = pre if pre is not None else (None, None)
    qs, qt = post if post is not None else (None, None)
    check(_shim.lib().sar_conv2d_fwd(ptr(x), ptr(w_hwio), ptr(bias), ptr(ps), ptr(pt), ptr(qs), ptr(qt),
                                     ptr(residual), ptr(out), B, H, W, Cin, Ho, Wo, Cout, kh, kw,
                                     stride, pad_t, pad_l, ACT[act], stream_ptr()), "sar_conv2d_fwd")
    _count(1)
    return out


def dense(x, kernel, bias=None, *, act=None, pre=None, post=None, out=None):
    """Dense on the last axis as a 1x1 convolution: x (..., Din) @ kernel (Din, Dout)."""
    x = _f32(x)
    lead = x.shape[:-1]
    M = 1
    for d in lead:
        M *= int(d)
    Din, Dout = kernel.shape
    y = conv2d(x.reshape(1, M, 1, Din), kernel.reshape(1, 1, Din, Dout), bias, out_hw=(M, 1),
               pre=pre, post=post, act=act,
               out=None if out is None else out.reshape(1, M, 1, Dout))
    return y.reshape(*lead, Dout)


def maxpool2d(x, *, k=3, stride=2, pad_t=0, pad_l=0, out_hw: Tuple[int, int]):
    x = _f32(x)
    B, H, W, Cc = x.shape
    Ho, Wo = out_hw
    out = torch.empty((B, Ho, Wo, Cc), device=x.device, dtype=torch.float32)
    check(_shim.lib().sar_maxpool2d_fwd(ptr(x), ptr(out), B, H, W, Cc, Ho, Wo, k, stride, pad_t, pad_l,
                                        stream_ptr()), "sar_maxpool2d_fwd")
    _count(1)
    return out


def affine_relu(x, scale, shift, relu=True):
    x = _f32(x)
    Cc = x.shape[-1]
    out = torch.empty_like(x)
    check(_shim.lib().sar_affine_relu_fwd(ptr(x), ptr(scale), ptr(shift), ptr(out), x.numel() // Cc, Cc,
                                          1 if relu else 0, stream_ptr()), "sar_affine_relu_fwd")
    _count(1)
    return out


def layernorm(x, gamma, beta, eps: float = LN_EPS, *, planes=None, want_dense: bool = True):
    """sar_layernorm_fwd / sar_layernorm_planes_fwd.  `planes` (tc.Planes of a (1, B, S) map, C channels): also
    write the rows as fp16 hi/lo flat-pad planes for a following tensor-core Dense; want_dense=False skips the
    fp32 output (returns None)."""
    x = _f32(x)
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    out = torch.empty_like(x) if want_dense or planes is None else None
    if planes is None:
        check(_shim.lib().sar_layernorm_fwd(ptr(x), ptr(gamma), ptr(beta), ptr(out), rows, Cc,
                                            float(eps), stream_ptr()), "sar_layernorm_fwd")
    else:
        assert planes.C == Cc and planes.B == 1 and planes.H * planes.W == rows and not planes.split
        seg = planes.W if planes.rows != rows else 0       # flat-pad (1, B, S) map: a pad row per S rows; plain rows: none
        check(_shim.lib().sar_layernorm_planes_fwd(ptr(x), ptr(gamma), ptr(beta), ptr(out), ptr(planes.t), planes.rows,
                                                   seg, rows, Cc, float(eps), stream_ptr()),
              "sar_layernorm_planes_fwd")
    _count(1)
    return out


def bigru(xp, rec, rbias, *, seq=True, nb: int = 0):
    """xp (B,S,2,3u) input projections; rec (2,u,3u); rbias (2,3u).  `nb`: utterances per cluster of the Bi-GRU
    kernel (0 = the kernel's choice, 16 or 32; 32 occupies half the SMs while several batches share the GPU)."""
    xp = _f32(xp)
    B, S = xp.shape[0], xp.shape[1]
    u = rec.shape[1]
    out = torch.empty((B, S, 2 * u) if seq else (B, 2 * u), device=xp.device, dtype=torch.float32)
    check(_shim.lib().sar_bigru_nb_fwd(ptr(xp), ptr(rec), ptr(rbias), ptr(out), B, S, u, 1 if seq else 0,
                                       int(nb), stream_ptr()), "sar_bigru_fwd")
    _count(1)
    return out


def vlad(feat, w_assign, b_assign, centers, K: int, G: int, score=None, *, planes=None, want_dense: bool = True):
    """feat (B,S,D) -> (B, K*D).  Scores from (w_assign, b_assign) or a given `score` (B,S,K+G).
    `planes` (tc.Planes of a (1, B, 1) map with K*D channels, see tc.alloc_rows): also write the descriptor as
    fp16 hi/lo planes for the tensor-core embedding GEMM; want_dense=False skips the fp32 output."""
    feat = _f32(feat)
    B, S, D = feat.shape
    assert centers.shape == (K + G, D)
    if score is None:
        assert w_assign.shape == (D, K + G)
    else:
        score = _f32(score)
        assert tuple(score.shape) == (B, S, K + G) and w_assign is None and b_assign is None
    out = torch.empty((B, K * D), device=feat.device, dtype=torch.float32) if (want_dense or planes is None) else None
    if planes is not None:
        assert tuple(planes.t.shape) == (2, B, K * D), (tuple(planes.t.shape), (2, B, K * D))
    check(_shim.lib().sar_vlad_planes_fwd(ptr(feat), ptr(w_assign), ptr(b_assign), ptr(score), ptr(centers), ptr(out),
                                          ptr(planes.t) if planes is not None else None,
                                          B, S, D, K, G, stream_ptr()), "sar_vlad_fwd")
    _count(1)
    return out


def softmax_rows(x, classes: Optional[int] = None):
    """sar_softmax_rows_fwd: softmax over the first `classes` columns of the last axis -> (..., classes)."""
    x = _f32(x)
    ld = int(x.shape[-1])
    Cc = int(classes) if classes else ld
    rows = x.numel() // ld
    out = torch.empty(tuple(x.shape[:-1]) + (Cc,), device=x.device, dtype=torch.float32)
    check(_shim.lib().sar_softmax_rows_fwd(ptr(x), ld, ptr(out), rows, Cc, stream_ptr()), "sar_softmax_rows_fwd")
    _count(1)
    return out


def avgpool(x):
    x = _f32(x)
    B, S, D = x.shape
    out = torch.empty((B, D), device=x.device, dtype=torch.float32)
    check(_shim.lib().sar_avgpool_fwd(ptr(x), ptr(out), B, S, D, stream_ptr()), "sar_avgpool_fwd")
    _count(1)
    return out


def gemm_splitk(a, w, bias=None):
    a = _f32(a)
    M, K = a.shape
    N = w.shape[1]
    l = _shim.lib()
    nbytes = l.sar_gemm_splitk_workspace_bytes(M, K, N)
    ws = torch.empty((nbytes // 4,), device=a.device, dtype=torch.float32)
    out = torch.empty((M, N), device=a.device, dtype=torch.float32)
    check(l.sar_gemm_splitk_fwd(ptr(a), ptr(w), ptr(bias), ptr(out), M, K, N, ptr(ws), nbytes, stream_ptr()),
          "sar_gemm_splitk_fwd")
    _count(2)
    return out


def head(emb, cls_w=None, *, emb_d=None, wd=None, onehot=None, n_classes=8, head_kind=None,
         margin=0.3, s=FACE_S, gamma=CIRCLE_GAMMA, want_logits=True, out=None):
    """sar_head_fwd.  cls_w = (w1,b1,w2,b2,w3,b3) or None.  Returns a dict of tensors; `out` may hold
    preallocated (B, n) / (B, 4) tensors under the same names (rows of a larger batch buffer)."""
    B = (emb if emb is not None else emb_d).shape[0]
    dev = (emb if emb is not None else emb_d).device
    D = emb.shape[1] if emb is not None else 0
    n = n_classes
    hk = HEAD[head_kind]
    res = {}
    out = out or {}

    def new(name, cols):
        t = out.get(name)
        if t is None:
            return torch.empty((B, cols), device=dev, dtype=torch.float32)
        assert tuple(t.shape) == (B, cols) and t.is_contiguous() and t.dtype == torch.float32, name
        return t
    w1 = b1 = w2 = b2 = w3 = b3 = None
    H1 = H2 = 0
    if cls_w is not None:
        w1, b1, w2, b2, w3, b3 = cls_w
        H1, H2 = w1.shape[1], w2.shape[1]
        res["y_accent"] = new("y_accent", n)
        if want_logits:
            res["y_accent_logits"] = new("y_accent_logits", n)
    if hk:
        res["y_disc"] = new("y_disc", n)
        if want_logits:
            res["y_disc_logits"] = new("y_disc_logits", n)
    res["sample_stats"] = new("sample_stats", 4)
    check(_shim.lib().sar_head_fwd(ptr(emb), D, ptr(w1), ptr(b1), H1, ptr(w2), ptr(b2), H2, ptr(w3), ptr(b3),
                                   ptr(emb_d), emb_d.shape[1] if emb_d is not None else 0, ptr(wd),
                                   ptr(onehot), n, hk, float(margin), float(s), float(gamma),
                                   ptr(res.get("y_accent")), ptr(res.get("y_accent_logits")),
                                   ptr(res.get("y_disc")), ptr(res.get("y_disc_logits")),
                                   ptr(res["sample_stats"]), B, stream_ptr()), "sar_head_fwd")
    _count(1)
    return res


def ctc(logits, labels, in_len, lab_len, *, want_probs=False, loss=None, status=None, classes=None):
    """logits (B,S,C) pre-softmax; labels (B,Lmax) float32; lens (B,) or (B,1) int32.  `classes` < logits.shape[-1]:
    the rows are padded (tensor-core ctc_pred) and only the first `classes` columns count."""
    logits = _f32(logits)
    labels = _f32(labels)
    B, S, ld = logits.shape
    Cc = int(classes) if classes else ld
    in_len = in_len.reshape(-1).to(torch.int32).contiguous()
    lab_len = lab_len.reshape(-1).to(torch.int32).contiguous()
    if loss is None:
        loss = torch.empty((B,), device=logits.device, dtype=torch.float32)
    if status is None:
        status = torch.empty((B,), device=logits.device, dtype=torch.int32)
    assert loss.numel() == B and status.numel() == B and loss.is_contiguous() and status.is_contiguous()
    probs = torch.empty((B, S, Cc), device=logits.device, dtype=torch.float32) if want_probs else None
    check(_shim.lib().sar_ctc_ld_fwd(ptr(logits), ld, ptr(labels), ptr(in_len), ptr(lab_len), ptr(loss), ptr(probs),
                                     ptr(status), B, S, Cc, labels.shape[1], stream_ptr()), "sar_ctc_fwd")
    _count(1)
    return loss, status, probs


def ctc_greedy(logits, in_len=None, *, fixed_len: int = 0, classes=None):
    """sar_ctc_greedy_fwd: logits (B,S,ld) pre-softmax -> (dec (B,S) int32 padded with -1, dec_len (B,) int32)."""
    logits = _f32(logits)
    B, S, ld = logits.shape
    Cc = int(classes) if classes else ld
    if in_len is not None:
        in_len = in_len.reshape(-1).to(torch.int32).contiguous()
    dec = torch.empty((B, S), device=logits.device, dtype=torch.int32)
    dec_len = torch.empty((B,), device=logits.device, dtype=torch.int32)
    check(_shim.lib().sar_ctc_greedy_fwd(ptr(logits), ld, ptr(in_len), int(fixed_len), ptr(dec), ptr(dec_len), B, S, Cc,
                                         stream_ptr()), "sar_ctc_greedy_fwd")
    _count(1)
    return dec, dec_len


def loss_reduce(sample_stats=None, ctc_loss=None, bn_stats=None, B: Optional[int] = None):
    ref = sample_stats if sample_stats is not None else (ctc_loss if ctc_loss is not None else bn_stats)
    B = B or ref.shape[0]
    out8 = torch.empty((8,), device=ref.device, dtype=torch.float32)
    check(_shim.lib().sar_loss_reduce_fwd(ptr(sample_stats), ptr(ctc_loss), ptr(bn_stats), ptr(out8), B,
                                          stream_ptr()), "sar_loss_reduce_fwd")
    _count(1)
    return out8


def feat_batch(feats, frame_offsets, T: int):
    """sar_feat_batch_fwd: feats (total frames, D) fp32 + frame_offsets (B+1) int64 -> x_data (B,T,D)."""
    feats = _f32(feats)
    B = frame_offsets.numel() - 1
    D = int(feats.shape[1])
    x = torch.empty((B, T, D), device=feats.device, dtype=torch.float32)
    check(_shim.lib().sar_feat_batch_fwd(ptr(feats), ptr(frame_offsets), ptr(x), B, T, D, stream_ptr()), "sar_feat_batch_fwd")
    _count(1)
    return x


def labels_pack(accent=None, n_classes=0, trans=None, trans_offsets=None, Lmax=0, encoder_len=0):
    """sar_labels_pack_fwd -> dict with x_accent (B,n) and/or x_ctc_label (B,Lmax) f32, x_ctc_out_len / x_ctc_in_len
    (B,1) int32, and `status` (1,) int32."""
    ref = accent if accent is not None else trans_offsets
    dev = ref.device
    B = accent.numel() if accent is not None else trans_offsets.numel() - 1
    out = {"status": torch.zeros((1,), device=dev, dtype=torch.int32)}
    onehot = lab = olen = ilen = None
    if accent is not None:
        onehot = out["x_accent"] = torch.empty((B, n_classes), device=dev, dtype=torch.float32)
    if trans_offsets is not None:
        lab = out["x_ctc_label"] = torch.empty((B, Lmax), device=dev, dtype=torch.float32)
        olen = out["x_ctc_out_len"] = torch.empty((B, 1), device=dev, dtype=torch.int32)
        ilen = out["x_ctc_in_len"] = torch.empty((B, 1), device=dev, dtype=torch.int32)
    check(_shim.lib().sar_labels_pack_fwd(ptr(accent), int(n_classes), ptr(onehot), ptr(trans), ptr(trans_offsets), int(Lmax),
                                          int(encoder_len), ptr(lab), ptr(olen), ptr(ilen), ptr(out["status"]), B,
                                          stream_ptr()), "sar_labels_pack_fwd")
    _count(1)
    return out


def fbank(wav, offsets, melfb_t, Fmax: int, T: int, out=None, feat_ws=None):
    """wav (N,) concatenated utterances -- float32 in [-1, 1) or int16 PCM -- offsets (B+1,) int64 -> x_data (B,T,80).
    `out` / `feat_ws`: preallocated (B,T,80) / (B,Fmax,80) fp32 (a graph's static buffers)."""
    if wav.dtype not in (torch.float32, torch.int16):
        raise _shim.SarnetError("fbank: wav must be float32 or int16, got %s" % wav.dtype)
    wav = wav.contiguous()
    B = offsets.numel() - 1
    feat = feat_ws if feat_ws is not None else torch.empty((B, Fmax, 80), device=wav.device, dtype=torch.float32)
    x = out if out is not None else torch.empty((B, T, 80), device=wav.device, dtype=torch.float32)
    assert tuple(feat.shape) == (B, Fmax, 80) and x.numel() == B * T * 80 and x.is_contiguous()
    fn = _shim.lib().sar_fbank_pcm16_fwd if wav.dtype == torch.int16 else _shim.lib().sar_fbank_fwd
    check(fn(ptr(wav), ptr(offsets), ptr(melfb_t), ptr(feat), ptr(x), B, Fmax, T, stream_ptr()), "sar_fbank_fwd")
    _count(2)
    return x, feat
