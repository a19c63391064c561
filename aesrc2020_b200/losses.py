"""Mirror of the reference's losses.py on device tensors.

    SphereFace / CosFace / ArcFace(n_classes=10, s=30.0, m=..., regularizer=None)   losses.py:13,59,106
        layer([x (B,D), y_onehot (B,n)]) -> softmax probabilities (B,n)            losses.py:28-47,74-91,121-144
        weight `W` (D, n_classes), glorot_uniform                                  losses.py:22-26
    circle_loss(y_true, y_pred, gamma=256, margin=0.25) -> per-sample loss (B,)     losses.py:157-172

All arithmetic runs in the fused head kernel (csrc/head.cu) through the C ABI.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import ops


class _FaceLayer:
    kind = None
    default_m = 0.0

    def __init__(self, n_classes=10, s=30.0, m=None, regularizer=None, name=None, W: Optional[np.ndarray] = None,
                 **kwargs):
        self.n_classes = int(n_classes)
        self.s = float(s)
        self.m = float(self.default_m if m is None else m)
        self.regularizer = regularizer          # training-only in the reference; ignored in the forward
        self.name = name
        self.W = None if W is None else np.ascontiguousarray(W, dtype=np.float32)
        self._dev = None

    def build(self, input_shape):
        """losses.py:20-26: W (D, n_classes), glorot_uniform (seeded here)."""
        if self.W is None:
            D = int(input_shape[0][-1])
            lim = np.sqrt(6.0 / (D + self.n_classes))
            self.W = np.random.RandomState(1234).uniform(-lim, lim, size=(D, self.n_classes)).astype(np.float32)

    def get_weights(self):
        return [self.W]

    def set_weights(self, ws):
        self.W = np.ascontiguousarray(ws[0], dtype=np.float32)
        self._dev = None

    def compute_output_shape(self, input_shape):
        return (None, self.n_classes)

    def __call__(self, inputs):
        return self.call(inputs)

    def call(self, inputs, return_logits: bool = False):
        x, y = inputs
        if self.W is None:
            self.build([tuple(x.shape), tuple(y.shape)])
        if self._dev is None or self._dev.device != x.device:
            self._dev = torch.from_numpy(self.W).to(x.device)
        r = ops.head(None, None, emb_d=x.contiguous(), wd=self._dev, onehot=y.contiguous().float(),
                     n_classes=self.n_classes, head_kind=self.kind, margin=self.m, s=self.s)
        return (r["y_disc"], r["y_disc_logits"]) if return_logits else r["y_disc"]


class SphereFace(_FaceLayer):
    kind = "sphereface"
    default_m = 1.35


class CosFace(_FaceLayer):
    kind = "cosface"
    default_m = 0.35


class ArcFace(_FaceLayer):
    kind = "arcface"
    default_m = 0.50


def circle_loss(y_true, y_pred, gamma: int = 256, margin: float = 0.25):
    """losses.py:157-172 on device: y_true one-hot (B,n), y_pred raw cosines (B,n) -> (B,).
    Runs the head kernel in SAR_HEAD_CIRCLE_RAW mode with W = I (no re-normalisation of
    y_pred; multiplying by the identity is exact in fp32)."""
    return circle_from_cos(y_true, y_pred, gamma, margin)


def circle_from_cos(y_true, cos, gamma, margin):
    n = cos.shape[1]
    eye = torch.eye(n, device=cos.device, dtype=torch.float32)
    r = ops.head(None, None, emb_d=cos.contiguous(), wd=eye, onehot=y_true.contiguous().float(), n_classes=n,
                 head_kind="circle_raw", margin=margin, gamma=float(gamma))
    return r["sample_stats"][:, 1].contiguous()
