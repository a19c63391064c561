"""Mirror of the reference's VLAD.py: `VladPooling(mode, k_centers, g_centers=0)`.

Call surface (VLAD.py:10,26-28,48): layer([feat (B,1,S,D), cluster_score (B,1,S,K+G)]) ->
(B, K*D); weight `centers` (K+G, D) (VLAD.py:16-19); ghost clusters are the last G rows and
are dropped in 'gvlad' mode (VLAD.py:44-45); per-cluster L2 normalisation, no final
whole-vector normalisation (VLAD.py:47-48).

Inputs are CUDA float32 tensors; the work is one launch of the single-pass kernel
(csrc/vlad.cu) fed with the externally computed scores.  Inside SAR_Net the 1x1 assignment
conv is fused into the same launch instead (model.vlad / engine).
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import ops


class VladPooling:
    def __init__(self, mode, k_centers, g_centers=0, name=None, centers: Optional[np.ndarray] = None, **kwargs):
        if mode not in ("vlad", "gvlad"):
            raise ValueError("mode must be 'vlad' or 'gvlad'")
        self.mode = mode
        self.k_centers = int(k_centers)
        self.g_centers = int(g_centers)
        self.name = name
        self.cluster = None if centers is None else np.ascontiguousarray(centers, dtype=np.float32)
        self._dev = None
        self.built = centers is not None

    def build(self, input_shape):
        """VLAD.py:16-20: orthogonal-initialised `centers` of shape (K+G, D) (seeded here)."""
        if self.cluster is None:
            D = int(input_shape[0][-1])
            kg = self.k_centers + self.g_centers
            rng = np.random.RandomState(1234)
            q, _ = np.linalg.qr(rng.randn(max(kg, D), min(kg, D)))
            q = q if kg >= D else q.T
            self.cluster = np.ascontiguousarray(q[:kg, :D], dtype=np.float32)
        self.built = True

    def compute_output_shape(self, input_shape):
        assert input_shape
        return (input_shape[0][0], self.k_centers * input_shape[0][-1])

    def get_weights(self):
        return [self.cluster]

    def set_weights(self, ws):
        self.cluster = np.ascontiguousarray(ws[0], dtype=np.float32)
        self._dev = None
        self.built = True

    def __call__(self, x):
        return self.call(x)

    def call(self, x):
        feat, cluster_score = x
        if not self.built:
            self.build([tuple(feat.shape), tuple(cluster_score.shape)])
        B, D = feat.shape[0], feat.shape[-1]
        KG = self.k_centers + self.g_centers
        if cluster_score.shape[-1] != KG or self.cluster.shape != (KG, D):
            raise ValueError("cluster_score/centers do not match k_centers+g_centers=%d, D=%d" % (KG, D))
        if self._dev is None or self._dev.device != feat.device:
            self._dev = torch.from_numpy(self.cluster).to(feat.device)
        f = feat.reshape(B, -1, D).contiguous()
        sc = cluster_score.reshape(B, -1, KG).contiguous()
        # 'vlad' mode keeps every cluster (VLAD.py:44 only slices for 'gvlad'); with g_centers=0
        # both modes coincide.
        G = self.g_centers if self.mode == "gvlad" else 0
        if self.mode == "vlad" and self.g_centers:
            raise ValueError("mode='vlad' with g_centers>0 makes VLAD.py:48's reshape fail in the reference too")
        return ops.vlad(f, None, None, self._dev, KG - G, G, score=sc)
