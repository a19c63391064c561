"""Training mode, sixth slice: the ResNet front-end (resnet.py:28-201) forward in TRAINING mode and its backward.

What changes against the inference engine: every BatchNormalization uses the batch statistics of its input (per replica,
as `multi_gpu_model` does) and updates its moving averages, so the fused `BN -> ReLU` epilogues of the tcgen05 kernels
(which fold the MOVING statistics into an affine) do not apply -- the convolutions run raw, a statistics pass follows, and
every intermediate the backward needs is kept.  This slice is the correctness-first form of that: fp32 CUDA-core kernels
behind the C ABI (`sar_conv2d_fwd`, `sar_bn_train_fwd/bwd`, `sar_conv2d_bwd_data`, `sar_conv2d_bwd_weight`,
`sar_maxpool2d_fwd/bwd`, `sar_relu_bwd`, `sar_axpy_fwd`), checked against float64 autograd (oracle/train_oracle.resnet_train),
not yet the tensor-core backward DESIGN.md §7 lists as the next step.

Structure (pre-activation basic blocks, resnet.py:105-125): x -> [BN1 -> ReLU ->] conv1 -> BN2 -> ReLU -> conv2 -> + shortcut(x),
shortcut = identity or a 1x1 'valid' convolution of the RAW block input (resnet.py:67-89); stem = conv 7x7/s2 -> BN -> ReLU ->
max-pool 3x3/s2; a final BN -> ReLU.  Conv kernels carry l2(1e-4) regularisers (resnet.py:36,56,86,116), biases and BNs none.
"""
from __future__ import annotations

from typing import Dict, List

import torch

from . import _shim, ops
from ._shim import check, ptr, stream_ptr
from . import training as T


def conv_bwd_data(spec, dy, w, B, dx=None, beta=0.0):
    if (spec.stride == 1 and beta == 0.0 and dx is None and spec.kh == spec.kw and spec.kh % 2 == 1
            and spec.pad_t == spec.kh // 2 and spec.pad_l == spec.kw // 2 and spec.hin == spec.hout and spec.win == spec.wout):
        # stride-1 SAME: d loss / d x is the forward convolution of dy with the kernel rotated by 180 degrees and its channel axes
        # swapped -- the implicit-GEMM forward kernel (sar_conv2d_fwd) is ~15x faster than the direct gradient kernel
        wf = w.flip(0, 1).permute(0, 1, 3, 2).contiguous()
        return ops.conv2d(dy, wf, None, stride=1, pad_t=spec.pad_t, pad_l=spec.pad_l, out_hw=(spec.hin, spec.win))
    if dx is None:
        dx = torch.empty((B, spec.hin, spec.win, spec.cin), device=dy.device, dtype=torch.float32)
    check(_shim.lib().sar_conv2d_bwd_data(ptr(dy), ptr(w), ptr(dx), B, spec.hin, spec.win, spec.cin, spec.hout, spec.wout, spec.cout,
                                          spec.kh, spec.kw, spec.stride, spec.pad_t, spec.pad_l, float(beta), stream_ptr()),
          "sar_conv2d_bwd_data")
    ops._count(1)
    return dx


def conv_bwd_weight(spec, x, dy, B):
    npos = B * spec.hout * spec.wout
    chunks = max(1, min(128, npos // 512))
    nw = spec.kh * spec.kw * spec.cin * spec.cout
    part = torch.empty((chunks, nw), device=dy.device, dtype=torch.float32)
    check(_shim.lib().sar_conv2d_bwd_weight(ptr(x), ptr(dy), ptr(part), chunks, B, spec.hin, spec.win, spec.cin, spec.hout, spec.wout,
                                            spec.cout, spec.kh, spec.kw, spec.stride, spec.pad_t, spec.pad_l, stream_ptr()),
          "sar_conv2d_bwd_weight")
    ops._count(1)
    return T.colsum(part).view(spec.kh, spec.kw, spec.cin, spec.cout)


def maxpool_bwd(x, dy, k, stride, pad_t, pad_l):
    B, H, W, C = x.shape
    _, Ho, Wo, _ = dy.shape
    dx = torch.empty_like(x)
    check(_shim.lib().sar_maxpool2d_bwd(ptr(x), ptr(dy), ptr(dx), B, H, W, C, Ho, Wo, k, stride, pad_t, pad_l, stream_ptr()),
          "sar_maxpool2d_bwd")
    ops._count(1)
    return dx


def axpy(x, y, alpha=1.0):
    check(_shim.lib().sar_axpy_fwd(ptr(x), ptr(y), float(alpha), x.numel(), stream_ptr()), "sar_axpy_fwd")
    ops._count(1)
    return y


class ResNetTrainer:
    """Training-mode forward + backward of the ResNet over the parameter dict `p` (device tensors, shared with HeadTrainer)."""

    def __init__(self, cfg, p: Dict[str, torch.Tensor]):
        self.cfg, self.p, self.plan = cfg, p, cfg.plan()
        self.tape = None

    @staticmethod
    def param_keys(cfg):
        """(trainable keys, l2-regularised keys, BN moving-statistic keys) in creation order."""
        plan = cfg.plan()
        keys: List[str] = ["resnet/stem/kernel", "resnet/stem/bias", "resnet/stem_bn/gamma", "resnet/stem_bn/beta"]
        l2, bns = ["resnet/stem/kernel"], ["resnet/stem_bn"]
        for b in plan.blocks:
            for c in (b.conv1, b.conv2, b.short):
                if c is None:
                    continue
                if c.pre_bn:
                    keys += [c.pre_bn + "/gamma", c.pre_bn + "/beta"]
                    bns.append(c.pre_bn)
                keys += [c.name + "/kernel", c.name + "/bias"]
                l2.append(c.name + "/kernel")
        keys += [plan.final_bn + "/gamma", plan.final_bn + "/beta"]
        bns.append(plan.final_bn)
        stats = [b + s for b in bns for s in ("/moving_mean", "/moving_variance")]
        return keys, l2, stats

    # ---- pieces
    def _conv(self, c, x, residual=None):
        p = self.p
        return ops.conv2d(x, p[c.name + "/kernel"], p[c.name + "/bias"], stride=c.stride, pad_t=c.pad_t, pad_l=c.pad_l,
                          out_hw=(c.hout, c.wout), residual=residual)

    def _bn_relu(self, x, name):
        p = self.p
        C = x.shape[-1]
        rows = x.reshape(-1, C)
        y, mean, inv = T.bn_train_fwd(rows, p[name + "/gamma"], p[name + "/beta"], p[name + "/moving_mean"], p[name + "/moving_variance"])
        a = T.bias_act(y, None, relu=True)
        return a.view(x.shape), (name, rows, a, mean, inv)

    def _bn_relu_bwd(self, g_a, saved, g: Dict[str, torch.Tensor]):
        name, rows, a, mean, inv = saved
        g_y = T.relu_bwd(g_a.reshape(rows.shape), a)
        g_x, g[name + "/gamma"], g[name + "/beta"] = T.bn_train_bwd(rows, g_y, self.p[name + "/gamma"], mean, inv)
        return g_x

    def _conv_bwd(self, c, x_in, g_out, g: Dict[str, torch.Tensor], B, want_dx=True, dx=None, beta=0.0):
        g[c.name + "/kernel"] = conv_bwd_weight(c, x_in, g_out, B)
        g[c.name + "/bias"] = T.colsum(g_out.reshape(-1, c.cout))
        return conv_bwd_data(c, g_out, self.p[c.name + "/kernel"], B, dx=dx, beta=beta) if want_dx else None

    # ---- forward (training mode), everything the backward needs kept on the tape
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        plan = self.plan
        if x.dim() == 3:
            x = x.unsqueeze(-1)
        x = x.contiguous()
        B = x.shape[0]
        tape = {"B": B, "x": x, "blocks": []}
        c0 = self._conv(plan.stem, x)
        a0, tape["stem_bn"] = self._bn_relu(c0, plan.stem.post_bn)
        h = ops.maxpool2d(a0, k=3, stride=2, pad_t=plan.pool_pad_t, pad_l=plan.pool_pad_l, out_hw=(plan.pool_hout, plan.pool_wout))
        tape["pool_in"] = a0
        for b in plan.blocks:
            rec = {"x": h}
            a1 = h
            if b.conv1.pre_bn:
                a1, rec["bn1"] = self._bn_relu(h, b.conv1.pre_bn)
            rec["a1"] = a1
            c1 = self._conv(b.conv1, a1)
            a2, rec["bn2"] = self._bn_relu(c1, b.conv2.pre_bn)
            rec["a2"] = a2
            short = self._conv(b.short, h) if b.short is not None else h
            h = self._conv(b.conv2, a2, residual=short)                      # Add(), resnet.py:89
            tape["blocks"].append(rec)
        out, tape["final_bn"] = self._bn_relu(h, plan.final_bn)
        self.tape = tape
        return out

    # ---- backward: g_out = d loss / d (final BN -> ReLU output) (B,Ho,Wo,C) -> gradients of every ResNet parameter
    def backward(self, g_out: torch.Tensor) -> Dict[str, torch.Tensor]:
        plan, tape = self.plan, self.tape
        B = tape["B"]
        g: Dict[str, torch.Tensor] = {}
        last = plan.blocks[-1].conv2
        gh = self._bn_relu_bwd(g_out.contiguous(), tape["final_bn"], g).view(B, last.hout, last.wout, last.cout)
        for b, rec in zip(reversed(plan.blocks), reversed(tape["blocks"])):
            # gh = d loss / d (block output) feeds both the residual branch and the shortcut
            g_a2 = self._conv_bwd(b.conv2, rec["a2"], gh, g, B)
            g_c1 = self._bn_relu_bwd(g_a2, rec["bn2"], g).view(B, b.conv1.hout, b.conv1.wout, b.conv1.cout)
            g_a1 = self._conv_bwd(b.conv1, rec["a1"], g_c1, g, B)
            gx = self._bn_relu_bwd(g_a1, rec["bn1"], g).view(rec["x"].shape) if "bn1" in rec else g_a1
            if b.short is not None:
                self._conv_bwd(b.short, rec["x"], gh, g, B, dx=gx, beta=1.0)
            else:
                axpy(gh, gx)
            gh = gx
        g_a0 = maxpool_bwd(tape["pool_in"], gh, 3, 2, plan.pool_pad_t, plan.pool_pad_l)
        g_c0 = self._bn_relu_bwd(g_a0, tape["stem_bn"], g).view(B, plan.stem.hout, plan.stem.wout, plan.stem.cout)
        self._conv_bwd(plan.stem, tape["x"], g_c0, g, B, want_dx=False)
        self.tape = None
        return g
