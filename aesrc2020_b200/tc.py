"""Host side of the tensor-core residual-block path (csrc/conv_tc.cu, csrc/planes.cu):
flat-pad hi/lo plane buffers, tap tables for TF-SAME stride-1/stride-2 3x3 convolutions,
fp16 hi/lo weight packing, and the ctypes mirror of `sar_tc_conv` (include/sarnet.h).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

from . import _shim, ops
from ._shim import check, ptr, stream_ptr

LO_SCALE = 2048.0


class sar_tc_conv(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("a_rows", C.c_longlong), ("a_ch", C.c_int), ("a_planes", C.c_int),
        ("ntaps", C.c_int), ("tap_row_off", C.c_int * 9), ("tap_plane", C.c_int * 9),
        ("s", C.c_void_p), ("s_rows", C.c_longlong), ("s_ch", C.c_int), ("s_planes", C.c_int), ("s_plane", C.c_int),
        ("w", C.c_void_p), ("cout", C.c_int),
        ("bias", C.c_void_p),
        ("res", C.c_void_p),
        ("out_raw", C.c_void_p), ("out_act", C.c_void_p),
        ("act_scale", C.c_void_p), ("act_shift", C.c_void_p),
        ("out_dense", C.c_void_p),
        ("out_split", C.c_int),
        ("B", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("dbg", C.c_void_p),
        ("act_kind", C.c_int),
        ("nopad", C.c_int), ("ksplit", C.c_int),
        ("res_f32", C.c_void_p), ("out_raw_f32", C.c_void_p),
    ]

ACT_KIND = {"bn_relu": 0, None: 1, "none": 1, "linear": 1, "tanh": 2}


@dataclass
class Planes:
    """A flat-pad hi/lo planes buffer (see include/sarnet.h) and its geometry."""
    t: torch.Tensor          # (nplanes, rows, C) float16
    B: int
    H: int
    W: int
    C: int
    split: bool

    @property
    def rows(self) -> int:
        return int(self.t.shape[1])

    @property
    def nplanes(self) -> int:
        return int(self.t.shape[0])


def plane_rows(B: int, H: int, W: int, split: bool) -> int:
    if split:
        return B * ((H + 1) // 2 + 1) * ((W + 1) // 2 + 1)
    return B * (H + 1) * (W + 1)


def alloc_planes(B, H, W, Cc, split, device) -> Planes:
    """Zero-initialised (pads must stay zero; kernels never write them)."""
    n = 8 if split else 2
    t = torch.zeros((n, plane_rows(B, H, W, split), Cc), device=device, dtype=torch.float16)
    return Planes(t, B, H, W, Cc, bool(split))


def alloc_raw32(B, H, W, Cc, device) -> torch.Tensor:
    """The residual stream of a stage as ONE fp32 plane: (B*(H+1)*(W+1), C) flat-pad rows (pad rows are never read
    for anything that is kept)."""
    return torch.zeros((plane_rows(B, H, W, False), Cc), device=device, dtype=torch.float32)


def pack(x: torch.Tensor, split=False, affine=None, relu=False, out: Optional[Planes] = None) -> Planes:
    """dense fp32 NHWC (B,H,W,C) -> planes (sar_planes_pack_fwd)."""
    x = x.contiguous()
    B, H, W, Cc = x.shape
    if out is None:
        out = alloc_planes(B, H, W, Cc, split, x.device)
    s, t = affine if affine is not None else (None, None)
    check(_shim.lib().sar_planes_pack_fwd(ptr(x), ptr(s), ptr(t), 1 if relu else 0, ptr(out.t), B, H, W, Cc,
                                          1 if out.split else 0, stream_ptr()), "sar_planes_pack_fwd")
    ops._count(1)
    return out


def unpack(p: Planes) -> torch.Tensor:
    x = torch.empty((p.B, p.H, p.W, p.C), device=p.t.device, dtype=torch.float32)
    check(_shim.lib().sar_planes_unpack_fwd(ptr(p.t), ptr(x), p.B, p.H, p.W, p.C, 1 if p.split else 0, stream_ptr()),
          "sar_planes_unpack_fwd")
    ops._count(1)
    return x


def maxpool_planes(x: torch.Tensor, out: Planes, k, stride, pad_t, pad_l) -> Planes:
    x = x.contiguous()
    B, H, W, Cc = x.shape
    check(_shim.lib().sar_maxpool_planes_fwd(ptr(x), ptr(out.t), B, H, W, Cc, out.H, out.W, k, stride, pad_t, pad_l,
                                             stream_ptr()), "sar_maxpool_planes_fwd")
    ops._count(1)
    return out


def stem_pool(x: torch.Tensor, w, bias, scale, shift, out: Planes) -> Planes:
    """sar_stem_pool_fwd: x (B,T,D[,1]) fp32 -> pooled stem map in planes `out`."""
    x = x.contiguous()
    B, T, D = int(x.shape[0]), int(x.shape[1]), int(x.shape[2])
    check(_shim.lib().sar_stem_pool_fwd(ptr(x), ptr(w), ptr(bias), ptr(scale), ptr(shift), ptr(out.t), B, T, D, out.C,
                                        stream_ptr()), "sar_stem_pool_fwd")
    ops._count(1)
    return out


def pack_weights(kernel_hwio: np.ndarray, short_kernel: Optional[np.ndarray] = None) -> np.ndarray:
    """Keras HWIO kernel (+ optional 1x1 projection kernel) -> [2][Cout][Ktot] fp16 hi/lo,
    K-major with k = tap*Cin + ci and the shortcut's Cin_s rows appended."""
    kh, kw, cin, cout = kernel_hwio.shape
    wk = np.asarray(kernel_hwio, dtype=np.float32).reshape(kh * kw * cin, cout)
    if short_kernel is not None:
        wk = np.concatenate([wk, np.asarray(short_kernel, dtype=np.float32).reshape(-1, cout)], axis=0)
    wt = np.ascontiguousarray(wk.T)                                  # (Cout, Ktot)
    hi = wt.astype(np.float16)
    lo = ((wt - hi.astype(np.float32)) * LO_SCALE).astype(np.float16)
    return np.stack([hi, lo])


def tap_table(kh: int, kw: int, stride: int, pad_t: int, pad_l: int, W_out: int) -> Tuple[list, list]:
    """(row offsets, plane bases) per tap for a TF-SAME convolution on flat-pad planes.
    stride 1: input is a plain planes tensor with the output geometry, tap (r,s) reads row
              q + (r-pad_t)*(W+1) + (s-pad_l) of plane pair 0.
    stride 2: input is phase-split; input row 2*ho + r - pad_t has parity a&1 and half-index
              ho + (a>>1) with a = r - pad_t (same along w): plane pair (a_h&1)*2 + (a_w&1), row
              offset (a_h>>1)*(W_out+1) + (a_w>>1)."""
    P = W_out + 1
    offs, planes = [], []
    for r in range(kh):
        for s in range(kw):
            ah, aw = r - pad_t, s - pad_l
            if stride == 1:
                offs.append(ah * P + aw)
                planes.append(0)
            else:
                assert stride == 2
                offs.append((ah >> 1) * P + (aw >> 1))
                planes.append(2 * ((ah & 1) * 2 + (aw & 1)))
    return offs, planes


def conv_desc(a: Planes, w_packed: torch.Tensor, bias: torch.Tensor, *, out_hw: Tuple[int, int], taps, cout: int,
            short: Optional[Planes] = None, res: Optional[Planes] = None, out_raw: Optional[Planes] = None,
            out_act: Optional[Planes] = None, act=None, out_dense: Optional[torch.Tensor] = None, dbg=None,
            act_kind: int = 0, nopad: bool = False, ksplit: int = 1, res_f32: Optional[torch.Tensor] = None,
            out_raw_f32: Optional[torch.Tensor] = None) -> sar_tc_conv:
    """The `sar_tc_conv` descriptor of one layer.  taps = (row_offsets, plane_bases).  res_f32 / out_raw_f32: the
    residual stream as one fp32 (R, cout) plane (alloc_raw32) instead of hi/lo planes."""
    d = sar_tc_conv()
    d.a, d.a_rows, d.a_ch, d.a_planes = ptr(a.t), a.rows, a.C, a.nplanes
    offs, planes = taps
    d.ntaps = len(offs)
    for i, (o, pl) in enumerate(zip(offs, planes)):
        d.tap_row_off[i] = int(o)
        d.tap_plane[i] = int(pl)
    if short is not None:
        d.s, d.s_rows, d.s_ch, d.s_planes, d.s_plane = ptr(short.t), short.rows, short.C, short.nplanes, 0
    d.w, d.cout = ptr(w_packed), cout
    d.bias = ptr(bias)
    H, W = out_hw
    if res is not None:
        assert not res.split and (res.H, res.W, res.C) == (H, W, cout)
        d.res = ptr(res.t)
    split = None
    for o in (out_raw, out_act):
        if o is not None:
            assert (o.H, o.W, o.C) == (H, W, cout), ((o.H, o.W, o.C), (H, W, cout))
            assert split is None or split == o.split
            split = o.split
    d.out_raw = ptr(out_raw.t) if out_raw is not None else None
    d.out_act = ptr(out_act.t) if out_act is not None else None
    if act is not None:
        d.act_scale, d.act_shift = ptr(act[0]), ptr(act[1])
    d.out_dense = ptr(out_dense) if out_dense is not None else None
    d.out_split = 1 if split else 0
    d.B, d.H, d.W = a.B, H, W
    d.dbg = ptr(dbg) if dbg is not None else None
    d.act_kind = int(act_kind)
    d.nopad, d.ksplit = (1 if nopad else 0), int(ksplit)
    R = plane_rows(a.B, H, W, False)
    for t_ in (res_f32, out_raw_f32):
        if t_ is not None:
            assert t_.dtype == torch.float32 and tuple(t_.shape) == (R, cout) and t_.is_contiguous(), (tuple(t_.shape), (R, cout))
    d.res_f32 = ptr(res_f32) if res_f32 is not None else None
    d.out_raw_f32 = ptr(out_raw_f32) if out_raw_f32 is not None else None
    return d


def conv_launch(d: sar_tc_conv):
    check(_shim.lib().sar_conv_tc_fwd(C.byref(d), stream_ptr()), "sar_conv_tc_fwd")
    ops._count(1)


def conv_tc(*args, **kwargs):
    """sar_conv_tc_fwd (arguments of conv_desc)."""
    conv_launch(conv_desc(*args, **kwargs))


def chain_workspace(descs, device) -> torch.Tensor:
    arr = (sar_tc_conv * len(descs))(*descs)
    n = _shim.lib().sar_conv_tc_chain_workspace_bytes(arr, len(descs))
    return torch.zeros((n + 3) // 4, device=device, dtype=torch.int32)


def conv_tc_chain(descs, workspace: torch.Tensor, max_ctas: int = 0):
    """sar_conv_tc_chain_grid_fwd: the stride-1 3x3 layers of one stage in one persistent (cooperative) launch;
    `max_ctas` caps the grid so that chains of several streams share the SMs."""
    arr = (sar_tc_conv * len(descs))(*descs)
    check(_shim.lib().sar_conv_tc_chain_grid_fwd(arr, len(descs), ptr(workspace), workspace.numel() * 4, int(max_ctas),
                                                 stream_ptr()), "sar_conv_tc_chain_fwd")
    ops._count(1)


def pack_dense_weights(kernel: np.ndarray) -> np.ndarray:
    """Keras Dense kernel (Din, Dout) -> the 1-tap packed form [2][Dout][Din] fp16 hi/lo."""
    din, dout = kernel.shape
    return pack_weights(np.asarray(kernel).reshape(1, 1, din, dout))


def dense_tc(a: Planes, w_packed: torch.Tensor, bias: torch.Tensor, *, act=None, nopad: bool = False) -> torch.Tensor:
    """Dense on the channel axis of a planes tensor as a 1-tap tensor-core "convolution"
    (Dense of model.py:35-42 / the GRU input projections of model.py:44-50).
    Returns fp32 (a.B, a.H, a.W, Dout) in the plain layout (pad rows are skipped)."""
    dout = int(w_packed.shape[1])
    out = torch.empty((a.B, a.H, a.W, dout), device=a.t.device, dtype=torch.float32)
    conv_tc(a, w_packed, bias, out_hw=(a.H, a.W), taps=([0], [0]), cout=dout, out_dense=out, act_kind=ACT_KIND[act],
            nopad=nopad)
    return out


def alloc_rows(M: int, Cc: int, device) -> Planes:
    """hi/lo planes [2][M][C] of a plain row-major (M, C) matrix (no pad rows): the A operand of gemm_splitk_tc."""
    t = torch.zeros((2, M, Cc), device=device, dtype=torch.float16)
    return Planes(t, 1, M, 1, Cc, False)


def gemm_splitk_tc(a: Planes, w_packed: torch.Tensor, bias: torch.Tensor, zero_bias: torch.Tensor, ksplit: int,
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out (M, N) = a (M, K) @ w (K, N) + bias on the tensor cores with the long K axis cut into `ksplit` slices
    (AR_EMBEDDING, model.py:286-289: M = batch, K = clusters * hidden = 16384): slice z is one tile of the 1-tap
    conv_tc kernel writing fp32 partial products, sar_splitk_reduce_fwd sums them in a fixed order and adds the bias."""
    M, N = a.H, int(w_packed.shape[1])
    ws = torch.empty((ksplit, M, N), device=a.t.device, dtype=torch.float32)
    conv_tc(a, w_packed, zero_bias, out_hw=(a.H, a.W), taps=([0], [0]), cout=N, out_dense=ws, act_kind=ACT_KIND[None],
            nopad=True, ksplit=ksplit)
    if out is None:
        out = torch.empty((M, N), device=a.t.device, dtype=torch.float32)
    assert tuple(out.shape) == (M, N) and out.is_contiguous()
    check(_shim.lib().sar_splitk_reduce_fwd(ptr(ws), ptr(bias), ptr(out), M, N, ksplit, stream_ptr()), "sar_splitk_reduce_fwd")
    ops._count(1)
    return out


def pack_vlad_assign(w_assign: np.ndarray) -> np.ndarray:
    """Assignment kernel (D, K+G) -> the packed form sar_vlad_tc_fwd takes: [2][KGP][D] fp16 hi/lo, KGP = K+G rounded up
    to 16 with zero rows."""
    d, kg = w_assign.shape
    pad = (-kg) % 16
    return pack_dense_weights(np.pad(np.asarray(w_assign, dtype=np.float32), ((0, 0), (0, pad))))


def vlad_tc_supported(B: int, S: int, D: int, K: int, G: int) -> bool:
    return bool(_shim.load_library().sar_vlad_tc_supported(int(B), int(S), int(D), int(K), int(G)))


def vlad_tc(x: Planes, wa_packed: torch.Tensor, b_assign: torch.Tensor, centers: torch.Tensor, B: int, S: int, K: int,
            G: int, *, planes: Optional[Planes] = None, want_dense: bool = True) -> Optional[torch.Tensor]:
    """sar_vlad_tc_fwd: x = hi/lo planes of the (B*S, 256) descriptors -> (B, K*256) fp32 and/or `planes` (alloc_rows(B, K*256))."""
    D = x.C
    assert x.rows >= B * S and tuple(centers.shape) == (K + G, D)
    out = torch.empty((B, K * D), device=x.t.device, dtype=torch.float32) if (want_dense or planes is None) else None
    if planes is not None:
        assert tuple(planes.t.shape) == (2, B, K * D), (tuple(planes.t.shape), (2, B, K * D))
    check(_shim.lib().sar_vlad_tc_fwd(ptr(x.t), x.rows, ptr(wa_packed), ptr(b_assign), ptr(centers), ptr(out),
                                      ptr(planes.t) if planes is not None else None, B, S, D, K, G, stream_ptr()),
          "sar_vlad_tc_fwd")
    ops._count(1)
    return out
