"""Device-side execution plan of the SAR-Net forward (model.predict, model.py:229-339).

`SARNetEngine` owns the device copies of the weights in kernel-native form (inference BN
folded to per-channel affines, GRU kernels of both directions concatenated, AR_BN1/AR_BN2
folded into AR_EMBEDDING) and launches the C-ABI kernels in order on the current stream.
No PyTorch math is on the path: torch only allocates buffers.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import torch

from . import ops
from .config import SARConfig, ResNetPlan

BN_EPS = 1e-3      # Keras BatchNormalization default epsilon (resnet.py:25, model.py:29-30)


@dataclass(frozen=True)
class StepOpts:
    """How ONE step is launched.  Passed explicitly down the call chain (never stored on the engine or in a module
    global), so that several steps -- pipeline slots, micro-batch lanes, threads -- can be captured concurrently.

    lane      buffer namespace: steps that may be in flight at the same time never share activation buffers
    no_chain  per-layer conv launches only (no persistent stage-chain launch: a chain needs its whole grid resident)
    gru_nb    utterances per Bi-GRU cluster (0 = the kernel's choice, 32 = fewer SMs per step)
    """
    lane: int = 0
    no_chain: bool = False
    gru_nb: int = 0
    chain_ctas: int = 0       # > 0: stage chains with at most this many CTAs (several steps in flight share the SMs)


DEFAULT_OPTS = StepOpts()
import os as _os
_SLOT_CHAIN = int(_os.environ.get("SAR_SLOT_CHAIN_CTAS", "0"))       # experiment: capped stage chains inside pipeline slots


def SLOT_OPTS(slot):
    if _SLOT_CHAIN > 0:
        return StepOpts(lane=slot, no_chain=False, gru_nb=32, chain_ctas=_SLOT_CHAIN)
    return StepOpts(lane=slot, no_chain=True, gru_nb=32)


def fold_bn(w: Dict[str, np.ndarray], name: str):
    """Inference BN -> (scale, shift) in float64."""
    g = w[name + "/gamma"].astype(np.float64)
    b = w[name + "/beta"].astype(np.float64)
    m = w[name + "/moving_mean"].astype(np.float64)
    v = w[name + "/moving_variance"].astype(np.float64)
    scale = g / np.sqrt(v + BN_EPS)
    return scale, b - m * scale


class ResNetDevice:
    """ResNet-18/34 front-end (resnet.py:170-201) on device buffers."""

    def __init__(self, plan: ResNetPlan, weights: Dict[str, np.ndarray], device):
        self.plan = plan
        self.device = device
        self.p: Dict[str, torch.Tensor] = {}

        def put(name, arr):
            self.p[name] = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32)).to(device)

        for c in plan.convs():
            put(c.name + "/kernel", weights[c.name + "/kernel"])
            put(c.name + "/bias", weights[c.name + "/bias"])
        bns = [plan.stem.post_bn, plan.final_bn]
        for b in plan.blocks:
            bns += [x for x in (b.conv1.pre_bn, b.conv2.pre_bn) if x]
        for name in bns:
            s, t = fold_bn(weights, name)
            put(name + "/scale", s)
            put(name + "/shift", t)

    def bn(self, name):
        return (self.p[name + "/scale"], self.p[name + "/shift"]) if name else None

    def conv(self, x, c, residual=None, act=None):
        return ops.conv2d(x, self.p[c.name + "/kernel"], self.p[c.name + "/bias"], stride=c.stride,
                          pad_t=c.pad_t, pad_l=c.pad_l, out_hw=(c.hout, c.wout), pre=self.bn(c.pre_bn),
                          post=self.bn(c.post_bn), residual=residual, act=act)

    def forward_raw(self, x: torch.Tensor) -> torch.Tensor:
        """x (B,T,80,1) -> the residual stream BEFORE the final BN->ReLU (B,H',W',C)."""
        pl = self.plan
        a = self.conv(x, pl.stem, act="relu")                                   # resnet.py:173/191
        a = ops.maxpool2d(a, k=3, stride=2, pad_t=pl.pool_pad_t, pad_l=pl.pool_pad_l,
                          out_hw=(pl.pool_hout, pl.pool_wout))                  # resnet.py:174/192
        for b in pl.blocks:                                                     # resnet.py:105-125
            c1 = self.conv(a, b.conv1)
            sc = self.conv(a, b.short) if b.short else a                       # resnet.py:67-89
            a = self.conv(c1, b.conv2, residual=sc)
        return a

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        a = self.forward_raw(x)
        s, t = self.bn(self.plan.final_bn)
        return ops.affine_relu(a, s, t, relu=True)                             # resnet.py:178/196


class ResNetTC:
    """ResNet-18/34 front-end with the residual blocks on tcgen05 tensor cores.

    stem 7x7/s2 (+BN+ReLU) : CUDA-core direct conv (Cin=1, K=49 is not GEMM-shaped)
    max-pool 3x3/s2        : writes flat-pad hi/lo planes directly
    every block conv       : csrc/conv_tc.cu; conv2 carries the shortcut (identity add or the 1x1
                             projection as extra K-steps) and writes the NEXT block's inputs: the raw
                             sum (its shortcut operand) and relu(bn1_next(.)) (its conv1 operand),
                             phase-split when the next block is strided.  The last conv2 writes the
                             final BN->ReLU map as dense fp32 (B,H',W',C) for CNN_LIN.
    """

    def __init__(self, plan: ResNetPlan, weights: Dict[str, np.ndarray], device):
        from . import tc
        self.tc = tc
        self.plan = plan
        self.device = device
        self.p: Dict[str, torch.Tensor] = {}
        self._bufs: Dict[tuple, list] = {}
        self._rot: Dict[tuple, int] = {}
        self.record_events = None
        import os as _os
        # stages whose stride-1 layers run as one persistent chain launch (SAR_CHAIN_STAGES="" disables)
        self.chain_stages = {int(x) for x in _os.environ.get("SAR_CHAIN_STAGES", "2,3,4").split(",") if x.strip()}
        # ... but only while a layer has few tiles per SM: the chain removes per-layer launch / prologue / epilogue tails
        # (B=64: conv segment 0.527 -> 0.495 ms), at large batches those are amortised and the per-layer kernels are
        # faster (weights resident or CTA pairs; B=512: 3.17 -> 2.72 ms without chains).  Items = 128 x 64 tiles.
        self.chain_max_items = int(_os.environ.get("SAR_CHAIN_MAX_ITEMS", str(3 * 148)))
        # residual stream between identity-shortcut blocks as one fp32 plane (SAR_RAW32=0: hi/lo planes, an A/B aid)
        self.raw32 = _os.environ.get("SAR_RAW32", "1") != "0" and _os.environ.get("SAR_TC_TMA_OUT", "1") != "0"

        def put(name, arr, dtype=np.float32):
            self.p[name] = torch.from_numpy(np.ascontiguousarray(arr, dtype=dtype)).to(device)

        put("stem/kernel", weights[plan.stem.name + "/kernel"])
        put("stem/bias", weights[plan.stem.name + "/bias"])
        bns = [plan.stem.post_bn, plan.final_bn]
        for b in plan.blocks:
            bns += [x for x in (b.conv1.pre_bn, b.conv2.pre_bn) if x]
            put(b.name + "/w1", tc.pack_weights(weights[b.conv1.name + "/kernel"]), np.float16)
            put(b.name + "/b1", weights[b.conv1.name + "/bias"])
            if b.short:
                put(b.name + "/w2", tc.pack_weights(weights[b.conv2.name + "/kernel"], weights[b.short.name + "/kernel"]),
                    np.float16)
                put(b.name + "/b2", weights[b.conv2.name + "/bias"].astype(np.float64)
                    + weights[b.short.name + "/bias"].astype(np.float64))
            else:
                put(b.name + "/w2", tc.pack_weights(weights[b.conv2.name + "/kernel"]), np.float16)
                put(b.name + "/b2", weights[b.conv2.name + "/bias"])
        for name in bns:
            s, t = fold_bn(weights, name)
            put(name + "/scale", s)
            put(name + "/shift", t)

    def bn(self, name):
        return (self.p[name + "/scale"], self.p[name + "/shift"]) if name else None

    def _buf(self, lane, B, H, W, C, split, role):
        """2-slot rotation of zero-initialised plane buffers per (lane, geometry, role)."""
        key = (lane, B, H, W, C, bool(split), role)
        if key not in self._bufs:
            self._bufs[key] = [self.tc.alloc_planes(B, H, W, C, split, self.device) for _ in range(2)]
            self._rot[key] = 0
        self._rot[key] ^= 1
        return self._bufs[key][self._rot[key]]

    def forward(self, x: torch.Tensor, as_planes: bool = False, opts: StepOpts = DEFAULT_OPTS):
        """x (B,T,80,1) fp32 -> (B,H',W',C) after the final BN->ReLU: fp32 NHWC, or (as_planes) the hi/lo
        planes a tensor-core Dense consumes directly."""
        tc, pl, p = self.tc, self.plan, self.p
        B = x.shape[0]
        st = pl.stem
        lane = opts.lane
        cur = self._buf(lane, B, pl.pool_hout, pl.pool_wout, st.cout, False, "raw")   # the first block is never strided
        if st.cout % 16 == 0 and st.cout <= 64 and st.wout % 4 == 0:
            s_, t_ = self.bn(st.post_bn)                                        # fused stem: conv+BN+ReLU+pool
            tc.stem_pool(x, p["stem/kernel"], p["stem/bias"], s_, t_, cur)
        else:                                                                   # wide stems: conv, then pool
            a = ops.conv2d(x, p["stem/kernel"], p["stem/bias"], stride=st.stride, pad_t=st.pad_t, pad_l=st.pad_l,
                           out_hw=(st.hout, st.wout), post=self.bn(st.post_bn), act="relu")
            tc.maxpool_planes(a, cur, 3, 2, pl.pool_pad_t, pl.pool_pad_l)
        cur_act = cur_raw = cur
        out_dense = None
        ev = None
        if self.record_events is not None:       # bench.py: device time of the block convolutions only
            # external events become event-record NODES when the step is being captured into a CUDA graph, so
            # the segment can be timed inside a replay (ordinary events cannot be recorded during capture)
            ext = torch.cuda.is_current_stream_capturing()
            ev = (torch.cuda.Event(enable_timing=True, external=ext), torch.cuda.Event(enable_timing=True, external=ext))
            ev[0].record()
        # The stride-1 3x3 layers of a stage (conv2 of its first block, then conv1/conv2 of every further block)
        # go out as ONE persistent launch (tc.conv_tc_chain) when the stage is in self.chain_stages; everything
        # else (the stage's first conv1: strided, or 64 -> 32 channels in stage 1) is a launch of its own.
        pending: list = []
        stage = 0

        def flush():
            if not pending:
                return
            if len(pending) == 1:
                tc.conv_launch(pending[0])
            else:
                key = ("chain", lane, B, stage, len(pending))
                if key not in self._bufs:
                    self._bufs[key] = tc.chain_workspace(pending, self.device)
                tc.conv_tc_chain(pending, self._bufs[key], max_ctas=opts.chain_ctas)
            pending.clear()

        def emit(desc, chainable):
            # lanes run concurrently on separate streams: a chain launch (CTAs spinning on tile counters of CTAs of
            # the SAME launch) needs its whole grid resident, which two lanes sharing the SMs cannot promise
            items = -(-(desc.B * (desc.H + 1) * (desc.W + 1)) // 128) * max(1, desc.cout // 64)
            if chainable and stage in self.chain_stages and not opts.no_chain and items <= self.chain_max_items:
                pending.append(desc)
            else:
                flush()
                tc.conv_launch(desc)

        raw32_in = None          # the previous block's raw sum as one fp32 plane (None: cur_raw planes)
        for i, b in enumerate(pl.blocks):
            nxt = pl.blocks[i + 1] if i + 1 < len(pl.blocks) else None
            c1, c2 = b.conv1, b.conv2
            first = i == 0 or c1.stride == 2
            if first:
                flush()
                stage += 1
            c1_act = self._buf(lane, B, c1.hout, c1.wout, c1.cout, False, "c1")
            emit(tc.conv_desc(cur_act, p[b.name + "/w1"], p[b.name + "/b1"], out_hw=(c1.hout, c1.wout),
                              taps=tc.tap_table(3, 3, c1.stride, c1.pad_t, c1.pad_l, c1.wout), cout=c1.cout,
                              out_act=c1_act, act=self.bn(c2.pre_bn)), chainable=not first)
            taps2 = tc.tap_table(3, 3, 1, 1, 1, c2.wout)
            # identity shortcut: the previous block's raw sum, as hi/lo planes or (raw32_in) one fp32 plane
            common = dict(out_hw=(c2.hout, c2.wout), taps=taps2, cout=c2.cout, short=cur_raw if b.short else None,
                          res=None if (b.short or raw32_in is not None) else cur_raw,
                          res_f32=None if b.short else raw32_in)
            if nxt is not None:
                ns = nxt.conv1.stride == 2
                nact = self._buf(lane, B, c2.hout, c2.wout, c2.cout, ns, "act")
                # The raw sum feeds the NEXT block's shortcut.  If that is an identity shortcut it is only added in
                # conv2's epilogue: one fp32 plane, no hi/lo split.  If it is a projection (first block of the next
                # stage) it is an MMA operand: hi/lo planes.
                if self.raw32 and nxt.short is None and not ns:
                    key = (lane, B, c2.hout, c2.wout, c2.cout, "raw32")
                    if key not in self._bufs:
                        self._bufs[key] = [tc.alloc_raw32(B, c2.hout, c2.wout, c2.cout, self.device) for _ in range(2)]
                        self._rot[key] = 0
                    self._rot[key] ^= 1
                    nraw32 = self._bufs[key][self._rot[key]]
                    emit(tc.conv_desc(c1_act, p[b.name + "/w2"], p[b.name + "/b2"], out_raw_f32=nraw32, out_act=nact,
                                      act=self.bn(nxt.conv1.pre_bn), **common), chainable=True)
                    cur_raw, raw32_in = None, nraw32
                else:
                    nraw = self._buf(lane, B, c2.hout, c2.wout, c2.cout, ns, "raw")
                    emit(tc.conv_desc(c1_act, p[b.name + "/w2"], p[b.name + "/b2"], out_raw=nraw, out_act=nact,
                                      act=self.bn(nxt.conv1.pre_bn), **common), chainable=True)
                    cur_raw, raw32_in = nraw, None
                cur_act = nact
            elif as_planes:
                out_dense = self._buf(lane, B, c2.hout, c2.wout, c2.cout, False, "final")
                emit(tc.conv_desc(c1_act, p[b.name + "/w2"], p[b.name + "/b2"], act=self.bn(pl.final_bn),
                                  out_act=out_dense, **common), chainable=True)
            else:
                out_dense = torch.empty((B, c2.hout, c2.wout, c2.cout), device=self.device, dtype=torch.float32)
                emit(tc.conv_desc(c1_act, p[b.name + "/w2"], p[b.name + "/b2"], act=self.bn(pl.final_bn),
                                  out_dense=out_dense, **common), chainable=False)
        flush()
        if ev is not None:
            ev[1].record()
            self.record_events.append(ev)
        return out_dense


class SARNetEngine:
    def __init__(self, cfg: SARConfig, weights: Dict[str, np.ndarray], device="cuda", conv_path: str = "tc"):
        if cfg.ar_enable and cfg.mto not in ("avg", "bigru", "vlad", "gvlad"):
            raise ValueError("Please specify avg/bigru/vlad/gvlad ..")          # model.py:136-138
        self.cfg = cfg
        self.device = torch.device(device)
        self.plan = cfg.plan()
        self.conv_path = conv_path
        tc_ok = all(c.cin % 32 == 0 and c.cout % 32 == 0 for c in self.plan.convs()[1:])
        if conv_path == "tc" and not tc_ok:
            # The tcgen05 kernels tile channels in chunks of 32.  Other widths (e.g. res_filters=16) run the
            # hand-written CUDA-core kernels (csrc/conv_ffma.cu) -- still this library on the GPU, but a different,
            # much slower code path: say so instead of selecting it silently ("Supported shapes").
            import warnings
            warnings.warn("aesrc2020_b200: res_filters=%d gives channel counts that are not multiples of 32; the residual "
                          "blocks run on the CUDA-core conv_ffma kernels instead of the tcgen05 tensor-core path"
                          % cfg.res_filters, RuntimeWarning, stacklevel=3)
            conv_path = self.conv_path = "ffma"
        self.resnet = (ResNetTC if conv_path == "tc" else ResNetDevice)(self.plan, weights, self.device)
        self._graphs: Dict[tuple, tuple] = {}
        self.p: Dict[str, torch.Tensor] = {}
        # Dense layers / GRU input projections on the tensor cores (1-tap conv_tc) when the residual blocks are
        self.dense_tc = conv_path == "tc" and cfg.hidden_dim % 64 == 0 and self.plan.cout % 32 == 0
        self._seq_bufs: Dict[tuple, object] = {}
        self._lane_streams: List[torch.cuda.Stream] = []
        self._lane_done: List[torch.cuda.Event] = []
        self._lane_sinks: Dict[tuple, Dict[str, torch.Tensor]] = {}
        self._lane_in: Optional[torch.cuda.Event] = None
        self._branch_streams: Dict[int, torch.cuda.Stream] = {}
        # CTC branch and accent branch (independent after CRNN_LN, model.py:261-296) on parallel graph branches
        import os as _os
        self.parallel_branches = _os.environ.get("SAR_PARALLEL_BRANCHES", "1") != "0"
        self._prepare(weights)

    # ------------------------------------------------------------------ weight preparation
    def _put(self, name, arr):
        self.p[name] = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32)).to(self.device)

    def _dense(self, w, name, bias=True, tc_ok=True):
        self._put(name + "/kernel", w[name + "/kernel"])
        if bias:
            self._put(name + "/bias", w[name + "/bias"])
        k = w[name + "/kernel"]
        if self.dense_tc and tc_ok and name in ("CNN_LIN", "CTC_DS", "AR_DS") and k.shape[0] % 32 == 0 and k.shape[1] % 32 == 0:
            from . import tc
            self.p[name + "/w_tc"] = torch.from_numpy(tc.pack_dense_weights(k)).to(self.device)

    def _ln(self, w, name):
        self._put(name + "/gamma", w[name + "/gamma"])
        self._put(name + "/beta", w[name + "/beta"])

    def _bigru(self, w, name):
        u = w[name + "/forward/recurrent_kernel"].shape[0]
        kf, kb = w[name + "/forward/kernel"], w[name + "/backward/kernel"]
        bf, bb = w[name + "/forward/bias"], w[name + "/backward/bias"]
        self._put(name + "/kernel_cat", np.concatenate([kf, kb], axis=1))                 # (Din, 6u)
        if self.dense_tc and kf.shape[0] % 32 == 0 and (6 * u) % 32 == 0:
            from . import tc
            self.p[name + "/w_tc"] = torch.from_numpy(tc.pack_dense_weights(np.concatenate([kf, kb], axis=1))).to(self.device)
        self._put(name + "/ibias_cat", np.concatenate([bf[:3 * u], bb[:3 * u]]))          # (6u,)
        self._put(name + "/rec", np.stack([w[name + "/forward/recurrent_kernel"],
                                           w[name + "/backward/recurrent_kernel"]]))      # (2,u,3u)
        self._put(name + "/rbias", np.stack([bf[3 * u:], bb[3 * u:]]))                    # (2,3u)

    def _prepare(self, w):
        cfg = self.cfg
        self._dense(w, "CNN_LIN"); self._ln(w, "CNN_LIN_LN")
        self._bigru(w, "CRNN"); self._ln(w, "CRNN_LN")
        if cfg.ctc_enable:
            self._bigru(w, "CTC_BIGRU"); self._ln(w, "CTC_BIGRU_LN")
            self._dense(w, "CTC_DS"); self._ln(w, "CTC_DS_LN")
            self._dense(w, "ctc_pred", tc_ok=False)
            if self.dense_tc and w["ctc_pred/kernel"].shape[0] % 32 == 0:
                # ctc_pred (Dout = bpe_classes, 1000 by default) on the tensor cores: columns padded with zero
                # weights to a multiple of 32; sar_ctc_ld_fwd reads the first bpe_classes columns of each row
                from . import tc
                k, b = w["ctc_pred/kernel"], w["ctc_pred/bias"]
                padc = (-k.shape[1]) % 32
                self.p["ctc_pred/w_tc"] = torch.from_numpy(tc.pack_dense_weights(np.pad(k, ((0, 0), (0, padc))))).to(self.device)
                self._put("ctc_pred/bias_tc", np.pad(b, (0, padc)))
        if cfg.ar_enable:
            self._dense(w, "AR_DS"); self._ln(w, "AR_DS_LN")
            if cfg.mto == "bigru":
                self._bigru(w, "AR_MERGE")
            elif cfg.mto in ("vlad", "gvlad"):
                pre = cfg.mto
                k = w[pre + "_center_assignment/kernel"]
                self._put(pre + "/w_assign", k.reshape(k.shape[2], k.shape[3]))
                self._put(pre + "/b_assign", w[pre + "_center_assignment/bias"])
                self._put(pre + "/centers", w[pre + "_pool/centers"])
                if self.dense_tc:
                    from . import tc
                    self.p[pre + "/w_assign_tc"] = torch.from_numpy(tc.pack_vlad_assign(k.reshape(k.shape[2], k.shape[3]))).to(self.device)
            # AR_BN1 -> AR_EMBEDDING -> AR_BN2 folded into one affine map (float64 on the host)
            s1, t1 = fold_bn(w, "AR_BN1")
            s2, t2 = fold_bn(w, "AR_BN2")
            W = w["AR_EMBEDDING/kernel"].astype(np.float64)
            b = w["AR_EMBEDDING/bias"].astype(np.float64)
            self._put("AR_EMBEDDING/kernel_folded", (s1[:, None] * W) * s2[None, :])
            self._put("AR_EMBEDDING/bias_folded", (t1 @ W + b) * s2 + t2)
            Wf = (s1[:, None] * W) * s2[None, :]
            self.embed_ksplit = 0
            if self.dense_tc and cfg.mto in ("vlad", "gvlad") and Wf.shape[0] % 2048 == 0 and Wf.shape[1] % 32 == 0:
                from . import tc
                # tensor-core split-K GEMM: K slices of 256 channels (4 k-steps), e.g. 64 slices at K = 16384
                self.embed_ksplit = Wf.shape[0] // 256
                self.p["AR_EMBEDDING/w_tc"] = torch.from_numpy(tc.pack_dense_weights(Wf.astype(np.float32))).to(self.device)
                self.p["AR_EMBEDDING/zero_bias"] = torch.zeros(Wf.shape[1], device=self.device, dtype=torch.float32)
            for n in ("AR_CF_DS1", "AR_CF_DS2", "y_accent"):
                self._dense(w, n)
            if cfg.disc_enable:
                key = "y_disc/W" if cfg.metric_loss in ("sphereface", "cosface", "arcface") else "y_disc/kernel"
                self._put("y_disc/w", w[key])
            if cfg.disc_enable and cfg.bn_dim:
                self._dense(w, "AR_BN_DS")
                s3, t3 = fold_bn(w, "AR_BN3")
                s4, t4 = fold_bn(w, "AR_BN4")
                Wb = w["bottleneck/kernel"].astype(np.float64)
                bb = w["bottleneck/bias"].astype(np.float64)
                self._put("bottleneck/kernel_folded", (s3[:, None] * Wb) * s4[None, :])
                self._put("bottleneck/bias_folded", (t3 @ Wb + bb) * s4 + t4)
                key = "y_disc_bn/W" if cfg.metric_loss in ("sphereface", "cosface", "arcface") else "y_disc_bn/kernel"
                self._put("y_disc_bn/w", w[key])

    # ------------------------------------------------------------------ timed segments (bench.py)
    # segment_events = {"vlad": []}: the named segment is bracketed by a CUDA event pair; while the step is being
    # captured the events are EXTERNAL ones, i.e. event-record nodes of the graph, re-recorded by every replay.
    segment_events: Optional[Dict[str, list]] = None

    def _seg_begin(self, name):
        if self.segment_events is None or name not in self.segment_events:
            return None
        ext = torch.cuda.is_current_stream_capturing()
        e0 = torch.cuda.Event(enable_timing=True, external=ext)
        e1 = torch.cuda.Event(enable_timing=True, external=ext)
        e0.record()
        return name, e0, e1

    def _seg_end(self, h):
        if h is not None:
            h[2].record()
            self.segment_events[h[0]].append((h[1], h[2]))

    # ------------------------------------------------------------------ building blocks
    def dense_ln(self, x, name, ln_name, act="tanh", pre=None):
        p = self.p
        y = ops.dense(x, p[name + "/kernel"], p[name + "/bias"], act=act, pre=pre)
        return ops.layernorm(y, p[ln_name + "/gamma"], p[ln_name + "/beta"])

    def bigru(self, x, name, seq=True, opts: StepOpts = DEFAULT_OPTS):
        p = self.p
        B, S, _ = x.shape
        xp = ops.dense(x, p[name + "/kernel_cat"], p[name + "/ibias_cat"])       # (B,S,6u) = (B,S,2,3u)
        return ops.bigru(xp.reshape(B, S, 2, -1), p[name + "/rec"], p[name + "/rbias"], seq=seq, nb=opts.gru_nb)

    # tensor-core variants: sequence activations travel as hi/lo planes of a (1, B, S) map
    def _seq_planes(self, lane, B, S, C, role):
        from . import tc
        key = (lane, B, S, C, role)
        if key not in self._seq_bufs:
            self._seq_bufs[key] = tc.alloc_rows(B * S, C, self.device)       # plain rows: a Dense has one tap, no halo
        return self._seq_bufs[key]

    def ln_planes(self, x, ln_name, role, want_dense=False, opts: StepOpts = DEFAULT_OPTS):
        """LayerNorm of x (B,S,C) -> (fp32 or None, planes)."""
        p = self.p
        B, S, C = x.shape
        pl = self._seq_planes(opts.lane, B, S, C, role)
        d = ops.layernorm(x, p[ln_name + "/gamma"], p[ln_name + "/beta"], planes=pl, want_dense=want_dense)
        return d, pl

    def dense_planes(self, planes, name, seq_shape, act=None, bias_name=None, conv_map=False):
        """Dense over the channels of a planes tensor -> fp32 (B, S, Dout): planes holds either plain rows
        (B*S, C) or (conv_map) the ResNet's flat-pad (B, H', W') map read as CNN2SEQ does (model.py:252)."""
        from . import tc
        p = self.p
        y = tc.dense_tc(planes, p[name + "/w_tc"], p[bias_name or (name + "/bias")], act=act, nopad=not conv_map)
        B, S = seq_shape
        return y.reshape(B, S, -1)

    def bigru_planes(self, planes, name, seq_shape, seq=True, opts: StepOpts = DEFAULT_OPTS):
        p = self.p
        xp = self.dense_planes(planes, name, seq_shape, bias_name=name + "/ibias_cat")      # (B,S,6u)
        B, S = seq_shape
        return ops.bigru(xp.reshape(B, S, 2, -1), p[name + "/rec"], p[name + "/rbias"], seq=seq, nb=opts.gru_nb)

    def embed(self, x):
        p = self.p
        W, b = p["AR_EMBEDDING/kernel_folded"], p["AR_EMBEDDING/bias_folded"]
        if W.shape[0] >= 2048:
            return ops.gemm_splitk(x, W, b)
        return ops.dense(x, W, b)

    # ------------------------------------------------------------------ CUDA-graph replay
    def forward_graphed(self, inputs: Dict[str, torch.Tensor], tag="", sink=None,
                        opts: StepOpts = DEFAULT_OPTS, decode_only: bool = False) -> Dict[str, torch.Tensor]:
        """Same as forward(), but the ~45 launches of a step are captured once per input signature
        into a CUDA graph and replayed (the step is launch-bound at small batches).  Inputs are
        copied into the graph's static buffers; the returned tensors are the graph's static outputs
        (valid until the next replay of the same signature)."""
        key = (tag, opts, decode_only) + tuple(sorted((k, tuple(v.shape), str(v.dtype)) for k, v in inputs.items()))
        entry = self._graphs.get(key)
        if entry is None:
            static_in = {k: torch.empty_like(v, device=self.device).copy_(v) for k, v in inputs.items()}   # inputs may be pinned host tensors
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):           # warm-up: allocates plane buffers, sets func attributes
                for _ in range(2):
                    self.forward(static_in, sink=sink, opts=opts, decode_only=decode_only)
            cur.wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = self.forward(static_in, sink=sink, opts=opts, decode_only=decode_only)
            entry = (graph, static_in, static_out)
            if len(self._graphs) >= 16:
                self._graphs.pop(next(iter(self._graphs)))
            self._graphs[key] = entry
        graph, static_in, static_out = entry
        for k, v in inputs.items():
            static_in[k].copy_(v, non_blocking=True)
        graph.replay()
        return static_out

    # ------------------------------------------------------------------ concurrent micro-batch lanes
    def forward_lanes(self, inputs: Dict[str, torch.Tensor], lanes: int, tag="") -> Dict[str, torch.Tensor]:
        """One step as `lanes` contiguous micro-batches, each a captured CUDA graph replayed on a stream of its own.
        Utterances are independent on this path (inference BN), so results are bitwise those of the unsplit step.
        Why: at B=64 most kernels are latency-bound and leave SMs idle (stage 3-4 convs fill 68-99 of 148 SMs, the
        Bi-GRU recurrence is 48 dependent steps); a second lane's kernels fill them, and with host inputs lane
        k+1's H2D runs under lane k's compute.  Every lane writes its rows of shared per-sample output buffers
        (`sink`); the batch loss vector is one sar_loss_reduce_fwd over all rows after the lanes join."""
        B = int(inputs["x_data"].shape[0])
        lanes = max(1, min(int(lanes), B))
        if lanes == 1:
            return self.forward_graphed(inputs, tag)
        cur = torch.cuda.current_stream()
        while len(self._lane_streams) < lanes:
            self._lane_streams.append(torch.cuda.Stream(device=self.device))
            self._lane_done.append(torch.cuda.Event())
        if self._lane_in is None:
            self._lane_in = torch.cuda.Event()
        bounds = [(l * B // lanes, (l + 1) * B // lanes) for l in range(lanes)]
        sink = self._lane_sinks.get((B, lanes))
        self._lane_in.record(cur)
        for l, (b0, b1) in enumerate(bounds):
            st = self._lane_streams[l]
            st.wait_event(self._lane_in)
            # lanes share the SMs: per-layer launches only (a chain launch needs its whole grid resident)
            lopts = StepOpts(lane=l + 1, no_chain=True)          # lane 0 is the unsplit path
            with torch.cuda.stream(st):
                sl = {k: v[b0:b1] for k, v in inputs.items()}
                if sink is None:                             # first call: learn the per-sample outputs from an eager probe
                    probe = self.forward({k: v.to(self.device) for k, v in sl.items()}, opts=lopts)
                    st.synchronize()
                    sink = {k: torch.zeros((B,) + tuple(v.shape[1:]), device=self.device, dtype=v.dtype)
                            for k, v in probe.items() if k != "loss_vector" and not k.startswith("__")}
                    if "_ctc_loss" in sink:                  # (B,1) model output = the kernel's (B,) vector
                        sink["y_ctc_loss"] = sink["_ctc_loss"].view(B, 1)
                    self._lane_sinks[(B, lanes)] = sink
                self.forward_graphed(sl, tag=(tag, "lanes", B, lanes), sink={k: t[b0:b1] for k, t in sink.items()}, opts=lopts)
                self._lane_done[l].record(st)
        for l in range(lanes):
            cur.wait_event(self._lane_done[l])
        out = {k: v for k, v in sink.items() if not k.startswith("_")}
        out["loss_vector"] = ops.loss_reduce(sink.get("_sample_stats"), sink.get("_ctc_loss"), sink.get("_bn_stats"), B=B)
        return out

    def forward_slot(self, inputs: Dict[str, torch.Tensor], slot: int, tag=""):
        """Enqueue one whole step on pipeline slot `slot`: its own stream, CUDA graph and activation buffers, so
        that consecutive (independent) batches overlap -- the tail of a step (Bi-GRU on 64 SMs, VLAD / head / Dense on
        24-96 CTAs) leaves most SMs idle, and the next batch's stem and stage-1 convolutions fill them.  Returns
        (outputs, stream); the caller orders the inputs before and the consumers after with events on that stream.
        Kernels captured for a slot are chosen for SM-time, not latency (SLOT_OPTS): no stage chains (a chain holds
        its SMs for a whole stage, and two concurrent chains could starve each other of SMs) and 32 utterances per
        Bi-GRU cluster (32 SMs instead of 64 at B=64); measured 98.1 -> 100.9 k utt/s device, 88.2 -> 95.2 k e2e,
        while the same choices cost a single stream 11 %."""
        st = self.slot_stream(slot)
        with torch.cuda.stream(st):
            out = self.forward_graphed(inputs, tag=(tag, "slot"), opts=SLOT_OPTS(slot))
        return out, st

    def slot_stream(self, slot: int) -> torch.cuda.Stream:
        while len(self._lane_streams) <= slot:
            self._lane_streams.append(torch.cuda.Stream(device=self.device))
            self._lane_done.append(torch.cuda.Event())
        return self._lane_streams[slot]

    def forward(self, inputs: Dict[str, torch.Tensor], want_intermediates: bool = False, sink=None,
                opts: StepOpts = DEFAULT_OPTS, decode_only: bool = False) -> Dict[str, torch.Tensor]:
        """`sink` (forward_lanes): per-sample outputs are written into these preallocated rows and the batch loss
        vector is left to the caller.  `decode_only` (model.ctc_pred on the x_data-only sub-model, model.py:385-389):
        encoder + ASR branch up to the ctc_pred logits, no CTC loss, no accent branch, no label inputs."""
        cfg, p = self.cfg, self.p
        x = inputs["x_data"]
        if x.dim() == 3:
            x = x.unsqueeze(-1)
        B = x.shape[0]
        out: Dict[str, torch.Tensor] = {}
        S, Cc = self.plan.seq_len, self.plan.cout
        if self.dense_tc:
            return self._forward_tc(inputs, x, want_intermediates, sink, opts, decode_only)
        if self.conv_path == "tc":
            raw = self.resnet.forward(x, opts=opts)                             # final BN->ReLU already applied
            cnn = self.dense_ln(raw.reshape(B, S, Cc), "CNN_LIN", "CNN_LIN_LN")   # CNN2SEQ, model.py:252
        else:
            raw = self.resnet.forward_raw(x)
            # final ResNet BN->ReLU (resnet.py:178/196) fused as the input op of CNN_LIN
            cnn = self.dense_ln(raw.reshape(B, S, Cc), "CNN_LIN", "CNN_LIN_LN",
                                pre=self.resnet.bn(self.plan.final_bn))
        crnn = ops.layernorm(self.bigru(cnn, "CRNN", opts=opts), p["CRNN_LN/gamma"], p["CRNN_LN/beta"])
        if want_intermediates:
            out["resnet_raw"], out["cnn_lin"], out["crnn"] = raw, cnn, crnn
        return self._forward_tail(inputs, out, crnn, None, want_intermediates, sink, opts, decode_only)

    def _forward_tc(self, inputs, x, want_intermediates, sink, opts, decode_only):
        """Encoder with every Dense / GRU input projection on the tensor cores: activations between the
        LayerNorms and the Dense layers travel as fp16 hi/lo planes (no fp32 round trip)."""
        B = x.shape[0]
        S = self.plan.seq_len
        out: Dict[str, torch.Tensor] = {}
        wi = want_intermediates
        P0 = self.resnet.forward(x, as_planes=True, opts=opts)                  # relu(final BN), resnet.py:178/196
        y = self.dense_planes(P0, "CNN_LIN", (B, S), act="tanh", conv_map=True)      # CNN2SEQ + CNN_LIN, model.py:252-253
        cnn, P1 = self.ln_planes(y, "CNN_LIN_LN", "cnn", want_dense=wi, opts=opts)
        crnn, P2 = self.ln_planes(self.bigru_planes(P1, "CRNN", (B, S), opts=opts), "CRNN_LN", "crnn", want_dense=wi, opts=opts)
        if wi:
            out["resnet_raw"], out["cnn_lin"], out["crnn"] = self.tc_unpack(P0), cnn, crnn
        return self._forward_tail(inputs, out, crnn, P2, want_intermediates, sink, opts, decode_only)

    def tc_unpack(self, planes):
        from . import tc
        return tc.unpack(planes)

    # -- the two branches after CRNN_LN (model.py:261-269 and 275-322): independent of each other
    def _ctc_branch(self, inputs, out, crnn, crnn_planes, want_intermediates, sink, opts, decode_only):
        cfg, p = self.cfg, self.p
        B = inputs["x_data"].shape[0]
        S_ = self.plan.seq_len
        if crnn_planes is not None:
            _, P3 = self.ln_planes(self.bigru_planes(crnn_planes, "CTC_BIGRU", (B, S_), opts=opts), "CTC_BIGRU_LN",
                                   "ctc_bigru", opts=opts)
            y = self.dense_planes(P3, "CTC_DS", (B, S_), act="tanh")
            if "ctc_pred/w_tc" in p:
                _, P5 = self.ln_planes(y, "CTC_DS_LN", "ctc_ds", opts=opts)
                asr = None
            else:
                asr = ops.layernorm(y, p["CTC_DS_LN/gamma"], p["CTC_DS_LN/beta"])
        else:
            asr = ops.layernorm(self.bigru(crnn, "CTC_BIGRU", opts=opts), p["CTC_BIGRU_LN/gamma"], p["CTC_BIGRU_LN/beta"])
            asr = self.dense_ln(asr, "CTC_DS", "CTC_DS_LN")
        if asr is None:
            logits = self.dense_planes(P5, "ctc_pred", (B, S_), bias_name="ctc_pred/bias_tc")
        else:
            logits = ops.dense(asr, p["ctc_pred/kernel"], p["ctc_pred/bias"])
        out["__ctc_logits"] = logits                 # (B, S, ld) pre-softmax, for the greedy decode (model.ctc_pred)
        if decode_only:
            return
        ctc_loss, status, probs = ops.ctc(logits, inputs["x_ctc_label"], inputs["x_ctc_in_len"],
                                          inputs["x_ctc_out_len"], want_probs=want_intermediates, classes=cfg.bpe_classes,
                                          loss=sink.get("_ctc_loss"), status=sink.get("ctc_status"))
        out["_ctc_loss"] = ctc_loss
        out["y_ctc_loss"] = ctc_loss.reshape(B, 1)
        out["ctc_status"] = status
        if want_intermediates:
            out["ctc_pred"] = probs

    def _ar_branch(self, inputs, out, crnn, crnn_planes, want_intermediates, sink, opts):
        cfg, p = self.cfg, self.p
        B = inputs["x_data"].shape[0]
        P4 = None
        S_ = self.plan.seq_len
        G = cfg.ghost_clusters if cfg.mto == "gvlad" else 0
        vlad_on_tc = False
        if crnn_planes is not None:
            y = self.dense_planes(crnn_planes, "AR_DS", (B, S_), act="tanh")
            if cfg.mto in ("vlad", "gvlad") and (cfg.mto + "/w_assign_tc") in p:
                from . import tc
                vlad_on_tc = tc.vlad_tc_supported(B, S_, y.shape[-1], cfg.vlad_clusters, G)
            if cfg.mto == "bigru" or vlad_on_tc:
                ar, P4 = self.ln_planes(y, "AR_DS_LN", "ar", want_dense=want_intermediates, opts=opts)
            else:
                ar = ops.layernorm(y, p["AR_DS_LN/gamma"], p["AR_DS_LN/beta"])
        else:
            ar = self.dense_ln(crnn, "AR_DS", "AR_DS_LN")
        if cfg.mto == "avg":
            integ = ops.avgpool(ar)
        elif cfg.mto == "bigru":
            integ = (self.bigru_planes(P4, "AR_MERGE", (B, S_), seq=False, opts=opts) if P4 is not None
                     else self.bigru(ar, "AR_MERGE", seq=False, opts=opts))
        else:
            vplanes = None
            if getattr(self, "embed_ksplit", 0):
                from . import tc
                key = (opts.lane, B, cfg.vlad_clusters * cfg.hidden_dim, "vlad")
                if key not in self._seq_bufs:
                    self._seq_bufs[key] = tc.alloc_rows(B, key[2], self.device)
                vplanes = self._seq_bufs[key]
            seg = self._seg_begin("vlad")
            if vlad_on_tc:                       # both contractions on the tensor cores (csrc/vlad_tc.cu)
                integ = tc.vlad_tc(P4, p[cfg.mto + "/w_assign_tc"], p[cfg.mto + "/b_assign"], p[cfg.mto + "/centers"],
                                   B, S_, cfg.vlad_clusters, G, planes=vplanes,
                                   want_dense=want_intermediates or vplanes is None)
            else:                                # shapes outside the tensor-core kernel's tiles: CUDA-core kernel (vlad.cu)
                integ = ops.vlad(ar, p[cfg.mto + "/w_assign"], p[cfg.mto + "/b_assign"], p[cfg.mto + "/centers"],
                                 cfg.vlad_clusters, G, planes=vplanes, want_dense=want_intermediates or vplanes is None)
            self._seg_end(seg)
        if cfg.mto in ("vlad", "gvlad") and getattr(self, "embed_ksplit", 0):
            from . import tc
            emb = tc.gemm_splitk_tc(vplanes, p["AR_EMBEDDING/w_tc"], p["AR_EMBEDDING/bias_folded"],
                                    p["AR_EMBEDDING/zero_bias"], self.embed_ksplit, out=sink.get("embedding"))
        else:
            emb = self.embed(integ)
        if want_intermediates:
            out["ar_ds"], out["integration"] = ar, integ
        out["embedding"] = emb
        onehot = inputs.get("x_accent") if cfg.disc_enable else inputs.get("y_true")
        cls = tuple(p[k] for k in ("AR_CF_DS1/kernel", "AR_CF_DS1/bias", "AR_CF_DS2/kernel", "AR_CF_DS2/bias",
                                   "y_accent/kernel", "y_accent/bias"))
        h = ops.head(emb, cls, wd=p.get("y_disc/w") if cfg.disc_enable else None, onehot=onehot,
                     n_classes=cfg.accent_classes, head_kind=cfg.metric_loss if cfg.disc_enable else None,
                     margin=cfg.margin,
                     out={k: sink.get(n) for k, n in (("y_accent", "y_accent"), ("y_accent_logits", "y_accent_logits"),
                                                      ("y_disc", "y_disc"), ("y_disc_logits", "y_disc_logits"),
                                                      ("sample_stats", "_sample_stats"))})
        out["_sample_stats"] = h.pop("sample_stats")
        out.update(h)
        if cfg.disc_enable and cfg.bn_dim:
            bn = ops.dense(emb, p["AR_BN_DS/kernel"], p["AR_BN_DS/bias"], act="relu")
            bn = ops.dense(bn, p["bottleneck/kernel_folded"], p["bottleneck/bias_folded"])
            hb = ops.head(None, None, emb_d=bn, wd=p["y_disc_bn/w"], onehot=onehot, n_classes=cfg.accent_classes,
                          head_kind=cfg.metric_loss, margin=cfg.margin)
            out["_bn_stats"] = hb["sample_stats"]
            out["y_disc_bn"] = hb["y_disc"]

    def _forward_tail(self, inputs, out, crnn, crnn_planes, want_intermediates, sink_in, opts, decode_only=False):
        cfg = self.cfg
        B = inputs["x_data"].shape[0]
        sink = sink_in or {}
        if decode_only:
            if not cfg.ctc_enable:
                raise ValueError("decode_only needs a model built with ctc_enable=True")
            self._ctc_branch(inputs, out, crnn, crnn_planes, False, sink, opts, True)
            return out
        # While the step is being captured into a CUDA graph the CTC branch goes onto a second stream: fork after
        # CRNN_LN, join before the loss reduction -> two parallel branches of the graph (CTC_BIGRU's 48-75 dependent
        # steps run beside AR_DS -> VLAD -> embedding -> head instead of in front of them).  Eager launches stay on
        # one stream (their temporaries belong to the caller's stream).
        fork = (self.parallel_branches and cfg.ctc_enable and cfg.ar_enable and torch.cuda.is_current_stream_capturing())
        if cfg.ctc_enable:                                                      # model.py:261-269
            if fork:
                cur = torch.cuda.current_stream()
                side = self._branch_streams.get(opts.lane)
                if side is None:
                    side = self._branch_streams[opts.lane] = torch.cuda.Stream(device=self.device)
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    self._ctc_branch(inputs, out, crnn, crnn_planes, want_intermediates, sink, opts, False)
            else:
                self._ctc_branch(inputs, out, crnn, crnn_planes, want_intermediates, sink, opts, False)
        if cfg.ar_enable:                                                       # model.py:275-322
            self._ar_branch(inputs, out, crnn, crnn_planes, want_intermediates, sink, opts)
        if fork:
            torch.cuda.current_stream().wait_stream(side)
        if sink_in is not None:
            # anything a kernel did not write in place (the views of y_ctc_loss share _ctc_loss's rows)
            for k, dst in sink_in.items():
                src = out.get(k)
                if src is not None and src.data_ptr() != dst.data_ptr():
                    dst.copy_(src)
            return dict(sink_in)
        out["loss_vector"] = ops.loss_reduce(out.get("_sample_stats"), out.get("_ctc_loss"), out.get("_bn_stats"), B=B)
        return out
