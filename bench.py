#!/usr/bin/env python
"""bench.py -- utterances/sec of the SAR-Net forward path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One "step" = one forward of the hot path over one per-GPU batch of synthetic fbank input
(BASELINE.json configs[1] by default: B=64 x 500 frames x 80 mel, thin ResNet-34 + Bi-GRU +
GhostVLAD(64c/8g) + ArcFace).  Weak scaling: every rank processes its own B utterances; the
step ends with the path's single collective, an all-reduce(SUM) of the 8-float loss vector.

Printed JSON (rank 0, ONE line):
  value          whole-job utterances/s with inputs resident in HBM (CUDA events, max over ranks), K steps
  sustained      the same loop run for >= 2 s
  single_stream  the same steps strictly one after the other on one stream (per-batch latency)
  e2e            same metric through model.predict_generator() with HOST inputs in pinned memory: H2D of every
                 step's inputs and D2H of its outputs inside the timed region; e2e.blocking_predict = model.predict
  e2e_pcm        host 16-bit PCM -> on-device fbank front-end -> forward (the metric's "fbank ->" stage included)
  roofline       residual-block convolution kernels: algorithmic FLOPs / bytes per step over their device time,
                 measured live inside a replayed CUDA graph (event-record nodes)
  roofline_vlad  the GhostVLAD kernel (north_star names it): isolated and in-graph, at B=64 and B=512
  roofline_fbank the front-end kernels
  extra_configs  the other BASELINE.json configurations at this N: configs[2] (B=256, T=800, CTC+Circle),
                 configs[3] shard (NetVLAD+CosFace, 256/GPU), configs[4] shard (512/GPU): value / single_stream /
                 e2e / conv roofline each
  strong         configs[4] strong scaling: global B=4096 split over the N ranks, micro-batched (512/step)
  train_slice    SURVEY 8f-1: one optimisation step of everything above the frozen ResNet (accent path), B=64/GPU
  train_full     SURVEY 8f-1: one optimisation step of the WHOLE model, multi-task (correctness-first ResNet backward), B=64/GPU
  cpu_baseline   the oracle's torch-CPU fp32 restatement of the Keras forward on the host cores (bounded sample)
`--impl reference` times that CPU restatement as the reference arm (the literal Keras/TF graph cannot run here: no
tensorflow/keras in the image and CuDNNGRU has no CPU kernel).  `--quick` skips extra_configs / strong / sustained.
"""
from __future__ import annotations

import argparse
import contextlib
import hashlib
import io
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

METRIC = "utterances/sec (fbank->ResNet->GhostVLAD->margin-softmax fwd)"
UNIT = "utt/s"

THIN34 = dict(res_type="res34", res_filters=32)
CONFIGS = {
    # BASELINE.json configs[1]
    "cfg2": dict(B=64, T=500, kw=dict(ctc_enable=False, ar_enable=True, disc_enable=True, mto="gvlad", vlad_clusters=64,
                                      ghost_clusters=8, metric_loss="arcface", margin=0.3, **THIN34),
                 workload="configs[1]: B=64/GPU x 500 frames x 80 mel, thin-ResNet34+BiGRU+GhostVLAD(64c/8g)+ArcFace fwd"),
    # BASELINE.json configs[2]
    "cfg3": dict(B=256, T=800, kw=dict(ctc_enable=True, ar_enable=True, disc_enable=True, mto="bigru",
                                       metric_loss="circleloss", margin=0.2, **THIN34),
                 workload="configs[2]: B=256 x 200-800 frames zero-padded to 800, multi-task CTC+Circle-Loss fwd"),
    # BASELINE.json configs[3] per-GPU shard (1024 over 4 GPUs)
    "cfg4": dict(B=256, T=500, kw=dict(ctc_enable=False, ar_enable=True, disc_enable=True, mto="vlad", vlad_clusters=64,
                                       metric_loss="cosface", margin=0.3, **THIN34),
                 workload="configs[3] shard: B=256/GPU (1024 over 4 GPUs) x 500 frames, NetVLAD(64c)+CosFace fwd"),
    # BASELINE.json configs[4] per-GPU shard (4096 over 8 GPUs)
    "cfg5": dict(B=512, T=500, kw=dict(ctc_enable=True, ar_enable=True, disc_enable=True, mto="gvlad", vlad_clusters=64,
                                       ghost_clusters=8, metric_loss="circleloss", margin=0.2, **THIN34),
                 workload="configs[4] shard: B=512/GPU (4096 over 8 GPUs) x 500 frames, CRNN+GhostVLAD+Circle-Loss+CTC fwd"),
}
ACT_BYTES = 4.0            # activations as stored by this build: fp16 hi + fp16 lo planes


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "tflops": float(p["bf16_tflops_sustained"]),
                "tflops_burst": float(p["bf16_tflops"]), "source": "measured (MEASURED_PEAKS.json)"}
    except Exception:
        return {"hbm_gbs": 6650.0, "tflops": 1400.0, "tflops_burst": 1590.0, "source": "fallback (B200_PROFILING.md)"}


def build_id(files=None) -> str:
    """Hash of the CUDA sources (all of them, or the named ones): ties committed ncu figures (profiles/*.json) to the
    build they were taken on."""
    h = hashlib.sha1()
    d = os.path.join(ROOT, "aesrc2020_b200", "csrc")
    for fn in sorted(os.listdir(d)):
        if fn.endswith((".cu", ".cuh")) and (files is None or fn in files):
            with open(os.path.join(d, fn), "rb") as f:
                h.update(fn.encode() + f.read())
    return h.hexdigest()[:12]


def conv_build_id() -> str:
    """The sources the residual-block convolution kernels are compiled from (roofline.traffic is tied to these)."""
    return build_id(("conv_tc.cu", "tc_common.cuh", "common.cuh"))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


def conv_algorithmic_work(plan, B, act_bytes):
    """per-layer F = 2*B*Ho*Wo*Cout*Cin*kh*kw; Bytes = e*B*(in + out [+ residual]) + weights.
    Returns (flops, bytes, launches) per step for the residual-block convolutions (stem excluded)."""
    F = 0.0
    Bt = 0.0
    n = 0
    for blk in plan.blocks:
        for c, has_res in ((blk.conv1, False), (blk.conv2, True), (blk.short, False)):
            if c is None:
                continue
            F += 2.0 * B * c.hout * c.wout * c.cout * c.cin * c.kh * c.kw
            Bt += act_bytes * B * (c.hin * c.win * c.cin + c.hout * c.wout * c.cout * (2 if has_res else 1))
            Bt += 4.0 * (c.kh * c.kw * c.cin * c.cout + c.cout)
            n += 1
    return F, Bt, n


def conv_issued_flops(plan, B):
    """Tensor-core FLOPs the residual-block kernels actually ISSUE per step: the three fp16 hi/lo products (A_hi B_hi,
    A_lo B_hi, A_hi B_lo -- DESIGN.md 4) on whole 128-row tiles of the flat-pad layout ((H+1)(W+1) rows per image)."""
    F = 0.0
    for blk in plan.blocks:
        for c in (blk.conv1, blk.conv2, blk.short):
            if c is None:
                continue
            rows = B * (c.hout + 1) * (c.wout + 1)
            F += 3 * 2.0 * (-(-rows // 128) * 128) * c.cout * c.cin * c.kh * c.kw
    return F


def vlad_algorithmic_work(B, S, D, K, G, e=ACT_BYTES):
    """Bytes = e*B*S*D + 4*B*K*D + 4*(2*(K+G)*D + (K+G)); F = 4*B*S*D*(K+G)."""
    return 4.0 * B * S * D * (K + G), e * B * S * D + 4.0 * B * K * D + 4.0 * (2 * (K + G) * D + (K + G))


def workload_config(cfgd, B, T, plan, in_bytes, rotate, world):
    """The `config` object of the JSON line -- built by the SAME function for both arms (`--impl ours` / `--impl reference`),
    so the driver's same-config check compares equal dicts; arm-specific details live in other keys of the line."""
    act_mb = conv_algorithmic_work(plan, B, ACT_BYTES)[1] / 1e6 * 2 / 3
    return {"workload": cfgd["workload"], "per_gpu_batch": B, "frames": T, "seq_len": plan.seq_len,
            "l2": "inputs rotate over %d distinct batches (%.0f MB > 126 MB L2); activations %.0f MB/step"
                  % (rotate, rotate * in_bytes / 1e6, act_mb),
            "parallelism": "dp%d (batch-sharded, 1 all-reduce of 8 floats/step)" % world}


def conv_roofline(plan, B, conv_ms, step_ms, peaks):
    F, Bt, nconv = conv_algorithmic_work(plan, B, ACT_BYTES)
    t = conv_ms * 1e-3
    tensor_time, hbm_time = F / (peaks["tflops"] * 1e12), Bt / (peaks["hbm_gbs"] * 1e9)
    if hbm_time >= tensor_time:
        roof = {"bound": "hbm", "achieved": Bt / t / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s"}
    else:
        roof = {"bound": "tensor", "achieved": F / t / 1e12, "peak": peaks["tflops"], "unit": "TFLOP/s"}
    roof["frac"] = roof["achieved"] / roof["peak"]
    roof.update({"conv_ms_per_step": conv_ms, "launches_per_step": nconv, "avg_launch_us": conv_ms * 1e3 / nconv,
                 "algorithmic_bytes_per_launch": Bt / nconv, "algorithmic_mb_per_step": Bt / 1e6,
                 "algorithmic_gflop_per_step": F / 1e9, "tflops_achieved": F / t / 1e12,
                 "frac_of_tensor_peak": F / t / 1e12 / peaks["tflops"], "frac_at_e2": (Bt - 0.0) / 2.0 / t / 1e9 / peaks["hbm_gbs"],
                 "act_bytes": ACT_BYTES, "share_of_step": conv_ms / step_ms if step_ms else None,
                 "ms_per_step_single_stream_graph": step_ms})
    # what bounds these kernels in practice: the fp32-accurate hi/lo scheme issues 3 fp16 products per algorithmic one
    # (+ flat-pad rows), and the chip's tensor pipes are power-limited to the measured cuBLAS rate -- the ceiling of
    # `frac` is hbm_time / (issued / sustained tensor rate), not 1
    Fi = conv_issued_flops(plan, B)
    roof["issued_gflop_per_step"] = Fi / 1e9
    roof["issued_tflops_achieved"] = Fi / t / 1e12
    roof["frac_issued_of_tensor_peak"] = Fi / t / 1e12 / peaks["tflops"]
    roof["frac_ceiling_at_tensor_peak"] = hbm_time / (Fi / (peaks["tflops"] * 1e12))
    return roof


# ====================================================================================== one configuration
class Runner:
    """One model configuration on this rank: synthetic batches, the timed loops."""

    def __init__(self, name, cfgd, dev, rank, world, rotate, eager=False, pipeline=4):
        from aesrc2020_b200 import model as mdl, utils as us
        self.name, self.cfgd, self.dev, self.rank, self.world = name, cfgd, dev, rank, world
        self.B, self.T = cfgd["B"], cfgd["T"]
        with contextlib.redirect_stdout(io.StringIO()):
            self.model, _ = mdl.SAR_Net((self.T, 80, 1), device=dev, **cfgd["kw"])
        self.cfg = self.model.config
        self.eng = self.model.engine()
        self.plan = self.cfg.plan()
        self.rotate = rotate
        self.eager = eager
        self.depth = 1 if eager else max(1, pipeline)
        self.host, self.devb = [], []
        for i in range(rotate):
            lengths = None
            if name == "cfg3":       # variable 200-800 frames, zero-padded (reference semantics Q4: compute on padding)
                lengths = np.random.RandomState(100 + i + 1000 * rank).randint(200, 801, size=self.B)
            x, _ = us.synthetic_batch(self.cfg, self.B, seed=2020 + i + 1000 * rank, lengths=lengths)
            self.host.append(us.pinned_like(x))      # e2e: inputs start in pinned host memory (a loader's ring buffer)
            self.devb.append({k: self.model._to_device(k, v).clone() for k, v in x.items()})
        torch.cuda.synchronize()
        self.in_bytes = sum(v.nbytes for v in self.host[0].values())
        if eager:
            self.model.use_graph = False
        self.main = torch.cuda.current_stream()

    def close(self):
        self.model = self.eng = self.host = self.devb = None
        import gc
        gc.collect()
        torch.cuda.empty_cache()

    # ---- helpers
    def barrier(self):
        if self.world > 1:
            import torch.distributed as tdist
            tdist.barrier()
        torch.cuda.synchronize()

    def step(self, i, graphed=True):
        import torch.distributed as tdist
        x = self.devb[i % self.rotate]
        if graphed and self.depth > 1:              # consecutive batches alternate between `depth` streams / graphs
            out, st = self.eng.forward_slot(x, i % self.depth)
            if self.world > 1:
                with torch.cuda.stream(st):
                    tdist.all_reduce(out["loss_vector"])
            return out
        out = self.eng.forward_graphed(x) if (graphed and not self.eager) else self.eng.forward(x)
        if self.world > 1:
            tdist.all_reduce(out["loss_vector"])
        return out

    def join_slots(self):
        for st in self.eng._lane_streams[:self.depth] if self.depth > 1 else []:
            self.main.wait_stream(st)

    def fork_slots(self):
        for st in (self.eng._lane_streams[:self.depth] if self.depth > 1 else []):
            st.wait_stream(self.main)               # no slot starts before the start event

    def max_over_ranks(self, *vals):
        if self.world == 1:
            return vals if len(vals) > 1 else vals[0]
        import torch.distributed as tdist
        tt = torch.tensor(list(vals), device=self.dev, dtype=torch.float64)
        tdist.all_reduce(tt, op=tdist.ReduceOp.MAX)
        out = [float(v) for v in tt.tolist()]
        return out if len(out) > 1 else out[0]

    # ---- launches per step (our kernels, counted through the C ABI wrappers)
    def count_launches(self):
        from aesrc2020_b200 import ops
        from aesrc2020_b200.engine import SLOT_OPTS
        l0 = ops.LAUNCHES["n"]
        if self.depth > 1:
            self.eng.forward(self.devb[0], opts=SLOT_OPTS(1))       # the kernel sequence a pipeline slot captures
        else:
            self.eng.forward(self.devb[0])
        torch.cuda.synchronize()
        return ops.LAUNCHES["n"] - l0

    def warm(self, warmup):
        self.step(0, graphed=False)
        for i in range(self.depth if self.depth > 1 else 0):        # capture every slot's CUDA graph first
            self.step(i)
        for i in range(warmup):
            self.step(i)
        self.barrier()

    # ---- timed region A: K steps, device-resident inputs, `depth` batches in flight
    def time_value(self, steps, first=0):
        self.barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        self.fork_slots()
        for i in range(steps):
            self.step(first + i)
        self.join_slots()
        t1.record()
        self.barrier()
        ms = self.max_over_ranks(t0.elapsed_time(t1))
        return ms / steps

    # ---- the same steps strictly one after the other on ONE stream
    def time_single(self, steps, first=0):
        import torch.distributed as tdist
        for i in range(3):
            self.eng.forward_graphed(self.devb[i % self.rotate])
        self.barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for i in range(steps):
            o = self.eng.forward_graphed(self.devb[(first + i) % self.rotate])
            if self.world > 1:
                tdist.all_reduce(o["loss_vector"])
        s1.record()
        self.barrier()
        return self.max_over_ranks(s0.elapsed_time(s1)) / steps

    # ---- timed region B: the step captured once more with external event pairs around the block convs (and VLAD)
    def time_segments(self, steps, first=0):
        conv_ev = []
        self.eng.resnet.record_events = conv_ev
        self.eng.segment_events = {"vlad": []}
        fn = (lambda i: self.eng.forward(self.devb[i % self.rotate])) if self.eager else \
             (lambda i: self.eng.forward_graphed(self.devb[i % self.rotate], tag="segments"))
        fn(0)
        torch.cuda.synchronize()
        conv, vlad, step = [], [], []
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(steps):
            a0.record()
            fn(first + i)
            a1.record()
            torch.cuda.synchronize()
            if conv_ev:
                conv.append(conv_ev[-1][0].elapsed_time(conv_ev[-1][1]))
            ve = self.eng.segment_events["vlad"]
            if ve:
                vlad.append(ve[-1][0].elapsed_time(ve[-1][1]))
            step.append(a0.elapsed_time(a1))
        self.eng.resnet.record_events = None
        self.eng.segment_events = None
        self.barrier()
        return (float(np.mean(conv)) if conv else None, float(np.mean(vlad)) if vlad else None, float(np.mean(step)))

    # ---- e2e: the public API with HOST inputs
    def time_e2e(self, steps):
        def gen(first, count):
            for i in range(count):
                yield self.host[(first + i) % self.rotate]
        outs = self.model.predict_generator(gen(0, 3))
        out_bytes = sum(o.nbytes for o in (outs if isinstance(outs, list) else [outs])) // 3
        self.barrier()
        t0 = time.perf_counter()
        self.model.predict_generator(gen(3, steps))
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        for i in range(2):
            self.model.predict(self.host[i % self.rotate], batch_size=self.B)
        self.barrier()
        t0 = time.perf_counter()
        nb = max(3, steps // 2)
        for i in range(nb):
            self.model.predict(self.host[(3 + i) % self.rotate], batch_size=self.B)
        torch.cuda.synchronize()
        sync_s = (time.perf_counter() - t0) * steps / nb
        e2e_s, sync_s = self.max_over_ranks(e2e_s, sync_s)
        n = self.world * self.B * steps
        return {"value": n / e2e_s, "unit": UNIT, "h2d_bytes_per_step": self.in_bytes, "d2h_bytes_per_step": out_bytes,
                "blocking_predict": n / sync_s,
                "api": "model.predict_generator(batches in pinned host memory, steps=K): %d batches in flight (copy stream + one "
                       "compute stream/graph per slot), wall clock over K steps" % self.model.PIPE_DEPTH}

    def measure(self, steps, warmup, peaks, e2e=True):
        """value / single_stream / conv roofline / e2e of this configuration."""
        self.warm(warmup)
        single_ms = self.time_single(steps, warmup) if self.depth > 1 else None
        ms = self.time_value(steps, warmup)
        conv_ms, vlad_ms, step_ms = self.time_segments(min(steps, 10), warmup)
        res = {"workload": self.cfgd["workload"], "per_gpu_batch": self.B, "frames": self.T,
               "value": self.world * self.B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
               "pipeline": self.depth}
        if single_ms is not None:
            res["single_stream"] = {"ms_per_step": single_ms, "value": self.world * self.B / (single_ms * 1e-3)}
        if conv_ms:
            res["roofline"] = conv_roofline(self.plan, self.B, conv_ms, step_ms, peaks)
        if vlad_ms:
            res["vlad_ms_in_graph"] = vlad_ms
        if e2e:
            res["e2e"] = self.time_e2e(steps)
        return res


# ====================================================================================== the GhostVLAD kernel alone
def vlad_roofline(dev, peaks, runner64=None, vlad_ms_in_graph=None):
    """north_star: '>= 70 % of the per-kernel roofline on the ResNet-conv AND GhostVLAD paths'.  SURVEY 8d bytes
    (114.7 KB/utterance with X and the descriptor stored as fp16 hi+lo planes, e = 4 B): isolated (CUDA events around
    back-to-back launches, inputs rotated over > L2 worth of buffers at B=512) and inside the replayed step graph."""
    from aesrc2020_b200 import tc
    S, D, K, G = 48, 256, 64, 8
    rng = np.random.RandomState(5)
    wa = torch.from_numpy(tc.pack_vlad_assign((rng.randn(D, K + G) / 16 * 3).astype(np.float32))).to(dev)
    ba = torch.from_numpy((rng.randn(K + G) * 0.1).astype(np.float32)).to(dev)
    cen = torch.from_numpy((rng.randn(K + G, D) / 16).astype(np.float32)).to(dev)
    out = {"kernel": "vlad_tc_kernel (tcgen05 scores + residual GEMMs, fused softmax / L2 norm)", "bound": "hbm",
           "peak": peaks["hbm_gbs"], "unit": "GB/s", "bytes_per_utt": vlad_algorithmic_work(1, S, D, K, G)[1], "points": []}
    for B in (64, 512):
        nbuf = 8 if B == 512 else 16                   # 512: 8 x (25 + 34) MB of inputs/outputs > 126 MB L2
        xs = [tc.Planes((torch.randn(2, B * S, D, device=dev) * 0.5).half(), 1, B * S, 1, D, False) for _ in range(nbuf)]
        ys = [tc.alloc_rows(B, K * D, dev) for _ in range(nbuf)]
        for i in range(nbuf):
            tc.vlad_tc(xs[i], wa, ba, cen, B, S, K, G, planes=ys[i], want_dense=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for r in range(reps):
            for i in range(nbuf):
                tc.vlad_tc(xs[i], wa, ba, cen, B, S, K, G, planes=ys[i], want_dense=False)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (reps * nbuf)
        F, Bt = vlad_algorithmic_work(B, S, D, K, G)
        pt = {"B": B, "isolated_us": us, "achieved": Bt / (us * 1e-6) / 1e9, "frac": Bt / (us * 1e-6) / 1e9 / peaks["hbm_gbs"],
              "algorithmic_bytes_per_launch": Bt, "gflop_per_launch": F / 1e9}
        if vlad_ms_in_graph and B in vlad_ms_in_graph:
            g_us = vlad_ms_in_graph[B] * 1e3
            pt.update({"in_graph_us": g_us, "frac_in_graph": Bt / (g_us * 1e-6) / 1e9 / peaks["hbm_gbs"]})
        out["points"].append(pt)
        del xs, ys
        torch.cuda.empty_cache()
    p512 = [p for p in out["points"] if p["B"] == 512][0]
    out["achieved"] = p512.get("achieved")
    out["frac"] = p512.get("frac_in_graph", p512["frac"])
    out["note"] = "frac = the B=512 point inside the replayed step graph when available (events include launch gaps), else isolated"
    return out


# ====================================================================================== front-end (PCM -> x_data)
def fbank_bench(runner, steps, peaks):
    """roofline_fbank (Bytes = 2*n_samples [int16 PCM] + 4*T*80 per utterance; SURVEY 8d uses 4*n_samples for float
    waveforms) and e2e_pcm: pinned host PCM -> H2D -> sar_fbank_pcm16_fwd -> the step -> D2H of the outputs."""
    from aesrc2020_b200 import fbank as fb, ops
    dev, B, T = runner.dev, runner.B, runner.T
    n = fb.FRAME_LEN + (T - 1) * fb.FRAME_STEP                     # samples that give exactly T frames
    rng = np.random.RandomState(77 + runner.rank)
    nb = 4
    pcm_host = [torch.from_numpy((rng.randn(B * n) * 3000).clip(-32768, 32767).astype(np.int16)).pin_memory() for _ in range(nb)]
    offs = torch.from_numpy(np.arange(B + 1, dtype=np.int64) * n).to(dev)
    melfb = torch.from_numpy(np.ascontiguousarray(fb.mel_filterbank().T, dtype=np.float32)).to(dev)
    feat_ws = torch.empty((B, T, 80), device=dev, dtype=torch.float32)
    pcm_dev = [p.to(dev) for p in pcm_host]
    x_out = torch.empty((B, T, 80), device=dev, dtype=torch.float32)
    for i in range(3):
        ops.fbank(pcm_dev[i % nb], offs, melfb, T, T, out=x_out, feat_ws=feat_ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        ops.fbank(pcm_dev[i % nb], offs, melfb, T, T, out=x_out, feat_ws=feat_ws)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / steps
    Bt = B * (2.0 * n + 4.0 * T * 80)
    roof = {"kernel": "fbank_frame_kernel<int16> + fbank_norm_kernel (2 launches)", "bound": "hbm", "us_per_batch": us,
            "algorithmic_bytes_per_batch": Bt, "achieved": Bt / (us * 1e-6) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": Bt / (us * 1e-6) / 1e9 / peaks["hbm_gbs"], "samples_per_utt": n}
    # e2e from PCM: the slot pipeline of predict_generator, with the front-end in front of every step
    model, eng = runner.model, runner.eng
    labels = {k: v for k, v in runner.devb[0].items() if k != "x_data"}
    depth = 3
    streams = [eng.slot_stream(s) for s in range(depth)]
    stage = [torch.empty(B * n, dtype=torch.int16, device=dev) for _ in range(depth)]
    xin = [torch.empty((B, T, 80, 1), device=dev, dtype=torch.float32) for _ in range(depth)]
    fws = [torch.empty((B, T, 80), device=dev, dtype=torch.float32) for _ in range(depth)]
    names = model._outputs
    pins = [None] * depth
    done = [torch.cuda.Event() for _ in range(depth)]

    def submit(i):
        s = i % depth
        with torch.cuda.stream(streams[s]):
            stage[s].copy_(pcm_host[i % nb], non_blocking=True)
            ops.fbank(stage[s], offs, melfb, T, T, out=xin[s], feat_ws=fws[s])
            inp = dict(labels)
            inp["x_data"] = xin[s]
        out, st = eng.forward_slot(inp, s)
        with torch.cuda.stream(st):
            outs = [out[k] for k in names]
            tot = sum(o.numel() for o in outs)
            if pins[s] is None:
                pins[s] = torch.empty(tot, dtype=torch.float32, pin_memory=True)
            off = 0
            for o in outs:
                pins[s][off:off + o.numel()].view(o.shape).copy_(o, non_blocking=True)
                off += o.numel()
            done[s].record(st)

    def run(count):
        for i in range(count):
            if i >= depth:
                done[i % depth].synchronize()
            submit(i)
        for s in range(depth):
            done[s].synchronize()
    run(4)
    runner.barrier()
    t0 = time.perf_counter()
    run(steps)
    torch.cuda.synchronize()
    dt = runner.max_over_ranks(time.perf_counter() - t0)
    e2e_pcm = {"value": runner.world * B * steps / dt, "unit": UNIT, "h2d_bytes_per_step": B * n * 2 + sum(v.nbytes for k, v in runner.host[0].items() if k != "x_data"),
               "d2h_bytes_per_step": int(sum(pins[0].shape)) * 4,
               "api": "pinned int16 PCM -> H2D -> sar_fbank_pcm16_fwd -> engine.forward_slot (3 slots) -> D2H, wall clock"}
    return roof, e2e_pcm


# ====================================================================================== training slice
def train_slice_bench(dev, rank, world, B=64, steps=10):
    """one optimisation step of everything above the frozen ResNet on the accent path --
    training.HeadTrainer(train_crnn=True).train_on_batch: ResNet inference forward, then CNN_LIN / CNN_LIN_LN / CRNN Bi-GRU
    (BPTT) / CRNN_LN / AR_DS / AR_DS_LN / GhostVLAD / embedding / classifier / ArcFace forward + backward in training mode,
    ONE flat gradient all-reduce over the ranks, Keras Adam.  Weak scaling."""
    import torch.distributed as tdist
    from aesrc2020_b200 import model as mdl, training as T, utils as us
    with contextlib.redirect_stdout(io.StringIO()):
        model, _ = mdl.SAR_Net((500, 80, 1), **dict(CONFIGS["cfg2"]["kw"]))
    x, y = us.synthetic_batch(model.config, B, seed=100 + rank)
    tr = T.HeadTrainer(model, lr=0.01, train_crnn=True)
    xd = {k: model._to_device(k, v) for k, v in x.items()}
    for _ in range(4):                                # two eager steps, the graph capture, one replay
        tr.train_on_batch(xd, y)
    torch.cuda.synchronize()
    if world > 1:
        tdist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        last = tr.train_on_batch(xd, y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        ms = float(t.item())
    return {"workload": "everything above the frozen ResNet on the accent path (CNN_LIN, CRNN Bi-GRU with BPTT, AR_DS .. y_accent / y_disc), B=%d/GPU, fwd + bwd + Adam" % B,
            "ms_per_step": ms, "value": world * B / (ms * 1e-3), "unit": "utt/s (training, partial graph)",
            "trainable_parameters": int(sum(tr.p[k].numel() for k in tr.keys)), "allreduce_bytes_per_step": 4 * int(sum(tr.p[k].numel() for k in tr.keys)) if world > 1 else 0,
            "loss": float(last["loss"]), "scaling": "weak",
            "note": "the step above the frozen ResNet replays as one CUDA graph on a single GPU (eager under torchrun: the gradient all-reduce); the Bi-GRU is one GEMM + one gate kernel per time step and direction, fp32 CUDA cores; includes the frozen ResNet's inference forward"}


def train_full_bench(dev, rank, world, B=64, steps=2):
    """SURVEY 8f-1, complete but correctness-first: one optimisation step of the WHOLE model (configs[4] graph: ResNet in training
    mode + CRNN + CTC branch + GhostVLAD + Circle-Loss), training.HeadTrainer(train_resnet=True, train_ctc=True).  The ResNet's
    backward is fp32 CUDA-core code (training_resnet.py), so this number says what the slice costs today, not what B200 can do."""
    import torch.distributed as tdist
    from aesrc2020_b200 import model as mdl, training as T, utils as us
    with contextlib.redirect_stdout(io.StringIO()):
        model, _ = mdl.SAR_Net((500, 80, 1), **dict(CONFIGS["cfg5"]["kw"]))
    x, y = us.synthetic_batch(model.config, B, seed=300 + rank)
    tr = T.HeadTrainer(model, lr=0.005, train_resnet=True, train_ctc=True)
    xd = {k: model._to_device(k, v) for k, v in x.items()}
    first = tr.train_on_batch(xd, y)
    for _ in range(3):                                # warm-up: two eager steps (the allocator still grows), the third is captured as a
        tr.train_on_batch(xd, y)                      # CUDA graph (training.HeadTrainer.step_graphed), the fourth is the first pure replay
    torch.cuda.synchronize()
    if world > 1:
        tdist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        last = tr.train_on_batch(xd, y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        ms = float(t.item())
    npar = int(sum(tr.p[k].numel() for k in tr.keys))
    return {"workload": "whole model, nothing frozen (configs[4] graph: thin-ResNet34 in training mode + CRNN + CTC branch + GhostVLAD + Circle-Loss), "
                        "B=%d/GPU x 500 frames, fwd + bwd + Adam" % B,
            "ms_per_step": ms, "value": world * B / (ms * 1e-3), "unit": "utt/s (training, whole graph)", "trainable_parameters": npar,
            "allreduce_bytes_per_step": 4 * npar if world > 1 else 0, "loss_first": float(first["loss"]), "loss_last": float(last["loss"]),
            "scaling": "weak",
            "note": "correctness-first: the ResNet's training forward / backward are fp32 CUDA-core kernels (row-parallel BN statistics, "
                    "smem-tiled weight gradients, data gradients through the implicit-GEMM forward kernel); one CUDA-graph replay per step on a single GPU, eager under torchrun; parity with the "
                    "float64 autograd oracle is tested, tensor-core speed is not the claim"}


# ====================================================================================== strong scaling
def strong_scaling(dev, rank, world, peaks, global_b=4096, micro=512, reps=2):
    """configs[4]: global B=4096, split contiguously over the ranks (dist.shard_slice), every rank runs its share as
    micro-batches of `micro` utterances through the slot pipeline; time = max over ranks, value = 4096 / time."""
    from aesrc2020_b200 import dist as sdist
    import torch.distributed as tdist
    sl = sdist.shard_slice(global_b, rank, world)
    mine = sl.stop - sl.start
    nmicro = (mine + micro - 1) // micro
    mb = min(micro, mine)
    cfgd = dict(CONFIGS["cfg5"]); cfgd["B"] = mb
    r = Runner("cfg5", cfgd, dev, rank, world, rotate=min(4, max(2, nmicro)))
    r.warm(3)
    times = []
    for rep in range(reps):
        r.barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        r.fork_slots()
        for i in range(nmicro):
            r.step(i)
        r.join_slots()
        t1.record()
        r.barrier()
        times.append(r.max_over_ranks(t0.elapsed_time(t1)))
    r.close()
    ms = float(np.min(times))
    return {"workload": "configs[4] strong scaling: global B=%d over %d GPU(s), %d micro-batch(es) of %d per GPU" % (global_b, world, nmicro, mb),
            "global_batch": global_b, "n_gpus": world, "ms_total": ms, "value": global_b / (ms * 1e-3), "unit": UNIT,
            "scaling": "strong"}


def run_reference(args, cfgd):
    """Reference arm: the oracle's torch-CPU fp32 restatement of the Keras forward, all host threads."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from aesrc2020_b200.config import SARConfig
    from aesrc2020_b200 import weights as W, utils as us
    from oracle import sarnet_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = SARConfig(input_shape=(cfgd["T"], 80, 1), **cfgd["kw"])
    w = W.init_weights(cfg, 1234)
    Bs = min(args.ref_batch, cfgd["B"])
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    # the same rotation over distinct synthetic batches as the GPU arm (seeds as in StepRunner); each timed step is a
    # Bs-utterance sample of one of them
    nrot = max(1, args.rotate)
    xs = [us.synthetic_batch(cfg, Bs, seed=2020 + i)[0] for i in range(nrot)]
    in_bytes = sum(np.asarray(v).nbytes for v in us.synthetic_batch(cfg, 1, seed=1)[0].values()) * cfgd["B"]
    fwd = lambda i: O.sar_net_forward(w, xs[i % nrot], **cfg.model_kwargs(), dtype=torch.float32)
    for i in range(max(1, min(args.warmup, 2))):
        fwd(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        fwd(i)
    dt = (time.perf_counter() - t0) / args.steps
    v = Bs / dt
    sample = ("%d of the %d utterances of one step per timed step, throughput normalised per utterance (torch-CPU fp32 "
              "restatement of the Keras forward = oracle port, %d threads)" % (Bs, cfgd["B"], cores))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(cfgd, cfgd["B"], cfgd["T"], cfg.plan(), in_bytes, args.rotate, world),
        "note": "reference arm = torch-CPU restatement of the Keras forward (oracle port, kind 'port'); the literal "
                "Keras/TF graph is not runnable in this image; each timed step is a %d-utterance sample of the "
                "%d-utterance batch, throughput normalised per utterance" % (Bs, cfgd["B"]),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=list(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch")
    ap.add_argument("--ref-batch", type=int, default=32, help="utterances per step of the CPU reference sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager", action="store_true", help="no CUDA-graph replay (for ncu launch lists)")
    ap.add_argument("--pipeline", type=int, default=4, help="independent steps in flight on separate streams (1 = one stream)")
    ap.add_argument("--rotate", type=int, default=16, help="distinct input batches rotated through (L2 hygiene)")
    ap.add_argument("--quick", action="store_true", help="primary workload only: no extra_configs / strong / sustained / fbank / vlad sweeps")
    ap.add_argument("--sustain-s", type=float, default=2.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    cfgd = dict(CONFIGS[args.config])
    if args.batch:
        cfgd["B"] = args.batch
    if args.impl == "reference":
        return run_reference(args, cfgd)

    from aesrc2020_b200 import dist as sdist
    import torch.distributed as tdist
    env = sdist.init_from_env()
    rank, world, local = env["rank"], env["world_size"], env["local_rank"]
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    dev = torch.device("cuda", local)
    peaks = load_peaks()
    quick = args.quick or args.eager
    t_wall0 = time.perf_counter()

    # ---------------------------------------------------------------- primary workload
    R = Runner(args.config, cfgd, dev, rank, world, args.rotate, eager=args.eager, pipeline=args.pipeline)
    B, T, plan = R.B, R.T, R.plan
    launches_per_step = R.count_launches()
    R.warm(args.warmup)
    single_ms = R.time_single(args.steps, args.warmup) if R.depth > 1 else None
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_per_step = R.time_value(args.steps, args.warmup)            # timed region A: the `value`
    conv_ms, vlad_ms, step_b_ms = R.time_segments(args.steps, args.warmup)
    sustained = None
    if not quick:
        n_sus = int(max(args.steps, np.ceil(args.sustain_s * 1e3 / ms_per_step)))
        sus_ms = R.time_value(n_sus, args.warmup)
        sustained = {"value": world * B / (sus_ms * 1e-3), "unit": UNIT, "steps": n_sus, "seconds": n_sus * sus_ms * 1e-3,
                     "ms_per_step": sus_ms}
    clocks = sampler.stop() if rank == 0 else None
    value = world * B / (ms_per_step * 1e-3)
    e2e = R.time_e2e(args.steps)
    roof = conv_roofline(plan, B, conv_ms, step_b_ms, peaks)
    roofline_fbank = e2e_pcm = None
    if not quick:
        roofline_fbank, e2e_pcm = fbank_bench(R, args.steps, peaks)

    # ---- CPU baseline (oracle port, bounded sample) on rank 0, N=1 only
    cpu = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        from oracle import sarnet_oracle as O
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        Bs = min(args.ref_batch, B)
        xs = {k: v[:Bs] for k, v in R.host[0].items()}
        fwd = lambda: O.sar_net_forward(R.model.weights, xs, **R.cfg.model_kwargs(), dtype=torch.float32)
        fwd()
        t0 = time.perf_counter()
        reps = 0
        while reps < 3 or (time.perf_counter() - t0 < 10.0 and reps < 50):
            fwd()
            reps += 1
        dt = (time.perf_counter() - t0) / reps
        cpu = {"value": Bs / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d utterances of step 0 x %d repeats, torch-CPU fp32 restatement of the Keras forward" % (Bs, reps)}
    in_bytes = R.in_bytes
    R.close()

    # ---------------------------------------------------------------- the other north-star configurations
    extra, strong, roofline_vlad, train_slice, train_full = None, None, None, None, None
    if not quick:
        extra = {}
        vlad_in_graph = {64: vlad_ms} if (vlad_ms and args.config == "cfg2" and B == 64) else {}
        for name in ("cfg3", "cfg4", "cfg5"):
            if name == args.config:
                continue
            try:
                r = Runner(name, dict(CONFIGS[name]), dev, rank, world, rotate=4, pipeline=args.pipeline)
                extra[name] = r.measure(max(5, args.steps // 2), 3, peaks)
                if name == "cfg5" and extra[name].get("vlad_ms_in_graph"):
                    vlad_in_graph[512] = extra[name]["vlad_ms_in_graph"]
                r.close()
            except Exception as ex:                                        # keep the primary line alive
                extra[name] = {"error": "%s: %s" % (type(ex).__name__, ex)}
                torch.cuda.empty_cache()
        try:
            strong = strong_scaling(dev, rank, world, peaks)
        except Exception as ex:
            strong = {"error": "%s: %s" % (type(ex).__name__, ex)}
        try:
            train_slice = train_slice_bench(dev, rank, world)
        except Exception as ex:
            train_slice = {"error": "%s: %s" % (type(ex).__name__, ex)}
        try:
            train_full = train_full_bench(dev, rank, world)
        except Exception as ex:
            train_full = {"error": "%s: %s" % (type(ex).__name__, ex)}
            torch.cuda.empty_cache()
        if rank == 0:
            try:
                roofline_vlad = vlad_roofline(dev, peaks, vlad_ms_in_graph=vlad_in_graph)
            except Exception as ex:
                roofline_vlad = {"error": "%s: %s" % (type(ex).__name__, ex)}
        if world > 1:
            tdist.barrier()

    if rank != 0:
        if world > 1:
            tdist.barrier()
            tdist.destroy_process_group()
        return

    # ---- roofline details of the residual-block convolution kernels
    traffic, traffic_src = None, None
    try:                                                # DRAM bytes per launch from the committed ncu pass of THIS build
        with open(os.path.join(ROOT, "profiles", "r2_conv_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("build_id") == conv_build_id() and args.config == "cfg2" and B == tj.get("B"):
            traffic, traffic_src = tj["traffic_bytes_per_launch"], tj.get("source")
    except Exception:
        pass
    roof.update({"traffic": traffic,
                 "traffic_unit": "bytes/launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu; null unless profiles/r2_conv_traffic.json was taken on this build of the conv sources: %s)" % (traffic_src or conv_build_id()),
                 "kernel": "residual-block conv (%d launches/step for thin-ResNet34)" % roof["launches_per_step"],
                 "timing": "external CUDA events (graph event-record nodes) around the block-conv launches of the single-stream "
                           "step graph, read after each of K replays run right after the value region",
                 "peak_source": peaks["source"] + ", sustained bf16 for a kernel timed inside a long step",
                 "build_id": build_id(), "conv_build_id": conv_build_id()})
    launches = launches_per_step * args.steps

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16x2 (fp16 hi+lo split operands, fp32 accumulate; fp32 elsewhere)", "data": "synthetic",
        "config": workload_config(cfgd, B, T, plan, in_bytes, args.rotate, world),
        "pipeline": "%d independent step(s) in flight (one stream + CUDA graph + buffer set each)" % (1 if args.eager else args.pipeline),
        "e2e": e2e, "gpu_launches": launches, "gpu_launches_per_step": launches_per_step, "roofline": roof, "clocks": clocks,
        "wall_s": None,
    }
    if single_ms is not None:
        line["single_stream"] = {"ms_per_step": single_ms, "value": world * B / (single_ms * 1e-3), "unit": UNIT,
                                 "note": "the same steps on ONE stream, one after the other (per-batch latency)"}
    if sustained:
        line["sustained"] = sustained
    if roofline_vlad:
        line["roofline_vlad"] = roofline_vlad
    if train_slice:
        line["train_slice"] = train_slice
        line["train_full"] = train_full
    if roofline_fbank:
        line["roofline_fbank"] = roofline_fbank
    if e2e_pcm:
        line["e2e_pcm"] = e2e_pcm
    if extra:
        line["extra_configs"] = extra
    if strong:
        line["strong"] = strong
    if cpu:
        line["cpu_baseline"] = cpu
    line["wall_s"] = time.perf_counter() - t_wall0
    print(json.dumps(line))
    if world > 1:
        tdist.barrier()
        tdist.destroy_process_group()


if __name__ == "__main__":
    main()
