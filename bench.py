#!/usr/bin/env python
"""bench.py -- utterances/sec of the SAR-Net forward path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One "step" = one forward of the hot path over one per-GPU batch of synthetic fbank input
(BASELINE.json configs[1] by default: B=64 x 500 frames x 80 mel, thin ResNet-34 + Bi-GRU +
GhostVLAD(64c/8g) + ArcFace).  Weak scaling: every rank processes its own B utterances; the
step ends with the path's single collective, an all-reduce(SUM) of the 8-float loss vector.

Printed JSON (rank 0, one line):
  value      whole-job utterances/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e        same metric through model.predict() with HOST numpy inputs: pinned H2D copy of the
             step's inputs and D2H of the outputs inside the timed region
  roofline   the ResNet residual-block convolution kernel (the dominant kernel): algorithmic
             FLOPs / bytes per step over its measured device time inside the timed region
  cpu_baseline  the oracle's torch-CPU fp32 restatement of the Keras forward on the host cores
`--impl reference` times that CPU restatement as the reference arm (the literal Keras/TF
graph cannot run here: no tensorflow/keras in the image and CuDNNGRU has no CPU kernel).
"""
from __future__ import annotations

import argparse
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

CONFIGS = {
    # BASELINE.json configs[1]
    "cfg2": dict(B=64, T=500, kw=dict(ctc_enable=False, ar_enable=True, disc_enable=True, res_type="res34",
                                      res_filters=32, mto="gvlad", vlad_clusters=64, ghost_clusters=8,
                                      metric_loss="arcface", margin=0.3),
                 workload="configs[1]: B=64/GPU x 500 frames x 80 mel, thin-ResNet34+BiGRU+GhostVLAD(64c/8g)+ArcFace fwd"),
    # BASELINE.json configs[4] per-GPU shard (512 utt/GPU), for manual sweeps
    "cfg5": dict(B=512, T=500, kw=dict(ctc_enable=True, ar_enable=True, disc_enable=True, res_type="res34",
                                       res_filters=32, mto="gvlad", vlad_clusters=64, ghost_clusters=8,
                                       metric_loss="circleloss", margin=0.2),
                 workload="configs[4] shard: B=512/GPU x 500 frames, CRNN+GhostVLAD+Circle-Loss+CTC fwd"),
    "cfg3": dict(B=256, T=800, kw=dict(ctc_enable=True, ar_enable=True, disc_enable=True, res_type="res34",
                                       res_filters=32, mto="bigru", metric_loss="circleloss", margin=0.2),
                 workload="configs[2]: B=256 x 200-800 frames padded to 800, CTC+Circle-Loss fwd"),
}


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "tflops": float(p["bf16_tflops_sustained"]),
                "tflops_burst": float(p["bf16_tflops"]), "source": "measured (MEASURED_PEAKS.json)"}
    except Exception:
        return {"hbm_gbs": 6650.0, "tflops": 1400.0, "tflops_burst": 1590.0, "source": "fallback (B200_PROFILING.md)"}


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
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


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


def run_reference(args, cfgd):
    """Reference arm: the oracle's torch-CPU fp32 restatement of the Keras forward, all host threads."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
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
    x, _ = us.synthetic_batch(cfg, Bs, seed=2020)
    fwd = lambda: O.sar_net_forward(w, x, **cfg.model_kwargs(), dtype=torch.float32)
    for _ in range(args.warmup):
        fwd()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fwd()
    dt = (time.perf_counter() - t0) / args.steps
    v = Bs / dt
    sample = "%d of the %d utterances of one step per timed step (torch-CPU fp32, %d threads)" % (Bs, cfgd["B"], cores)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfgd["workload"], "note": "reference arm = torch-CPU restatement of the Keras forward "
                   "(oracle port); the literal Keras/TF graph is not runnable in this image"},
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
    ap.add_argument("--lanes", type=int, default=0, help="concurrent micro-batch lanes per step (0 = the model's default)")
    ap.add_argument("--pipeline", type=int, default=4, help="independent steps in flight on separate streams (1 = one stream)")
    ap.add_argument("--rotate", type=int, default=16, help="distinct input batches rotated through (L2 hygiene)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    cfgd = dict(CONFIGS[args.config])
    if args.batch:
        cfgd["B"] = args.batch
    if args.impl == "reference":
        return run_reference(args, cfgd)

    from aesrc2020_b200 import dist as sdist, model as mdl, utils as us, ops
    from aesrc2020_b200.engine import StepOpts, SLOT_OPTS
    import torch.distributed as tdist
    env = sdist.init_from_env()
    rank, world, local = env["rank"], env["world_size"], env["local_rank"]
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    dev = torch.device("cuda", local)
    B, T = cfgd["B"], cfgd["T"]
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        model, _ = mdl.SAR_Net((T, 80, 1), device=dev, **cfgd["kw"])
    cfg = model.config
    eng = model.engine()
    plan = cfg.plan()

    # ---- synthetic inputs: `rotate` distinct batches, resident in HBM (value) and in pinned host memory (e2e)
    host_batches, dev_batches = [], []
    lengths = None
    for i in range(args.rotate):
        if args.config == "cfg3":
            lengths = np.random.RandomState(100 + i + 1000 * rank).randint(200, 801, size=B)
        x, _ = us.synthetic_batch(cfg, B, seed=2020 + i + 1000 * rank, lengths=lengths)
        host_batches.append(us.pinned_like(x))      # e2e: inputs start in pinned host memory (a loader's ring buffer)
        dev_batches.append({k: model._to_device(k, v).clone() for k, v in x.items()})
    torch.cuda.synchronize()
    in_bytes = sum(v.nbytes for v in host_batches[0].values())

    if args.eager:
        model.use_graph = False      # profiling aid (ncu launch lists): no graph capture anywhere, e2e included
    if args.lanes:
        model.lanes = args.lanes
    lanes = 1 if args.eager else model._lanes_for(B)

    depth = 1 if (args.eager or lanes > 1) else max(1, args.pipeline)
    main_stream = torch.cuda.current_stream()

    def step_device(i, graphed=True):
        x = dev_batches[i % args.rotate]
        if graphed and depth > 1:                   # consecutive batches alternate between `depth` streams / graphs
            out, st = eng.forward_slot(x, i % depth)
            if world > 1:
                with torch.cuda.stream(st):
                    tdist.all_reduce(out["loss_vector"])
            return out
        out = eng.forward_lanes(x, lanes) if (graphed and not args.eager) else eng.forward(x)
        vec = out["loss_vector"]
        if world > 1:
            tdist.all_reduce(vec)
        return out

    def join_slots():
        for st in eng._lane_streams[:depth] if depth > 1 else []:
            main_stream.wait_stream(st)

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (captures the CUDA graph of the step on first use)
    l0 = ops.LAUNCHES["n"]
    step_device(0, graphed=False)
    launches_per_step = ops.LAUNCHES["n"] - l0
    if lanes > 1:       # every lane launches the whole kernel sequence on its rows (no chain launches), + 1 batch loss reduce
        l0 = ops.LAUNCHES["n"]
        eng.forward({k: v[:B // lanes] for k, v in dev_batches[0].items()}, opts=StepOpts(lane=1, no_chain=True))
        launches_per_step = (ops.LAUNCHES["n"] - l0 - 1) * lanes + 1
    for i in range(depth if depth > 1 else 0):      # capture every slot's CUDA graph before the W warm-up steps
        step_device(i)
    for i in range(args.warmup):
        step_device(i)
    barrier()

    # ---- single-stream reference: the same K steps strictly one after the other (per-step latency, reported beside
    # the pipelined throughput)
    single_ms = None
    if depth > 1:
        for i in range(3):
            eng.forward_lanes(dev_batches[i % args.rotate], 1)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for i in range(args.steps):
            out1 = eng.forward_lanes(dev_batches[(args.warmup + i) % args.rotate], 1)
            if world > 1:
                tdist.all_reduce(out1["loss_vector"])
        s1.record()
        barrier()
        single_ms = s0.elapsed_time(s1) / args.steps

    # ---- timed region A (the `value`): K graph replays, device-resident inputs
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for st in (eng._lane_streams[:depth] if depth > 1 else []):
        st.wait_stream(main_stream)                 # no slot starts before the start event
    for i in range(args.steps):
        step_device(args.warmup + i)
    join_slots()
    t_end.record()
    barrier()
    launches = launches_per_step * args.steps       # kernels replayed from the graph
    if depth > 1:                                   # pipeline slots launch every layer separately (no stage chains)
        l0 = ops.LAUNCHES["n"]
        eng.forward(dev_batches[0], opts=SLOT_OPTS(1))
        per_other = ops.LAUNCHES["n"] - l0
        torch.cuda.synchronize()
        launches = per_other * args.steps
    ms = t_start.elapsed_time(t_end)

    # ---- timed region B (roofline): K more replays of the SAME step captured a second time with two external
    # CUDA events (event-record graph nodes) bracketing the 36 block-conv launches; read after every replay
    conv_ev = []
    eng.resnet.record_events = conv_ev
    step_b = lambda i: eng.forward_graphed(dev_batches[i % args.rotate], tag="conv-events")
    if args.eager:
        step_b = lambda i: eng.forward(dev_batches[i % args.rotate])
    step_b(0)                                   # capture (records the event pair once) + first replay
    torch.cuda.synchronize()
    conv_times, step_times = [], []
    tb0, tb1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(args.steps):
        tb0.record()
        step_b(args.warmup + i)
        tb1.record()
        torch.cuda.synchronize()
        a, b = conv_ev[-1]      # graph: the captured pair (re-recorded by every replay); eager: this step's pair
        conv_times.append(a.elapsed_time(b))
        step_times.append(tb0.elapsed_time(tb1))
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    eng.resnet.record_events = None
    eager_ms_per_step = float(np.mean(step_times))
    conv_ms = float(np.mean(conv_times))
    if world > 1:
        tt = torch.tensor([ms], device=dev)
        tdist.all_reduce(tt, op=tdist.ReduceOp.MAX)
        ms = float(tt.item())
    ms_per_step = ms / args.steps
    value = world * B / (ms_per_step * 1e-3)

    # ---- e2e: the public API with HOST inputs; every step's pinned H2D and the D2H of its outputs are inside the
    # timed region.  (1) model.predict_generator over the K batches (Keras' queued generator loop: copies of step
    # i+1 / i-1 run under step i's kernels) -- the `e2e` value; (2) K blocking model.predict() calls, one full
    # H2D -> kernels -> D2H -> host-sync round trip each -- reported beside it as `e2e.blocking_predict`.
    def gen_host(first, count):
        for i in range(count):
            yield host_batches[(first + i) % args.rotate]
    outs = model.predict_generator(gen_host(0, 3))
    out_bytes = sum(o.nbytes for o in (outs if isinstance(outs, list) else [outs])) // 3
    barrier()
    t0 = time.perf_counter()
    model.predict_generator(gen_host(3, args.steps))
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    for i in range(3):
        model.predict(host_batches[i % args.rotate], batch_size=B)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        model.predict(host_batches[(3 + i) % args.rotate], batch_size=B)
    torch.cuda.synchronize()
    sync_s = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([e2e_s, sync_s], device=dev)
        tdist.all_reduce(tt, op=tdist.ReduceOp.MAX)
        e2e_s, sync_s = float(tt[0].item()), float(tt[1].item())
    e2e_value = world * B * args.steps / e2e_s
    e2e_sync_value = world * B * args.steps / sync_s

    if rank != 0:
        if world > 1:
            tdist.barrier()
            tdist.destroy_process_group()
        return

    # ---- roofline of the residual-block convolution kernel
    peaks = load_peaks()
    act_bytes = 4.0            # activations as stored by this build: fp16 hi + fp16 lo planes
    F, Bt, nconv = conv_algorithmic_work(plan, B, act_bytes)
    t_conv = conv_ms * 1e-3
    tensor_time = F / (peaks["tflops"] * 1e12)
    hbm_time = Bt / (peaks["hbm_gbs"] * 1e9)
    if hbm_time >= tensor_time:
        roof = {"bound": "hbm", "achieved": Bt / t_conv / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s"}
    else:
        roof = {"bound": "tensor", "achieved": F / t_conv / 1e12, "peak": peaks["tflops"], "unit": "TFLOP/s"}
    roof["frac"] = roof["achieved"] / roof["peak"]
    # DRAM bytes per launch from the committed ncu pass of this workload (profiles/r1_conv_traffic_v6.json); ncu is
    # never run inside the bench.  Only valid for the default workload it was captured on.
    traffic = None
    try:
        if args.config == "cfg2" and B == 64:
            with open(os.path.join(ROOT, "profiles", "r1_conv_traffic_v6.json")) as f:
                traffic = json.load(f)["traffic_bytes_per_launch"]
    except Exception:
        traffic = None
    roof.update({"traffic": traffic, "traffic_unit": "bytes/launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu, profiles/r1_conv_traffic_v6.json)",
                 "algorithmic_bytes_per_launch": Bt / nconv,
                 "kernel": "residual-block conv (36 launches/step for thin-ResNet34)",
                 "launches_per_step": nconv, "avg_launch_us": conv_ms * 1e3 / nconv, "conv_ms_per_step": conv_ms,
                 "share_of_step": conv_ms / eager_ms_per_step, "ms_per_step_region_b": eager_ms_per_step,
                 "timing": "external CUDA events (graph event-record nodes) around the 36 block-conv launches, read after each of K graph replays run right after the value region", "algorithmic_gflop_per_step": F / 1e9,
                 "algorithmic_mb_per_step": Bt / 1e6, "tflops_achieved": F / t_conv / 1e12,
                 "peak_source": peaks["source"] + ", sustained bf16 for a kernel timed inside a long step"})

    # ---- CPU baseline (oracle port, bounded sample) on rank 0, N=1 only
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import sarnet_oracle as O
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        Bs = min(args.ref_batch, B)
        xs = {k: v[:Bs] for k, v in host_batches[0].items()}
        fwd = lambda: O.sar_net_forward(model.weights, xs, **cfg.model_kwargs(), dtype=torch.float32)
        fwd()
        t0 = time.perf_counter()
        reps = 0
        while reps < 3 or (time.perf_counter() - t0 < 10.0 and reps < 50):
            fwd()
            reps += 1
        dt = (time.perf_counter() - t0) / reps
        cpu = {"value": Bs / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d utterances of step 0 x %d repeats, torch-CPU fp32 restatement of the Keras forward" % (Bs, reps)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16x2 (fp16 hi+lo split operands, fp32 accumulate; fp32 elsewhere)", "data": "synthetic",
        "config": {"workload": cfgd["workload"], "per_gpu_batch": B, "frames": T, "seq_len": plan.seq_len,
                   "l2": "inputs rotate over %d distinct batches (%.0f MB > 126 MB L2); activations %.0f MB/step"
                         % (args.rotate, args.rotate * in_bytes / 1e6, 2 * Bt / 1e6 / 3),
                   "parallelism": "dp%d (batch-sharded, 1 all-reduce of 8 floats/step)" % world,
                   "lanes": "%d concurrent micro-batch graph(s) per step on separate streams" % lanes,
                   "pipeline": "%d independent step(s) in flight (one stream + CUDA graph + buffer set each)" % depth},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes,
                "api": "model.predict_generator(batches in pinned host memory, steps=K): %d batches in flight (copy stream + one compute stream/graph per slot), wall clock over K steps" % model.PIPE_DEPTH,
                "blocking_predict": e2e_sync_value},
        "gpu_launches": launches, "roofline": roof, "clocks": clocks,
    }
    if single_ms is not None:
        line["single_stream"] = {"ms_per_step": single_ms, "value": world * B / (single_ms * 1e-3), "unit": UNIT,
                                 "note": "the same steps on ONE stream, one after the other (per-batch latency)"}
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))
    if world > 1:
        tdist.barrier()
        tdist.destroy_process_group()


if __name__ == "__main__":
    main()
