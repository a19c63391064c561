"""Per-op device times INSIDE the replayed CUDA graph of one forward step: every C-ABI wrapper in ops.py / tc.py
is bracketed by external CUDA events (cudaEventRecordExternal nodes) while the step is captured, the graph is
replayed `--steps` times and the event pairs are read after each replay.  Unlike scripts/profile_step.py (eager,
host-launch-rate bound at B=64) these are the times the bench's graph replay actually sees.
Usage: python scripts/profile_graph.py [--config cfg2] [--batch 64] [--steps 10] [--json out.json]
"""
import argparse
import collections
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from aesrc2020_b200 import model as mdl, ops, tc, utils as us

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cfg2")
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--json", default="")
args = ap.parse_args()
cfgd = dict(bench.CONFIGS[args.config])
B = args.batch or cfgd["B"]
model, _ = mdl.SAR_Net((cfgd["T"], 80, 1), **cfgd["kw"])
eng = model.engine()
batches = []
for i in range(8):
    x, _ = us.synthetic_batch(model.config, B, seed=2020 + i)
    batches.append({k: model._to_device(k, v).clone() for k, v in x.items()})

records = []          # (label, e0, e1) in capture order
capturing = [False]


def wrap(mod, name):
    fn = getattr(mod, name)

    def inner(*a, **k):
        if not capturing[0]:
            return fn(*a, **k)
        e0 = torch.cuda.Event(enable_timing=True, external=True)
        e1 = torch.cuda.Event(enable_timing=True, external=True)
        e0.record()
        r = fn(*a, **k)
        e1.record()
        lab = name
        if name == "conv2d":
            lab = "dense/conv_ffma %s" % (tuple(a[1].shape),)
        if name == "conv_tc":
            pa = a[0]
            lab = "conv_tc %dx%d cin%d cout%d%s" % (k["out_hw"][0], k["out_hw"][1], pa.C, k["cout"],
                                                    " s2" if pa.split else "")
        records.append((lab, e0, e1))
        return r
    setattr(mod, name, inner)


for n in [x for x in dir(ops) if callable(getattr(ops, x)) and not x.startswith("_") and
          getattr(getattr(ops, x), "__module__", "") == ops.__name__ and x not in ("dense",)]:
    wrap(ops, n)
for n in [x for x in dir(tc) if callable(getattr(tc, x)) and not x.startswith("_") and
          getattr(getattr(tc, x), "__module__", "") == tc.__name__ and
          x not in ("alloc_planes", "plane_rows", "pack_weights", "tap_table", "Planes", "sar_tc_conv")]:
    wrap(tc, n)

static_in = {k: v.clone() for k, v in batches[0].items()}
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3):
        eng.forward(static_in)
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
graph = torch.cuda.CUDAGraph()
capturing[0] = True
g0 = torch.cuda.Event(enable_timing=True, external=True)
g1 = torch.cuda.Event(enable_timing=True, external=True)
with torch.cuda.graph(graph):
    g0.record()
    eng.forward(static_in)
    g1.record()
capturing[0] = False

agg = collections.OrderedDict()
tot_graph = 0.0
for i in range(args.steps + 3):
    for k, v in batches[i % 8].items():
        static_in[k].copy_(v, non_blocking=True)
    graph.replay()
    torch.cuda.synchronize()
    if i < 3:
        continue
    tot_graph += g0.elapsed_time(g1) * 1e3
    for lab, e0, e1 in records:
        agg.setdefault(lab, []).append(e0.elapsed_time(e1) * 1e3)
tot = 0.0
rows = []
for lab, v in agg.items():
    per_step = sum(v) / args.steps
    rows.append((lab, len(v) // args.steps, per_step))
    tot += per_step
print("graph step %.1f us with event nodes (op sum %.1f us), B=%d" % (tot_graph / args.steps, tot, B))
for lab, n, us_ in rows:
    print("%-44s x%-2d %8.1f us %5.1f%%  (%.1f us each)" % (lab, n, us_, 100 * us_ / tot, us_ / n))
if args.json:
    json.dump({"B": B, "config": args.config, "rows": rows, "op_sum_us": tot, "graph_us": tot_graph / args.steps},
              open(args.json, "w"), indent=1)
