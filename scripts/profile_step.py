"""In-situ per-op device times of one eager forward step (warm L2, real data flow): every C-ABI wrapper in
ops.py / tc.py is bracketed by CUDA events.  Complements the ncu launch list (whose per-kernel times are
cold-cache and serialised).  Usage: python scripts/profile_step.py [--config cfg2] [--batch 64] [--steps 10]
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

records = []          # (name, start, end)
cur_step = [0]


def wrap(mod, name, label=None):
    fn = getattr(mod, name)

    def inner(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*a, **k)
        e1.record()
        lab = label or name
        if name == "conv2d":
            lab = "dense/conv_ffma %s" % (tuple(a[1].shape),)
        if name == "conv_tc":
            pa = a[0]
            lab = "conv_tc %dx%d cin%d cout%d%s" % (k["out_hw"][0], k["out_hw"][1], pa.C, k["cout"],
                                                    " s2" if pa.split else "")
        records.append((cur_step[0], lab, e0, e1))
        return r
    setattr(mod, name, inner)


for n in ("conv2d", "layernorm", "bigru", "vlad", "avgpool", "gemm_splitk", "head", "ctc", "loss_reduce", "maxpool2d",
          "affine_relu"):
    wrap(ops, n)
for n in ("conv_tc", "stem_pool", "maxpool_planes"):
    wrap(tc, n)

for i in range(3):
    eng.forward(batches[i % 8])
torch.cuda.synchronize()
del records[:]
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for i in range(args.steps):
    cur_step[0] = i
    eng.forward(batches[(3 + i) % 8])
t1.record()
torch.cuda.synchronize()
agg = collections.OrderedDict()
for st, lab, e0, e1 in records:
    agg.setdefault(lab, []).append(e0.elapsed_time(e1) * 1e3)
tot = 0.0
rows = []
for lab, v in agg.items():
    per_step = sum(v) / args.steps
    rows.append((lab, len(v) // args.steps, per_step))
    tot += per_step
print("eager step %.1f us (event sum %.1f us), B=%d" % (t0.elapsed_time(t1) * 1e3 / args.steps, tot, B))
for lab, n, us_ in rows:
    print("%-44s x%-2d %8.1f us %5.1f%%" % (lab, n, us_, 100 * us_ / tot))
if args.json:
    json.dump({"B": B, "config": args.config, "rows": rows, "event_sum_us": tot}, open(args.json, "w"), indent=1)
