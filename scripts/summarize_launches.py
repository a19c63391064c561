"""Summarise an ncu launch list (--metrics gpu__time_duration.sum[,dram__bytes_*] --csv): per kernel count / total us / share,
optionally only launches [first, last).  `--traffic-json OUT --build-id ID --B n`: also write the DRAM bytes per
residual-block conv launch (36 per thin-ResNet34 step) for bench.py's roofline.traffic.
Usage: python scripts/summarize_launches.py <csv> [--first N] [--last M] [--steps K]"""
import argparse, csv, json, re, sys
ap = argparse.ArgumentParser()
ap.add_argument("csv"); ap.add_argument("--first", type=int, default=0); ap.add_argument("--last", type=int, default=10**9)
ap.add_argument("--steps", type=float, default=1.0); ap.add_argument("--align", action="store_true"); ap.add_argument("--only-steps-with"); ap.add_argument("--traffic-json"); ap.add_argument("--build-id"); ap.add_argument("--B", type=int, default=64)
a = ap.parse_args()
rows = list(csv.reader(open(a.csv)))
hdr = next(r for r in rows if r and r[0] == "ID")
ix = {h: i for i, h in enumerate(hdr)}
launch = {}
for r in rows:
    if len(r) != len(hdr) or not r[0].isdigit():
        continue
    i = int(r[0])
    d = launch.setdefault(i, {"name": r[ix["Kernel Name"]], "grid": r[ix["Grid Size"]]})
    d[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", "")) if r[ix["Metric Value"]] not in ("", "n/a") else 0.0
sel = [v for k, v in sorted(launch.items()) if a.first <= k < a.last]
if a.align:          # whole steps only: from the first stem launch to the last loss_reduce launch
    i0 = next(i for i, v in enumerate(sel) if "stem_tc" in v["name"])
    i1 = max(i for i, v in enumerate(sel) if "loss_reduce" in v["name"]) + 1
    sel = sel[i0:i1]
    if a.only_steps_with:      # keep the steps that contain this kernel (e.g. bigru_tc_kernel<32>: the pipeline-slot graphs)
        steps, cur = [], []
        for v in sel:
            if "stem_tc" in v["name"] and cur:
                steps.append(cur); cur = []
            cur.append(v)
        steps.append(cur)
        steps = [st for st in steps if any(a.only_steps_with in v["name"].replace("(int)", "") for v in st)]
        sel = [v for st in steps for v in st]
    a.steps = float(sum(1 for v in sel if "stem_tc" in v["name"]))
    print("aligned window: %d launches = %d whole steps" % (len(sel), a.steps))
def short(n):
    n = n.replace("void ", "").replace("sar::", "")
    n = re.sub(r"\(.*", "", n)
    return n
agg = {}
for v in sel:
    k = short(v["name"])
    g = agg.setdefault(k, [0, 0.0, 0.0])
    g[0] += 1; g[1] += v.get("gpu__time_duration.sum", 0.0) / 1e3
    g[2] += v.get("dram__bytes_read.sum", 0.0) + v.get("dram__bytes_write.sum", 0.0)
tot = sum(g[1] for g in agg.values())
print("| kernel | launches/step | us/step | share | DRAM MB/step |\n|---|---|---|---|---|")
for k, g in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %.1f | %.1f | %.1f %% | %.1f |" % (k, g[0] / a.steps, g[1] / a.steps, 100 * g[1] / tot, g[2] / a.steps / 1e6))
print("| total | %.1f | %.1f | | %.1f |" % (sum(g[0] for g in agg.values()) / a.steps, tot / a.steps, sum(g[2] for g in agg.values()) / a.steps / 1e6))
if a.traffic_json:
    # the three strided conv1's are the conv_tc_kernel launches right before a chain / slab run; Dense layers come after the ResNet
    # (sel holds whole steps: every conv_tc launch before a step's first LayerNorm minus the CNN_LIN GEMM right in front of it)
    conv, in_resnet = [], False
    for i, v in enumerate(sel):
        if "stem_tc" in v["name"]:
            in_resnet = True
        elif "layernorm" in v["name"]:
            if in_resnet and conv and "conv_tc_kernel" in conv[-1]["name"]:
                conv.pop()                      # CNN_LIN
            in_resnet = False
        elif in_resnet and "conv_tc" in v["name"]:
            conv.append(v)
    byt = sum(v.get("dram__bytes_read.sum", 0.0) + v.get("dram__bytes_write.sum", 0.0) for v in conv) / a.steps
    json.dump({"build_id": a.build_id, "B": a.B, "traffic_bytes_per_launch": byt / 36.0, "conv_kernel_launches_in_capture": len(conv),
               "dram_bytes_conv_per_step": byt,
               "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum (cold cache, serialised) over the block-conv launches of one eager step: %s" % a.csv},
              open(a.traffic_json, "w"), indent=1)
    print("conv launches", len(conv), "DRAM MB/step", byt / 1e6, "per layer-launch (36)", byt / 36 / 1e6)
