"""Per-CTA, per-item clock64 stamps of the persistent stage-chain launches (ChainParams.dbg, conv_tc.cu): where a
stage's time goes -- waiting for the tiles of the previous layer, lead-in (slab TMA), MMA issue, epilogue, publish.
Usage: python scripts/chain_dbg.py [B] [stage ...]      (stages 2 3 4 by default)"""
import os, sys, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from aesrc2020_b200 import model as mdl, utils as us, tc
from aesrc2020_b200.engine import StepOpts

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
want = [int(v) for v in sys.argv[2:]] or [2, 3, 4]
STRIDE, ITEMS = 8 + 32 * 8, 32
with contextlib.redirect_stdout(io.StringIO()):
    model, _ = mdl.SAR_Net((500, 80, 1), disc_enable=True, res_type="res34", res_filters=32, mto="gvlad", vlad_clusters=64,
                           ghost_clusters=8, metric_loss="arcface")
eng = model.engine(); rn = eng.resnet
x, _ = us.synthetic_batch(model.config, B, seed=1)
xd = model._to_device("x_data", x["x_data"])
opts = StepOpts()
orig = tc.conv_tc_chain
bufs = []
def traced(descs, ws, max_ctas=0):
    d = torch.zeros(148 * STRIDE + 512, dtype=torch.int64, device="cuda")
    descs[0].dbg = d.data_ptr()
    bufs.append((d, len(descs), int(descs[0].cout), int(descs[0].B * (descs[0].H + 1) * (descs[0].W + 1))))
    return orig(descs, ws, max_ctas=max_ctas)
for _ in range(3):
    rn.forward(xd, opts=opts)
torch.cuda.synchronize()
tc.conv_tc_chain = traced
rn.forward(xd, opts=opts); torch.cuda.synchronize()
tc.conv_tc_chain = orig
for ci, (d, nph, cout, rows) in enumerate(bufs):
    stage = ci + 2
    if stage not in want:
        continue
    full = d.cpu().numpy()
    a = full[:148 * STRIDE].reshape(148, STRIDE)
    e = full[148 * STRIDE:]
    used = a[:, 0] != 0
    ncta = int(used.sum())
    it = a[:, 8:].reshape(148, ITEMS, 8)
    bn = 64 if cout % 64 == 0 else 32
    tiles = ((rows + 127) // 128) * (cout // bn)
    print("== stage %d: %d layers, cout %d, %d tiles/layer, %d CTAs, %.2f rounds/layer" % (stage, nph, cout, tiles, ncta, tiles / ncta))
    life = (a[used, 2] - a[used, 0])
    print("   CTA lifetime cycles: min %d median %d max %d" % (life.min(), np.median(life), life.max()))
    gt = a[used, 1] - a[used, 1].min()
    print("   CTA start skew (globaltimer ns): max %d" % gt.max())
    # per-item phases, averaged over CTAs
    tot = dict(wait=0.0, lead=0.0, mma=0.0, epi0=0.0, epil=0.0, n=0)
    for c in range(148):
        if not used[c]:
            continue
        for k in range(ITEMS):
            r = it[c, k]
            if r[0] == 0:
                break
            tot["wait"] += r[2] - r[1]; tot["lead"] += r[3] - r[2]; tot["mma"] += r[4] - r[3]
            tot["epi0"] += r[6] - r[5]; tot["epil"] += r[7] - r[4]; tot["n"] += 1
    n = max(tot["n"], 1)
    print("   per item (mean over %d items): dependency wait %.0f | deps met -> first slab %.0f | MMA issue %.0f | chunk-0 epilogue (enter->published) %.0f | all issued -> last chunk published %.0f"
          % (n, tot["wait"] / n, tot["lead"] / n, tot["mma"] / n, tot["epi0"] / n, tot["epil"] / n))
    for c in (0, 1, 40, 100, ncta - 1):
        if c >= 148 or not used[c]:
            continue
        t0 = a[c, 0]
        print("   CTA %d (exit %d):" % (c, a[c, 2] - t0))
        for k in range(ITEMS):
            r = it[c, k]
            if r[0] == 0:
                break
            item = int(r[0]) - 1
            print("     item %5d (layer %2d tile %3d): reached %7d  deps met %7d (+%d)  first slab %7d  all issued %7d (+%d)  epi enter %7d  chunk0 published %7d  last published %7d"
                  % (item, item // tiles, item % tiles, r[1] - t0, r[2] - t0, r[2] - r[1], r[3] - t0, r[4] - t0, r[4] - r[3], r[5] - t0, r[6] - t0, r[7] - t0))
    if os.environ.get("SAR_CHAIN_DBG_PHASE"):
        t0 = a[0, 0]
        print("   epilogue items of CTA 0, layer %s (first two items):" % os.environ["SAR_CHAIN_DBG_PHASE"])
        for k in range(2):
            for c in range(4):
                for q in range(4):
                    o = 64 + ((k * 4 + c) * 4 + q) * 8
                    if e[o] == 0:
                        continue
                    v = [e[o + j] - t0 for j in range(7)]
                    print("     item %d chunk %d quad %d: entered %7d  acc ready %7d (+%d)  math done %7d (+%d)  stores issued +%d  landed +%d  fence +%d  published %7d (+%d)"
                          % (k, c, q, v[0], v[1], v[1] - v[0], v[2], v[2] - v[1], v[4] - v[2], v[5] - v[4], v[6] - v[5], v[3], v[3] - v[2]))
