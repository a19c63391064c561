"""Per-layer device time of the tcgen05 block convolutions (CUDA events around every launch, warm, no stage chains),
next to each layer's tensor-pipe floor for the three hi/lo products (issued flops incl. flat-pad rows / 8192 flop/clk/SM).
Usage: python scripts/conv_layers.py [B] [T]"""
import os, sys, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from aesrc2020_b200 import model as mdl, utils as us, tc
from aesrc2020_b200.engine import StepOpts

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 500
with contextlib.redirect_stdout(io.StringIO()):
    model, _ = mdl.SAR_Net((T, 80, 1), disc_enable=True, res_type="res34", res_filters=32, mto="gvlad",
                           vlad_clusters=64, ghost_clusters=8, metric_loss="arcface")
eng = model.engine()
rn = eng.resnet
x, _ = us.synthetic_batch(model.config, B, seed=1)
xd = model._to_device("x_data", x["x_data"])
opts = StepOpts(no_chain=True)
calls = []
orig = tc.conv_launch
def timed(d):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); orig(d); e1.record()
    calls.append((d, e0, e1))
tc.conv_launch = timed
for _ in range(3):
    calls.clear(); rn.forward(xd, opts=opts)
torch.cuda.synchronize()
N = 10
acc = [0.0] * len(calls)
for _ in range(N):
    calls.clear(); rn.forward(xd, opts=opts); torch.cuda.synchronize()
    for i, (d, e0, e1) in enumerate(calls):
        acc[i] += e0.elapsed_time(e1) * 1e3
SMS, CLK = 148, 1.965e3        # MHz -> cycles per us
tot = totf = 0.0
stage_t, stage_f = {}, {}
for i, (d, e0, e1) in enumerate(calls):
    rows = d.B * (d.H + 1) * (d.W + 1)
    tiles = -(-rows // 128)
    k = d.ntaps * d.a_ch + (d.s_ch if d.s else 0)
    issued = 3 * 2.0 * tiles * 128 * k * d.cout                  # three hi/lo products on whole tiles
    floor_us = issued / 8192.0 / SMS / CLK
    us_ = acc[i] / N
    tot += us_; totf += floor_us
    key = (d.H, d.W)
    stage_t[key] = stage_t.get(key, 0.0) + us_; stage_f[key] = stage_f.get(key, 0.0) + floor_us
    print("%2d  %3dx%-2d cin=%3d cout=%3d taps=%d %s%s rows=%7d mtiles=%5d  %7.1f us  floor %6.1f us  %4.0f %%" % (
        i, d.H, d.W, d.a_ch, d.cout, d.ntaps, "proj " if d.s else "     ", "res " if (d.res or d.res_f32) else "    ",
        rows, tiles, us_, floor_us, 100 * floor_us / us_))
for key in stage_t:
    print("stage %3dx%-2d  %7.1f us  floor %6.1f us  %4.0f %%" % (key[0], key[1], stage_t[key], stage_f[key], 100 * stage_f[key] / stage_t[key]))
print("total %.1f us  floor %.1f us (nominal 8192 flop/clk/SM, %d SMs, %.0f MHz)" % (tot, totf, SMS, CLK))
