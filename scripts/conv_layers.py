"""Per-layer device time of the tcgen05 block convolutions at the bench shape (CUDA events, warm)."""
import os, sys, json, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from aesrc2020_b200 import model as mdl, utils as us, tc, ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
with contextlib.redirect_stdout(io.StringIO()):
    model, _ = mdl.SAR_Net((500, 80, 1), disc_enable=True, res_type="res34", res_filters=32, mto="gvlad",
                           vlad_clusters=64, ghost_clusters=8, metric_loss="arcface")
eng = model.engine()
rn = eng.resnet
x, _ = us.synthetic_batch(model.config, B, seed=1)
xd = model._to_device("x_data", x["x_data"])
# monkeypatch tc.conv_tc to time each call
calls = []
orig = tc.conv_tc
def timed(*a, **k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); orig(*a, **k); e1.record()
    calls.append((a[0], k, e0, e1))
tc.conv_tc = timed
for _ in range(3):
    calls.clear(); rn.forward(xd)
torch.cuda.synchronize()
acc = [0.0] * len(calls)
N = 10
for _ in range(N):
    calls.clear(); rn.forward(xd); torch.cuda.synchronize()
    for i, (a, k, e0, e1) in enumerate(calls):
        acc[i] += e0.elapsed_time(e1) * 1e3
tot = 0
for i, (a, k, e0, e1) in enumerate(calls):
    H, W = k["out_hw"]; cout = k["cout"]; ntaps = len(k["taps"][0]); cin = a.C
    sc = k.get("short"); flops = 2.0 * B * H * W * cout * (ntaps * cin + (sc.C if sc is not None else 0))
    us_ = acc[i] / N; tot += us_
    rows = B * (H + 1) * (W + 1)
    print("%2d  %3dx%-2d cin=%3d cout=%3d %s%s%s rows=%7d tiles=%5d  %7.1f us  %6.1f TF/s" % (
        i, H, W, cin, cout, "S2 " if a.split else "   ", "proj " if sc is not None else "     ",
        "res " if k.get("res") is not None else "    ", rows, -(-rows // 128) * (cout // min(cout, 128)), us_, flops / us_ / 1e6))
print("total %.1f us" % tot)
import ctypes
