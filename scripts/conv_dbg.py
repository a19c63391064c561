"""clock64 stamps of CTA 0 of chosen residual-block conv layers (sar_tc_conv.dbg): prologue, per-tile TMA / MMA issue, and
per epilogue item (quadrant, 32-column chunk): entered / accumulator ready / math+staging done / finished.
Usage: python scripts/conv_dbg.py [B] [layer indices...]"""
import os, sys, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from aesrc2020_b200 import model as mdl, utils as us, tc
from aesrc2020_b200.engine import StepOpts
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
layers = [int(v) for v in sys.argv[2:]] or [2, 3, 9, 17, 31]
with contextlib.redirect_stdout(io.StringIO()):
    model, _ = mdl.SAR_Net((500, 80, 1), disc_enable=True, res_type="res34", res_filters=32, mto="gvlad", vlad_clusters=64, ghost_clusters=8, metric_loss="arcface")
eng = model.engine(); rn = eng.resnet
x, _ = us.synthetic_batch(model.config, B, seed=1)
xd = model._to_device("x_data", x["x_data"])
opts = StepOpts(no_chain=True)
orig = tc.conv_desc
dbgs = []
def timed(*a, **k):
    d = torch.zeros(512, dtype=torch.int64, device="cuda"); dbgs.append(d)
    return orig(*a, dbg=d, **k)
rn.forward(xd, opts=opts); torch.cuda.synchronize()
tc.conv_desc = timed
rn.forward(xd, opts=opts); torch.cuda.synchronize()
for li in layers:
    d = dbgs[li].cpu().tolist()
    t0 = d[62]
    print("layer %d: prologue done %d  exit %d" % (li, d[61] - t0, d[63] - t0))
    for it in range(8):
        m = [d[it*4+j] - t0 for j in range(4)]
        if d[it*4] == 0: continue
        print("  tile %d: mma loop start %6d  acc free %6d  first slab %6d  all issued %6d" % (it, m[0], m[1], m[2], m[3]))
    for it in range(2):
        for c in range(4):
            for q in range(4):
                o = 64 + ((it * 4 + c) * 4 + q) * 8
                if d[o] == 0: continue
                e = [d[o + j] - t0 for j in range(4)]
                print("    item tile %d chunk %d quad %d: entered %6d  acc ready %6d  math done %6d (+%d)  finished %6d (+%d)" % (it, c, q, e[0], e[1], e[2], e[2]-e[1], e[3], e[3]-e[2]))
