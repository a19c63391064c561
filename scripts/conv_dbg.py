import os, sys, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from aesrc2020_b200 import model as mdl, utils as us, tc
B = 64
with contextlib.redirect_stdout(io.StringIO()):
    model, _ = mdl.SAR_Net((500, 80, 1), disc_enable=True, res_type="res34", res_filters=32, mto="gvlad", vlad_clusters=64, ghost_clusters=8, metric_loss="arcface")
eng = model.engine(); rn = eng.resnet; rn.chain_stages = set()
x, _ = us.synthetic_batch(model.config, B, seed=1)
xd = model._to_device("x_data", x["x_data"])
calls = []
orig = tc.conv_desc
dbgs = []
def timed(*a, **k):
    d = torch.zeros(64, dtype=torch.int64, device="cuda"); dbgs.append(d)
    return orig(*a, dbg=d, **k)
rn.forward(xd); torch.cuda.synchronize()
tc.conv_desc = timed
rn.forward(xd); torch.cuda.synchronize()
for li in (2, 3, 5, 8, 9):
    d = dbgs[li].cpu().tolist()
    t0 = d[62]
    print("layer", li, "entry 0  prologue done %d  exit %d" % (d[61] - t0, d[63] - t0))
    for it in range(4):
        m = [d[it*4+j] - t0 for j in range(4)]; e = [d[32+it*6+j] - t0 for j in range(6)]
        if m[0] < 0: continue
        print("  tile %d: mma start %6d  tempty ok %6d  slab ok %6d  issued %6d | epi chunk0: tfull %6d  +tmem %5d  +res/math %5d  +split/STS %5d  +write-out %5d" % (it, m[0], m[1], m[2], m[3], e[0], e[1]-e[0], e[2]-e[1], e[3]-e[2], e[5]-e[3]))
