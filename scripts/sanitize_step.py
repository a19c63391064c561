"""One small step of the hot path, eager and as a replayed CUDA graph (single stream and a pipeline slot) --
the workload `scripts/sanitize.sh` runs under compute-sanitizer and ncu (graph-node profiling)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import io
import contextlib

import numpy as np
import torch

from aesrc2020_b200 import model as mdl, utils as us

B = int(os.environ.get("SAN_B", "4"))
T = int(os.environ.get("SAN_T", "200"))
kw = dict(ctc_enable=True, ar_enable=True, disc_enable=True, res_type="res34", res_filters=32, mto="gvlad",
          vlad_clusters=64, ghost_clusters=8, metric_loss="arcface", margin=0.3)
with contextlib.redirect_stdout(io.StringIO()):
    model, _ = mdl.SAR_Net((T, 80, 1), **kw)
x, _ = us.synthetic_batch(model.config, B, seed=3)
eng = model.engine()
dev = {k: model._to_device(k, v) for k, v in x.items()}
names = ("y_accent", "y_disc", "y_ctc_loss")
eager = {k: v.clone() for k, v in eng.forward(dev).items() if k in names}
torch.cuda.synchronize()
for rep in range(2):                       # capture + replay, then a second replay
    g = {k: v.clone() for k, v in eng.forward_graphed(dev).items() if k in names}
torch.cuda.synchronize()
out, st = eng.forward_slot(dev, 1)
st.synchronize()
slot = {k: v.clone() for k, v in out.items() if k in names}
for k in names:
    a, b, c = eager[k].cpu().numpy(), g[k].cpu().numpy(), slot[k].cpu().numpy()
    assert np.array_equal(a, b), ("graph replay differs from eager", k)
    assert np.allclose(a, c, rtol=1e-4, atol=1e-6), ("slot path differs", k)
print("sanitize_step ok B=%d T=%d" % (B, T))
