"""Why three tensor-core products per k-step?  Model-level error of cheaper operand schemes against the float64 oracle.
Emulated with the shipped kernels: rounding the conv / Dense / GRU weights to fp16 on the host makes every B_lo plane zero,
i.e. the `A x B_lo` product contributes nothing -- the two-product scheme "activations hi+lo, weights fp16".  (The other
two-product scheme, fp16 activations with hi+lo weights, has the same rounding magnitude on the other operand.)
Prints the error of the full scheme and of the emulated two-product scheme on the bench configuration's outputs."""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from aesrc2020_b200 import model as mdl, utils as us
from oracle import sarnet_oracle as O

kw = dict(ctc_enable=True, ar_enable=True, disc_enable=True, res_type="res34", res_filters=32, mto="gvlad", vlad_clusters=64,
          ghost_clusters=8, metric_loss="arcface", margin=0.3)
T, B = 500, 8
with contextlib.redirect_stdout(io.StringIO()):
    model, _ = mdl.SAR_Net((T, 80, 1), **kw)
x, _ = us.synthetic_batch(model.config, B, seed=7)
ref = O.sar_net_forward(model.weights, x, **model.config.model_kwargs())

def errs(m):
    outs = m.predict(x, batch_size=B)
    res = {}
    for name, got in zip(m.config.output_names(), outs):
        want = ref[name].numpy()
        res[name] = float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-6)))
    return res

print("3 products (shipped):                         ", {k: "%.2e" % v for k, v in errs(model).items()})
w16 = {k: (v.astype(np.float16).astype(np.float32) if (k.endswith("kernel") and v.ndim >= 2) else v) for k, v in model.weights.items()}
with contextlib.redirect_stdout(io.StringIO()):
    m2, _ = mdl.SAR_Net((T, 80, 1), weights=w16, **kw)
print("2 products (weights rounded to fp16, B_lo = 0):", {k: "%.2e" % v for k, v in errs(m2).items()})
print("tolerance (north_star): 1e-3 relative")
