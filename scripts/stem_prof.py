"""Times sar_stem_pool_fwd alone (B=64, T=500).  With a -DSAR_STEM_PROFILE build CTA 0 prints per-role cycles."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from aesrc2020_b200 import tc
B, T, F0 = int(os.environ.get("B", 64)), 500, 64
rng = np.random.RandomState(0)
d = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()
x = d(rng.rand(B, T, 80, 1)); w = d(rng.randn(7, 7, 1, F0) * 0.2); b = d(rng.randn(F0) * 0.1)
s = d(rng.uniform(0.7, 1.3, F0)); t = d(rng.randn(F0) * 0.1)
out = tc.alloc_planes(B, 125, 20, F0, False, "cuda")
tc.stem_pool(x, w, b, s, t, out); torch.cuda.synchronize()
if os.environ.get("ONCE"): sys.exit(0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): tc.stem_pool(x, w, b, s, t, out)
e1.record(); torch.cuda.synchronize()
print("stem_pool B=%d: %.1f us" % (B, e0.elapsed_time(e1) * 1e3 / 20))
