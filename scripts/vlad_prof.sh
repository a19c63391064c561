#!/bin/bash
# per-phase cycle stamps of vlad_tc_kernel (CTA 0, its second item): rebuilds the library with -DSAR_VLAD_PROFILE,
# runs B=512 (several items per CTA), restores the normal build.
SAR_NVCC_EXTRA=-DSAR_VLAD_PROFILE python aesrc2020_b200/csrc/build.py --force > /dev/null
python - <<'PY'
import torch, numpy as np
from aesrc2020_b200 import tc
for B in (512,):
    S, D, K, G = 48, 256, 64, 8
    rng = np.random.RandomState(5)
    wa = torch.from_numpy(tc.pack_vlad_assign((rng.randn(D, K + G) / 16 * 3).astype(np.float32))).cuda()
    ba = torch.zeros(K + G, device="cuda"); cen = torch.randn(K + G, D, device="cuda") / 16
    xp = tc.Planes((torch.randn(2, B * S, D, device="cuda") * 0.5).half(), 1, B * S, 1, D, False)
    y = tc.alloc_rows(B, K * D, "cuda")
    for _ in range(2):
        tc.vlad_tc(xp, wa, ba, cen, B, S, K, G, planes=y, want_dense=False)
    torch.cuda.synchronize()
PY
python aesrc2020_b200/csrc/build.py --force > /dev/null
