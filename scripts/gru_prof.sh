#!/bin/bash
# per-phase cycle stamps of the tcgen05 Bi-GRU step loop (a profiling build in a scratch copy of the library)
set -e
cd "$(dirname "$0")/.."
SAR_NVCC_EXTRA=-DSAR_GRU_PROFILE python aesrc2020_b200/csrc/build.py --force > /dev/null
python - <<'PY'
import torch
from aesrc2020_b200 import _shim
from aesrc2020_b200._shim import ptr, stream_ptr
lib = _shim.lib()
B, S = 64, 48
xp = torch.randn(B, S, 2, 768, device="cuda") * 0.1
rec = torch.randn(2, 256, 768, device="cuda") / 16
rb = torch.zeros(2, 768, device="cuda")
out = torch.empty(B, S, 512, device="cuda")
for _ in range(2):
    lib.sar_bigru_fwd(ptr(xp), ptr(rec), ptr(rb), ptr(out), B, S, 256, 1, stream_ptr())
    torch.cuda.synchronize()
PY
