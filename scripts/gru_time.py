"""Bi-GRU recurrence: device time of sar_bigru_fwd (tcgen05 kernel; SAR_GRU_FFMA=1 = the CUDA-core kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from aesrc2020_b200 import _shim
from aesrc2020_b200._shim import ptr, stream_ptr
lib = _shim.lib()
def run(B, S, n=20):
    xp = torch.randn(B, S, 2, 768, device="cuda") * 0.1
    rec = torch.randn(2, 256, 768, device="cuda") / 16
    rb = torch.zeros(2, 768, device="cuda")
    out = torch.empty(B, S, 512, device="cuda")
    f = lambda: lib.sar_bigru_fwd(ptr(xp), ptr(rec), ptr(rb), ptr(out), B, S, 256, 1, stream_ptr())
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for B in (8, 64, 128, 512):
    for S in (12, 48):
        us = run(B, S)
        print("%s B=%d S=%d  %.1f us  (%.2f us/step)" % ("ffma" if os.environ.get("SAR_GRU_FFMA") else "tc", B, S, us, us / S))
