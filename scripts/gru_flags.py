"""Device time of the Bi-GRU recurrence kernel alone (sar_bigru_fwd), B utterances x 48 steps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from aesrc2020_b200 import _shim
from aesrc2020_b200._shim import ptr, stream_ptr

B, S, U = int(sys.argv[1]) if len(sys.argv) > 1 else 64, 48, 256
xp = torch.randn(B, S, 2, 3 * U, device="cuda") * 0.1
rec = torch.randn(2, U, 3 * U, device="cuda") * 0.05
rb = torch.randn(2, 3 * U, device="cuda") * 0.1
out = torch.empty(B, S, 2 * U, device="cuda")
lib = _shim.lib()
for flags, name in ((1, "bigru recurrence"),):
    for _ in range(3):
        lib.sar_bigru_fwd(ptr(xp), ptr(rec), ptr(rb), ptr(out), B, S, U, flags, stream_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        lib.sar_bigru_fwd(ptr(xp), ptr(rec), ptr(rb), ptr(out), B, S, U, flags, stream_ptr())
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print("%-20s %8.1f us  (%.2f us/step)" % (name, us, us / S))
