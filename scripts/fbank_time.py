"""Front-end alone: sar_fbank_pcm16_fwd at B x T (CUDA events over back-to-back calls; run under
`ncu --metrics gpu__time_duration.sum -k regex:fbank` for the per-kernel split).  Usage: fbank_time.py [B] [T] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from aesrc2020_b200 import fbank as fb, ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 500
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
dev = torch.device("cuda")
n = fb.FRAME_LEN + (T - 1) * fb.FRAME_STEP
rng = np.random.RandomState(1)
pcm = torch.from_numpy((rng.randn(B * n) * 3000).clip(-32768, 32767).astype(np.int16)).to(dev)
offs = torch.from_numpy(np.arange(B + 1, dtype=np.int64) * n).to(dev)
melfb = torch.from_numpy(np.ascontiguousarray(fb.mel_filterbank().T, dtype=np.float32)).to(dev)
ws = torch.empty((B, T, 80), device=dev); x = torch.empty((B, T, 80), device=dev)
for _ in range(3):
    ops.fbank(pcm, offs, melfb, T, T, out=x, feat_ws=ws)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    ops.fbank(pcm, offs, melfb, T, T, out=x, feat_ws=ws)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / reps
byt = B * (2.0 * n + 4.0 * T * 80)
print("B=%d T=%d: %.1f us per batch, %.1f GB/s algorithmic" % (B, T, us, byt / us / 1e3))
