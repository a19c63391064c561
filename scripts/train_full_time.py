import contextlib, io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from aesrc2020_b200 import model as mdl, training as T, utils as us
B = int(os.environ.get("TF_B", "64"))
with contextlib.redirect_stdout(io.StringIO()):
    model, _ = mdl.SAR_Net((500, 80, 1), **dict(bench.CONFIGS["cfg5"]["kw"]))
x, y = us.synthetic_batch(model.config, B, seed=300)
tr = T.HeadTrainer(model, lr=0.005, train_resnet=True, train_ctc=True)
xd = {k: model._to_device(k, v) for k, v in x.items()}
for i in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    s0 = torch.cuda.memory_stats()
    out = tr.train_on_batch(xd, y)
    t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    s1 = torch.cuda.memory_stats()
    print("step %d: host %.1f ms, total %.1f ms, cudaMalloc calls %d, retries %d, reserved %.1f GB" % (i, (t1 - t0) * 1e3, (t2 - t0) * 1e3,
          s1["num_device_alloc"] - s0["num_device_alloc"], s1["num_alloc_retries"] - s0["num_alloc_retries"], s1["reserved_bytes.all.current"] / 1e9))
