"""One whole-model training step (HeadTrainer(train_resnet=True, train_ctc=True), configs[4] graph, B=16) -- run under
`ncu --metrics gpu__time_duration.sum` to see which training kernels the step time goes to."""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from aesrc2020_b200 import model as mdl, training as T, utils as us
B = int(os.environ.get("TF_B", "16"))
with contextlib.redirect_stdout(io.StringIO()):
    model, _ = mdl.SAR_Net((500, 80, 1), **dict(bench.CONFIGS["cfg5"]["kw"]))
x, y = us.synthetic_batch(model.config, B, seed=300)
tr = T.HeadTrainer(model, lr=0.005, train_resnet=True, train_ctc=True)
xd = {k: model._to_device(k, v) for k, v in x.items()}
for _ in range(int(os.environ.get("TF_STEPS", "1"))):
    print(tr.train_on_batch(xd, y))
torch.cuda.synchronize()
