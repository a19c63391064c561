"""torchrun --nproc-per-node 2 scripts/dp_fit_check.py -- data-parallel training check: SAR_Net(gpus=2) -> train_model.fit_generator
on GLOBAL batches (every rank shards them like multi_gpu_model, gradients all-reduced): the loss falls and both ranks end with
bitwise identical weights."""
import os, sys, warnings, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from aesrc2020_b200 import model as mdl, utils as us

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
kw = dict(ctc_enable=True, ar_enable=True, disc_enable=True, res_type="res34", res_filters=32, mto="gvlad", vlad_clusters=8,
          ghost_clusters=2, metric_loss="arcface", margin=0.3, bpe_classes=40, max_ctc_len=4)
with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
    warnings.simplefilter("ignore")
    model, train_model = mdl.SAR_Net((200, 80, 1), lr=0.004, gpus=world, **kw)
cfg = model.config
batches = [us.synthetic_batch(cfg, 8 * world, seed=50 + i) for i in range(3)]          # the same GLOBAL batches on every rank

def gen():
    while True:
        for b in batches:
            yield b

hist = train_model.fit_generator(gen(), steps_per_epoch=6, epochs=2, verbose=0)
flat = torch.cat([torch.from_numpy(model.weights[k].reshape(-1)) for k in sorted(model.weights)]).cuda()
ref = flat.clone()
dist.broadcast(ref, 0)
same = bool(torch.equal(flat, ref))
nbad = [k for k in sorted(model.weights)]
print("rank %d: loss %.3f -> %.3f, max |w - w_rank0| = %.3e" % (rank, hist[0]["loss"], hist[-1]["loss"], float((flat - ref).abs().max())), flush=True)
ok = torch.tensor([1 if (same and hist[-1]["loss"] < hist[0]["loss"]) else 0], device="cuda")
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
if rank == 0:
    print("dp_fit_check: world %d, loss %.3f -> %.3f, weights identical across ranks: %s -> %s"
          % (world, hist[0]["loss"], hist[-1]["loss"], same, "OK" if int(ok) else "FAILED"))
dist.destroy_process_group()
sys.exit(0 if int(ok) else 1)
