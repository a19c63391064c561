"""Data parallelism for the forward path: one process per GPU, utterance batch sharded by
rank, weights replicated, ONE collective per evaluation step -- an all-reduce(SUM) of the
8-float loss/metric vector over NCCL (NVLink 5 / NVSwitch).

Replaces `keras.utils.multi_gpu_model(model, gpus)` (model.py:193-194), which slices every
input on axis 0 into `gpus` contiguous parts inside one process and concatenates the
outputs on the CPU.  The same contiguous split rule is used here (`shard_slice`).

vector = [sum loss_accent, sum loss_disc, sum loss_ctc, sum loss_disc_bn,
          #correct_accent, #correct_disc, count, #correct_disc_bn]
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import numpy as np
import torch
import torch.distributed as dist

from .config import SARConfig


def init_from_env(backend: Optional[str] = None) -> Dict[str, int]:
    """torchrun-style init (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        be = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if be == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(be, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(be, rank=rank, world_size=world)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return {"rank": rank, "local_rank": local, "world_size": world}


def shard_slice(n: int, rank: int, world: int) -> slice:
    """Contiguous batch shard of rank `rank` -- multi_gpu_model's split rule: every replica
    gets n // world rows, the last one also takes the remainder."""
    step = n // world
    lo = rank * step
    hi = n if rank == world - 1 else lo + step
    return slice(lo, hi)


def shard_inputs(x: Dict[str, object], rank: int, world: int) -> Dict[str, object]:
    n = len(x["x_data"])
    sl = shard_slice(n, rank, world)
    return {k: v[sl] for k, v in x.items()}


def all_reduce_loss_vector(vec: torch.Tensor, group=None) -> torch.Tensor:
    """The path's single exchange step.  No-op when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    return vec


def loss_vector_to_metrics(v: np.ndarray, cfg: SARConfig) -> Dict[str, float]:
    """Keras-style batch means, accuracies and weighted total (model.py:344-367)."""
    n = max(float(v[6]), 1.0)
    w = cfg.loss_weights()
    m: Dict[str, float] = {}
    if cfg.ar_enable:
        m["y_accent_loss"] = float(v[0]) / n
        m["y_accent_acc"] = float(v[4]) / n
        if cfg.disc_enable:
            m["y_disc_loss"] = float(v[1]) / n
            m["y_disc_acc"] = float(v[5]) / n
    if cfg.ctc_enable:
        m["y_ctc_loss_loss"] = float(v[2]) / n
    if cfg.bn_dim and cfg.disc_enable:
        m["y_disc_bn_loss"] = float(v[3]) / n
        m["y_disc_bn_acc"] = float(v[7]) / n
    m["loss"] = sum(wk * m.get(k + "_loss", 0.0) for k, wk in w.items())
    m["count"] = float(v[6])
    return m


class DataParallelModel:
    """What compile(model, gpus>1) returns: same predict/evaluate surface; each rank
    processes its contiguous shard of the global batch."""

    def __init__(self, model, gpus: int):
        self.model = model
        self.gpus = gpus

    def __getattr__(self, name):
        return getattr(self.model, name)

    def _rank_world(self):
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
        return 0, 1

    def predict(self, x, batch_size=32, gather: bool = False, **kw):
        rank, world = self._rank_world()
        local = self.model.predict(shard_inputs(self.model._as_dict(x), rank, world), batch_size=batch_size, **kw)
        if not gather or world == 1:
            return local
        outs = local if isinstance(local, list) else [local]
        res = []
        for o in outs:
            t = o if isinstance(o, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(o))
            if dist.get_backend() == "nccl":
                t = t.cuda()
            sizes = [None] * world
            dist.all_gather_object(sizes, int(t.shape[0]))
            mx = max(sizes)
            pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            pad[:t.shape[0]] = t
            parts = [torch.empty_like(pad) for _ in range(world)]
            dist.all_gather(parts, pad)
            full = torch.cat([p[:s] for p, s in zip(parts, sizes)], 0)
            res.append(full if isinstance(o, torch.Tensor) else full.cpu().numpy())
        return res if isinstance(local, list) else res[0]

    def _shard_batches(self, generator):
        """multi_gpu_model's split (model.py:193-194): every global batch the generator yields is cut into contiguous
        per-replica shards; this rank keeps its own."""
        rank, world = self._rank_world()
        for item in generator:
            x, y = item if isinstance(item, tuple) else (item, None)
            xd = self.model._as_dict(x)
            n = len(xd["x_data"])
            sl = shard_slice(n, rank, world)
            yield shard_inputs(xd, rank, world), (None if y is None else {k: v[sl] for k, v in y.items()})

    def train_on_batch(self, x, y=None):
        return next(self.fit_batches([(x, y)]))

    def fit_batches(self, batches):
        for xs, ys in self._shard_batches(batches):
            yield self.model.train_on_batch(xs, ys)

    def fit_generator(self, generator, steps_per_epoch, epochs=1, **kw):
        """train.py:38-44 on the data-parallel model: every rank trains on its shard of each global batch, the gradients are
        all-reduced (mean) inside training.HeadTrainer -- the synchronous data parallelism of multi_gpu_model."""
        return self.model.fit_generator(self._shard_batches(generator), steps_per_epoch, epochs, **kw)

    def evaluate(self, x, y=None, batch_size=32):
        rank, world = self._rank_world()
        xs = shard_inputs(self.model._as_dict(x), rank, world)
        ys = None
        if y is not None:
            n = len(self.model._as_dict(x)["x_data"])
            sl = shard_slice(n, rank, world)
            ys = {k: v[sl] for k, v in y.items()}
        return self.model.evaluate(xs, ys, batch_size=batch_size)
