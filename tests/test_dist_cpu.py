"""The N>1 host logic on CPU: world_size-2 gloo run of the loss-vector all-reduce and the
multi_gpu_model split rule (model.py:193-194)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from aesrc2020_b200 import dist as sdist
from aesrc2020_b200.config import SARConfig


def test_shard_slice_is_multi_gpu_model_rule():
    for n, g in ((64, 2), (65, 4), (7, 8), (4096, 8)):
        parts = [sdist.shard_slice(n, r, g) for r in range(g)]
        idx = np.concatenate([np.arange(n)[s] for s in parts])
        assert np.array_equal(idx, np.arange(n))
        sizes = [len(np.arange(n)[s]) for s in parts]
        assert all(s == n // g for s in sizes[:-1]) and sizes[-1] == n - (g - 1) * (n // g)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    env = sdist.init_from_env(backend="gloo")
    assert env["world_size"] == world
    rng = np.random.RandomState(0)
    stats = rng.rand(10, 8).astype(np.float32)               # per-utterance contributions of a global batch of 10
    stats[:, 6] = 1.0
    sl = sdist.shard_slice(10, rank, world)
    vec = torch.from_numpy(stats[sl].sum(0))
    vec = sdist.all_reduce_loss_vector(vec)
    if rank == 0:
        np.save(out, vec.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_loss_vector_allreduce_gloo_world2(tmp_path):
    out = str(tmp_path / "v.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.start_processes(_worker, args=(2, port, out), nprocs=2, join=True, start_method="spawn")
    got = np.load(out)
    rng = np.random.RandomState(0)
    stats = rng.rand(10, 8).astype(np.float32)
    stats[:, 6] = 1.0
    assert np.allclose(got, stats.sum(0), rtol=1e-6)
    cfg = SARConfig(ctc_enable=True, disc_enable=True, mto="gvlad", metric_loss="circleloss")
    m = sdist.loss_vector_to_metrics(got, cfg)
    assert m["count"] == 10 and abs(m["y_ctc_loss_loss"] - stats[:, 2].mean()) < 1e-6
    want_total = 0.01 * stats[:, 0].mean() + 0.6 * stats[:, 1].mean() + 0.01 * stats[:, 2].mean()
    assert abs(m["loss"] - want_total) < 1e-6                # model.py:344-367 weights (Q5)


def _grad_worker(rank, world, port, q):
    import os
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from aesrc2020_b200.training import HeadTrainer
    tr = HeadTrainer.__new__(HeadTrainer)                     # only the collective: no device, no model
    tr.keys, tr.group = ["a", "b"], None
    g = {"a": torch.full((3, 2), float(rank + 1)), "b": torch.arange(4, dtype=torch.float32) * (rank + 1)}
    tr._all_reduce(g)
    # ... and the BN moving averages every replica updated with its own shard's statistics
    tr.stat_keys, tr.p = ["bn/moving_mean", "bn/moving_variance"], {"bn/moving_mean": torch.full((3,), float(rank)),
                                                                    "bn/moving_variance": torch.full((3,), 2.0 + 2 * rank)}
    tr._sync_stats()
    assert tr.p["bn/moving_mean"].tolist() == [0.5] * 3 and tr.p["bn/moving_variance"].tolist() == [3.0] * 3
    q.put((rank, g["a"].tolist(), g["b"].tolist()))
    dist.destroy_process_group()


def test_gradient_all_reduce_is_the_mean_over_replicas():
    """training.HeadTrainer._all_reduce: ONE flat all-reduce, mean over the replicas (gloo here, NCCL on the GPUs); _sync_stats:
    the BN moving averages are averaged too (asserted inside the workers)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500) + 7
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, a, b in res:
        assert a == [[1.5, 1.5]] * 3 and b == [0.0, 1.5, 3.0, 4.5]


def test_data_parallel_fit_shards_every_global_batch_like_multi_gpu_model():
    """dist.DataParallelModel._shard_batches (the generator fit_generator trains on when gpus > 1): every (inputs, targets)
    batch is cut into the contiguous per-replica slices of multi_gpu_model (model.py:193-194); the slices of all ranks tile
    the global batch, inputs and targets stay aligned.  Host logic only (no device, no process group: rank/world stubbed)."""
    class _Model:
        @staticmethod
        def _as_dict(x):
            return dict(x)

    n = 11
    x = {"x_data": np.arange(n * 3, dtype=np.float32).reshape(n, 3), "x_accent": np.arange(n)[:, None].astype(np.float32)}
    y = {"y_accent": np.arange(n)[:, None].astype(np.float32) + 100}
    for world in (1, 2, 4):
        seen_x, seen_y = [], []
        for rank in range(world):
            dp = sdist.DataParallelModel(_Model(), world)
            dp._rank_world = lambda r=rank, w=world: (r, w)
            (xs, ys), = list(dp._shard_batches([(x, y)]))
            assert len(xs["x_data"]) == len(xs["x_accent"]) == len(ys["y_accent"])
            assert np.array_equal(ys["y_accent"][:, 0] - 100, xs["x_accent"][:, 0])      # inputs and targets of the same utterances
            seen_x.append(xs["x_data"]); seen_y.append(ys["y_accent"])
        assert np.array_equal(np.concatenate(seen_x), x["x_data"]) and np.array_equal(np.concatenate(seen_y), y["y_accent"])
    dp = sdist.DataParallelModel(_Model(), 2)
    dp._rank_world = lambda: (1, 2)
    (xs, ys), = list(dp._shard_batches([x]))                  # a generator that yields inputs only
    assert ys is None and len(xs["x_data"]) == n - n // 2
