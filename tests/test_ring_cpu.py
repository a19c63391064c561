"""utils.PinnedRing / data_generator host logic (the stand-in for Keras' generator queue, train.py:44)."""
import threading
import time

import numpy as np
import pytest

from aesrc2020_b200 import utils as us


def _batches(n, B=4, T=6):
    for i in range(n):
        yield {"x_data": np.full((B, T, 3, 1), float(i), np.float32), "x_accent": np.eye(8, dtype=np.float32)[[i % 8] * B]}, \
              {"y_accent": np.full((B,), i)}


def test_ring_preserves_order_values_and_targets():
    with us.PinnedRing(_batches(25), max_queue_size=4, keep=3) as ring:
        seen = []
        for x, y in ring:
            assert isinstance(x, us.RingBatch) and x.ready is None
            assert float(x["x_data"][0, 0, 0, 0]) == len(seen) and int(y["y_accent"][0]) == len(seen)
            seen.append(x)
        assert len(seen) == 25
        assert ring.staged_bytes == 25 * (4 * 6 * 3 * 4 + 4 * 8 * 4)
        with pytest.raises(StopIteration):
            next(ring)


def test_ring_slots_stay_valid_for_keep_batches_and_are_bounded():
    keep, mq = 3, 2
    ring = us.PinnedRing(_batches(40), max_queue_size=mq, keep=keep)
    held = []
    for i, (x, _) in enumerate(ring):
        held.append((i, x["x_data"]))
        for j, a in held[-keep:]:                       # the last `keep` hand-outs are still intact
            assert float(a.flat[0]) == j, (i, j)
    ring.close()
    bufs = {a.__array_interface__["data"][0] for _, a in held}
    assert len(bufs) == mq + keep                       # a RING: a fixed set of buffers, reused


def test_ring_backpressure_and_error_propagation():
    produced = []

    def gen():
        for i in range(100):
            produced.append(i)
            yield {"x_data": np.zeros((1, 2), np.float32)}
    ring = us.PinnedRing(gen(), max_queue_size=3, keep=2)
    time.sleep(0.3)
    assert len(produced) <= 3 + 2                       # the producer does not run ahead of the queue bound
    next(ring)
    ring.close()

    def bad():
        yield {"x_data": np.zeros((1, 2), np.float32)}
        raise RuntimeError("loader failed")
    ring = us.PinnedRing(bad(), max_queue_size=2)
    next(ring)
    with pytest.raises(RuntimeError, match="loader failed"):
        next(ring)
    ring.close()
    assert not any(t.name == "sarnet-pinned-ring" and t.is_alive() for t in threading.enumerate())


def test_data_generator_shuffles_and_cuts_batches_like_the_reference():
    """utils.py:120-154: n_batchs = len(lst) // batch_size per pass, reshuffled every pass, endless."""
    calls = []
    real = us.data_loader
    us.data_loader = lambda sub, **kw: (calls.append(list(sub)) or ({"x_data": sub}, {}))
    try:
        lst = ["u%d" % i for i in range(11)]
        g = us.data_generator(lst, batch_size=4, data_dct={}, seed=3)
        got = [next(g) for _ in range(6)]               # 2 batches per pass -> 3 passes
    finally:
        us.data_loader = real
    assert all(len(c) == 4 for c in calls) and len(calls) == 6
    for p in range(3):
        a, b = calls[2 * p], calls[2 * p + 1]
        assert not set(a) & set(b) and set(a + b) <= set(lst)
    assert calls[0] + calls[1] != calls[2] + calls[3]   # reshuffled
    with pytest.raises(ValueError):
        next(us.data_generator(lst, batch_size=64, data_dct={}))
