"""Mirror of the parts of the reference's utils.py that sit on the hot path's boundary:
the input dict contract (utils.py:102-116), label padding (utils.py:57-63) and the shape
known-answer `cal_descriptors` (utils.py:156-159), `data_loader` / `data_generator`
(utils.py:71-154) with the per-batch work on the device, and `PinnedRing`, the background
producer that stands in for Keras' generator queue (train.py:44 `max_queue_size=20`).
`synthetic_batch` provides the seeded synthetic inputs the benchmarks and tests use.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

from .config import SARConfig

UNK_ID, SOS_ID, EOS_ID = 0, 1, 2          # utils.py:53-55


def text_ids_norm(ids, max_len):
    """utils.py:57-63: truncate to max_len, pad with EOS_ID."""
    ids = list(ids)[:max_len]
    return ids + [EOS_ID] * (max_len - len(ids))


def feat_reshape(feat: np.ndarray, max_len: int = 1200) -> np.ndarray:
    """utils.py:39-46."""
    h, w = feat.shape
    if h >= max_len:
        return feat[:max_len]
    out = np.zeros((max_len, w), dtype=feat.dtype)
    out[:h] = feat
    return out


def cal_descriptors(T, D):
    """utils.py:156-159 (prints 114 for (1200, 80), utils.py:193)."""
    def pool(x):
        return np.ceil(x / 2)
    return int(pool(pool(pool(pool(pool(T))))) * pool(pool(pool(pool(pool(D))))))


def to_categorical(y, num_classes):
    out = np.zeros((len(y), num_classes), dtype=np.float32)
    out[np.arange(len(y)), np.asarray(y, dtype=np.int64)] = 1.0
    return out


def synthetic_batch(cfg: SARConfig, B: int, seed: int = 2020, lengths: Optional[np.ndarray] = None,
                    label_len_range: Tuple[int, int] = (4, 20)) -> Tuple[Dict[str, np.ndarray], Dict[str, np.ndarray]]:
    """Seeded synthetic (inputs, targets) with the dict keys / dtypes / shapes of
    data_loader (utils.py:102-116).  x_data ~ U[0,1) (the range feat_norm guarantees); rows
    beyond `lengths[i]` are zero (utils.py:39-46 padding, no masking downstream -- Q4);
    CTC labels in [0, C-2] (blank = C-1), feasible for the encoder length."""
    rng = np.random.RandomState(seed)
    T, D, _ = cfg.input_shape
    x = rng.rand(B, T, D, 1).astype(np.float32)
    if lengths is not None:
        for i, n in enumerate(lengths):
            x[i, int(n):] = 0.0
    inputs: Dict[str, np.ndarray] = {"x_data": x}
    targets: Dict[str, np.ndarray] = {}
    acc = rng.randint(0, cfg.accent_classes, size=B)
    onehot = to_categorical(acc, cfg.accent_classes)
    if cfg.ar_enable:
        targets["y_accent"] = onehot
    if cfg.disc_enable:
        inputs["x_accent"] = onehot
        targets["y_disc"] = onehot
    if cfg.ctc_enable:
        S = cfg.plan().seq_len
        lo, hi = label_len_range
        labels = np.full((B, cfg.max_ctc_len), EOS_ID, dtype=np.float32)
        lab_len = np.zeros((B, 1), dtype=np.int32)
        for i in range(B):
            L = int(rng.randint(lo, hi + 1))
            L = max(1, min(L, cfg.max_ctc_len, S // 2))
            ids = rng.randint(0, cfg.bpe_classes - 1, size=L)
            # feasibility (tf.nn.ctc_loss raises otherwise): L + #adjacent repeats <= S
            while L + int(np.sum(ids[1:] == ids[:-1])) > S:
                ids = rng.randint(0, cfg.bpe_classes - 1, size=L)
            labels[i, :L] = ids
            lab_len[i, 0] = L
        inputs["x_ctc_label"] = labels
        inputs["x_ctc_in_len"] = np.full((B, 1), S, dtype=np.int32)      # utils.py:96: constant encoder_len
        inputs["x_ctc_out_len"] = lab_len
        targets["y_ctc_loss"] = np.zeros([B])
    return inputs, targets


def load(path):
    """utils.py:20-22."""
    import pickle
    with open(path, "rb") as f:
        return pickle.load(f)


def data_loader(lst, ctc_enable=False, ar_enable=False, disc_enable=False, data_dct=None, accent_dct=None, trans_dct=None,
                max_input_len=1200, max_ctc_len=72, encoder_len=100, accent_classes=8, bn=0, device="cuda"):
    """utils.py:71-117 with the per-batch numpy / sklearn work on the device: the un-padded feature matrices of the
    batch go up in ONE contiguous pinned upload, and sar_feat_batch_fwd (per-utterance MinMaxScaler + truncate / pad,
    utils.py:35-46) and sar_labels_pack_fwd (text_ids_norm, to_categorical, the two CTC length vectors) assemble the
    model inputs in HBM.  `data_dct[utt]` is a pickle path as in the reference (utils.py:91 `load(data_dct[utt])`) or
    the (frames, 80) array itself.  Returns (input_data, output_data) like the reference, as CUDA tensors with the
    reference's shapes / dtypes; `model.predict(input_data)` consumes them without a host round trip."""
    import torch
    from . import ops
    from ._shim import SarnetError
    dev = torch.device(device)
    feats = []
    for utt in lst:
        f = data_dct[utt]
        f = load(f) if isinstance(f, (str, bytes)) else f
        f = np.asarray(f, dtype=np.float32)
        if f.ndim != 2 or f.shape[0] < 1:
            raise ValueError("features of %s must be a non-empty (frames, dims) matrix" % (utt,))
        feats.append(f)
    D = feats[0].shape[1]
    offs = np.zeros(len(lst) + 1, dtype=np.int64)
    offs[1:] = np.cumsum([f.shape[0] for f in feats])
    flat = torch.empty((int(offs[-1]), D), dtype=torch.float32, pin_memory=True)
    np.concatenate(feats, axis=0, out=flat.numpy())
    x = ops.feat_batch(flat.to(dev, non_blocking=True), torch.from_numpy(offs).to(dev, non_blocking=True), int(max_input_len))
    input_data = {"x_data": x.unsqueeze(-1)}
    output_data = {}
    want_ctc = bool(ctc_enable and trans_dct)
    want_acc = bool(ar_enable and accent_dct)
    if want_ctc or want_acc:
        acc = torch.from_numpy(np.asarray([int(accent_dct[u]) for u in lst], dtype=np.int32)).to(dev) if want_acc else None
        tr = toff = None
        if want_ctc:
            ids = [np.asarray(trans_dct[u], dtype=np.int32).reshape(-1) for u in lst]
            to = np.zeros(len(lst) + 1, dtype=np.int64)
            to[1:] = np.cumsum([len(i) for i in ids])
            tr = torch.from_numpy(np.concatenate(ids + [np.zeros(1, np.int32)])).to(dev)
            toff = torch.from_numpy(to).to(dev)
        packed = ops.labels_pack(acc, accent_classes, tr, toff, max_ctc_len, encoder_len)
        if want_acc and int(packed["status"].item()) != 0:
            raise SarnetError("data_loader: accent id outside [0, %d) (to_categorical raises IndexError)" % accent_classes)
    if ctc_enable:
        if not want_ctc:
            raise ValueError("ctc_enable needs trans_dct")          # the reference builds ragged arrays and fails later
        input_data["x_ctc_in_len"] = packed["x_ctc_in_len"]
        input_data["x_ctc_out_len"] = packed["x_ctc_out_len"]
        input_data["x_ctc_label"] = packed["x_ctc_label"]
        output_data["y_ctc_loss"] = np.zeros([len(lst)])
    if ar_enable:
        output_data["y_accent"] = packed["x_accent"] if want_acc else None
    if disc_enable:
        input_data["x_accent"] = packed["x_accent"] if want_acc else None
        output_data["y_disc"] = input_data["x_accent"]
        if bn:
            output_data["y_disc_bn"] = input_data["x_accent"]
    return input_data, output_data


def pinned_like(arrays: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """Copies of `arrays` in PAGE-LOCKED host memory (numpy views of pinned torch tensors) -- what a loader's
    ring buffer would hand to model.predict(): such arrays are DMA'd to the device in place, pageable ones are
    first staged through an internal pinned buffer."""
    import torch
    out = {}
    for k, v in arrays.items():
        a = np.ascontiguousarray(v)
        t = torch.empty(a.shape, dtype=torch.from_numpy(a).dtype, pin_memory=True)
        t.copy_(torch.from_numpy(a))
        out[k] = t.numpy()          # the array keeps the tensor alive through its .base
    return out


def data_generator(lst, ctc_enable=False, ar_enable=False, disc_enable=False, batch_size=32, data_dct=None, accent_dct=None,
                   trans_dct=None, max_input_len=1200, max_ctc_len=72, encoder_len=100, accent_classes=8, bn=0,
                   device="cuda", seed=None):
    """utils.py:120-154: endless generator of (inputs, targets) batches -- shuffle `lst`, cut it into len(lst) // batch_size
    batches, data_loader() each, start over.  Batches are assembled on the device (data_loader above); `seed` makes the
    shuffles reproducible (the reference uses the global `random.shuffle`)."""
    import random
    rnd = random.Random(seed) if seed is not None else random
    lst = list(lst)
    n_batchs = len(lst) // batch_size
    if n_batchs < 1:
        raise ValueError("data_generator: %d utterances do not fill one batch of %d" % (len(lst), batch_size))
    while True:
        rnd.shuffle(lst)
        for i in range(n_batchs):
            sub = lst[i * batch_size:(i + 1) * batch_size]
            yield data_loader(sub, ctc_enable=ctc_enable, ar_enable=ar_enable, disc_enable=disc_enable, data_dct=data_dct,
                              accent_dct=accent_dct, trans_dct=trans_dct, max_input_len=max_input_len,
                              max_ctc_len=max_ctc_len, encoder_len=encoder_len, accent_classes=accent_classes, bn=bn,
                              device=device)


class RingBatch(dict):
    """An input dict handed out by PinnedRing: `ready` is a CUDA event recorded after the producer thread's device work
    for this batch (None for host batches); consumers order their stream behind it."""
    ready = None


class PinnedRing:
    """Background batch producer: the twin of Keras' GeneratorEnqueuer behind `fit_generator(generator,
    max_queue_size=20)` (train.py:38-44), which the reference relies on to keep the GPU fed.

    A worker thread drains `generator` (anything yielding an input dict or an (inputs, targets) tuple) and keeps up to
    `max_queue_size` batches ready:
      * host (numpy) arrays are staged into a RING of page-locked buffers -- allocated once per slot, key and shape, then
        reused -- so that model.predict / predict_generator DMA them to the device in place (no per-batch pinning, no
        staging memcpy on the consumer's thread); arrays that already live in pinned memory pass through untouched;
      * CUDA tensors (utils.data_loader assembles batches on the device) pass through; the worker runs the generator
        under a stream of its own and records an event per batch (RingBatch.ready) that the consumer waits on.
    Slot lifetime: a handed-out batch stays valid until `keep` further batches have been taken (predict_generator has
    finished the H2D of batch i before it asks for batch i + PIPE_DEPTH + 1)."""

    _END = object()

    def __init__(self, generator, max_queue_size: int = 20, keep: int = 8, device=None):
        import queue
        import threading
        self.gen = iter(generator)
        self.mq = max(1, int(max_queue_size))
        self.keep = max(1, int(keep))
        self.nslots = self.mq + self.keep
        self.slots = [dict() for _ in range(self.nslots)]
        self.q = queue.Queue(maxsize=self.mq)
        self.cv = threading.Condition()
        self.handed = 0
        self.stop = False
        self.error = None
        self.device = device
        self.staged_bytes = 0
        self.thread = threading.Thread(target=self._run, name="sarnet-pinned-ring", daemon=True)
        self.thread.start()

    # ---- producer thread
    def _stage(self, slot, k, v):
        import torch
        if isinstance(v, torch.Tensor):
            return v
        a = np.ascontiguousarray(v)
        if a.dtype == np.float64:
            a = a.astype(np.float32)
        t = torch.from_numpy(a)
        if t.is_pinned():
            return a
        pin = slot.get(k)
        if pin is None or pin.shape != t.shape or pin.dtype != t.dtype:
            pin = slot[k] = torch.empty(t.shape, dtype=t.dtype, pin_memory=torch.cuda.is_available())
        pin.copy_(t)
        self.staged_bytes += a.nbytes
        return pin.numpy()

    def _run(self):
        import contextlib
        import torch
        i = 0
        try:
            ctx = contextlib.nullcontext()
            stream = None
            if torch.cuda.is_available():
                dev = torch.device(self.device) if self.device is not None else None
                if dev is not None and dev.type == "cuda" and dev.index is not None:
                    torch.cuda.set_device(dev)
                stream = torch.cuda.Stream()
                ctx = torch.cuda.stream(stream)
            with ctx:
                for item in self.gen:
                    with self.cv:                 # slot i % nslots: its previous batch must be `keep` hand-outs old
                        while not self.stop and self.handed < i - self.mq + 1:
                            self.cv.wait(0.05)
                        if self.stop:
                            return
                    inputs, targets = (item if isinstance(item, tuple) else (item, None))
                    slot = self.slots[i % self.nslots]
                    out = RingBatch((k, self._stage(slot, k, v)) for k, v in inputs.items())
                    if stream is not None and any(isinstance(v, torch.Tensor) and v.is_cuda for v in out.values()):
                        out.ready = torch.cuda.Event()
                        out.ready.record(stream)
                    while not self.stop:
                        try:
                            self.q.put((out, targets), timeout=0.05)
                            break
                        except Exception:
                            continue
                    if self.stop:
                        return
                    i += 1
        except BaseException as e:                # surfaced on the consumer's thread by __next__
            self.error = e
        finally:
            while not self.stop:
                try:
                    self.q.put(self._END, timeout=0.05)
                    break
                except Exception:
                    continue

    # ---- consumer
    def __iter__(self):
        return self

    def __next__(self):
        import queue
        while True:                               # never block forever on a producer that died
            try:
                item = self.q.get(timeout=0.2)
                break
            except queue.Empty:
                if not self.thread.is_alive() and self.q.empty():
                    if self.error is not None:
                        raise self.error
                    raise StopIteration
        if item is self._END:
            self.q.put(self._END)
            if self.error is not None:
                raise self.error
            raise StopIteration
        with self.cv:
            self.handed += 1
            self.cv.notify_all()
        out, targets = item
        return out if targets is None else (out, targets)

    def close(self):
        self.stop = True
        with self.cv:
            self.cv.notify_all()
        self.thread.join(timeout=2.0)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
