"""Mirror of the parts of the reference's utils.py that sit on the hot path's boundary:
the input dict contract (utils.py:102-116), label padding (utils.py:57-63) and the shape
known-answer `cal_descriptors` (utils.py:156-159).  Pickle/scp file handling and the
shuffling generator are data preparation and stay out of scope; `synthetic_batch` provides
the seeded synthetic inputs the benchmarks and tests use instead.
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
