"""Mirror of the reference's model.py call surface, executing on B200 through the C ABI.

Kept verbatim from the reference: `SAR_Net(...)` keyword set and `(model, train_model)`
return (model.py:204-224,371); the returned model's `predict / get_layer / load_weights /
save_weights / summary`; helper names `build, compile, integration, vlad, disc_loss,
ctc_module, ctc_lambda_func, sub_model, ctc_pred` and the layer factories `SQUEEZE, EXPAND,
BN, LN, DS, BIGRU, DP` (model.py:23-53).  `compile` keeps `lr` for `fit_generator` / `train_on_batch` (training.HeadTrainer:
Keras' Adam(lr, decay=2e-4), model.py:197) and wires the data-parallel wrapper for gpus > 1.

Tensors at this surface: numpy arrays (host, as in Keras) or CUDA torch tensors.  predict()
with host arrays does the H2D copy from pinned memory, runs the kernels, and copies the
outputs back; with device tensors it is zero-copy.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from . import ops, weights as _weights, utils as _utils
from . import losses as ls
from . import VLAD as vd
from .config import SARConfig
from .engine import SARNetEngine, fold_bn
from ._shim import SarnetError

ArrayLike = Union[np.ndarray, torch.Tensor]
AUTO_LANES = 1       # default micro-batch lanes for batches >= 32 (SARModel.lanes = 0)


# =========================
#         Layers            (model.py:23-53) -- eager, device tensors in/out
# =========================
class _Lambda:
    def __init__(self, fn, name=None):
        self.fn, self.name = fn, name

    def __call__(self, x):
        return self.fn(x)


def SQUEEZE(axis=3, name=None):
    return _Lambda(lambda x: x.squeeze(axis), name=name)


def EXPAND(axis=3, name=None):
    return _Lambda(lambda x: x.unsqueeze(axis), name=name)


def DP(rate, name=None):
    return _Lambda(lambda x: x, name=name)          # Dropout is the identity at inference


class _Layer:
    def __init__(self, name=None):
        self.name = name
        self.weights: Dict[str, np.ndarray] = {}
        self._dev: Dict[str, torch.Tensor] = {}

    def set_weights_dict(self, w: Dict[str, np.ndarray]):
        self.weights = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in w.items()}
        self._dev = {}

    def get_weights(self):
        return list(self.weights.values())

    def d(self, key, device, value=None):
        if key not in self._dev or self._dev[key].device != device:
            src = self.weights[key] if value is None else value
            self._dev[key] = torch.from_numpy(np.ascontiguousarray(src, dtype=np.float32)).to(device)
        return self._dev[key]


class _BN(_Layer):
    def __call__(self, x):
        if not self.weights:
            c = x.shape[-1]
            self.set_weights_dict({"gamma": np.ones(c), "beta": np.zeros(c), "moving_mean": np.zeros(c),
                                   "moving_variance": np.ones(c)})
        s, t = fold_bn({"bn/" + k: v for k, v in self.weights.items()}, "bn")
        return ops.affine_relu(x.contiguous(), self.d("s", x.device, s), self.d("t", x.device, t), relu=False)


def BN(name=None):
    return _BN(name)


class _LN(_Layer):
    def __call__(self, x):
        if not self.weights:
            c = x.shape[-1]
            self.set_weights_dict({"gamma": np.ones(c), "beta": np.zeros(c)})
        return ops.layernorm(x.contiguous(), self.d("gamma", x.device), self.d("beta", x.device))


def LN(name=None):
    return _LN(name)


class _DS(_Layer):
    def __init__(self, hidden, activation, use_bias=True, name=None):
        super().__init__(name)
        self.hidden, self.activation, self.use_bias = hidden, activation, use_bias

    def __call__(self, x):
        if not self.weights:
            din = x.shape[-1]
            rng = np.random.RandomState(1234)
            w = {"kernel": rng.randn(din, self.hidden) * np.sqrt(2.0 / din)}
            if self.use_bias:
                w["bias"] = np.zeros(self.hidden)
            self.set_weights_dict(w)
        act = self.activation if self.activation in ("relu", "tanh") else None
        y = ops.dense(x.contiguous(), self.d("kernel", x.device),
                      self.d("bias", x.device) if self.use_bias else None, act=act)
        if self.activation == "softmax":
            y = ops.softmax_rows(y)                    # any width (ctc_pred: DS(1000, 'softmax'), model.py:268)
        return y


def DS(hidden, activation, rgr=None, use_bias=True, name=None):
    return _DS(hidden, activation, use_bias=use_bias, name=name)


class _BIGRU(_Layer):
    def __init__(self, hidden, seq=True, name=None):
        super().__init__(name)
        self.hidden, self.seq = hidden, seq
        self.nb = 0          # utterances per cluster of the recurrence kernel (0 = its own choice, 16 or 32)

    def __call__(self, x):
        u = self.hidden
        if not self.weights:
            din = x.shape[-1]
            rng = np.random.RandomState(1234)
            w = {}
            for d in ("forward", "backward"):
                lim = np.sqrt(6.0 / (din + 3 * u))
                w[d + "/kernel"] = rng.uniform(-lim, lim, (din, 3 * u))
                w[d + "/recurrent_kernel"] = rng.randn(u, 3 * u) / np.sqrt(u)
                w[d + "/bias"] = np.zeros(6 * u)
            self.set_weights_dict(w)
        w = self.weights
        kcat = np.concatenate([w["forward/kernel"], w["backward/kernel"]], axis=1)
        bcat = np.concatenate([w["forward/bias"][:3 * u], w["backward/bias"][:3 * u]])
        rec = np.stack([w["forward/recurrent_kernel"], w["backward/recurrent_kernel"]])
        rb = np.stack([w["forward/bias"][3 * u:], w["backward/bias"][3 * u:]])
        B, S, _ = x.shape
        xp = ops.dense(x.contiguous(), self.d("kcat", x.device, kcat), self.d("bcat", x.device, bcat))
        return ops.bigru(xp.reshape(B, S, 2, 3 * u), self.d("rec", x.device, rec), self.d("rb", x.device, rb),
                         seq=self.seq, nb=self.nb)


def BIGRU(hidden, seq=True, rgr=None, name=None):
    return _BIGRU(hidden, seq=seq, name=name)


# =========================
#     ctc constructors      (model.py:62-73)
# =========================
def ctc_lambda_func(args):
    """K.ctc_batch_cost(labels, y_pred, input_length, label_length) on device.  `y_pred` are the
    ctc_pred softmax PROBABILITIES as in the reference; the kernel takes logits, and
    softmax(log p) == p, so log-probabilities are passed (zeros map to -inf safely)."""
    y_pred, labels, input_length, label_length = args
    loss, status, _ = ops.ctc(torch.log(y_pred).contiguous(), labels, input_length, label_length)
    if bool((status == 1).any()):
        raise SarnetError("Not enough time for target transition sequence (infeasible CTC label)")
    return loss.reshape(-1, 1)


def ctc_module(ctc_pred, max_label_len):
    raise SarnetError("ctc_module builds Keras Input placeholders (model.py:66-73); the device path takes the "
                      "x_ctc_label/x_ctc_in_len/x_ctc_out_len tensors through SAR_Net(...).predict")


# =========================
#          NetVLAD          (model.py:82-109)
# =========================
def vlad(x, aggregation, vlad_clusters, ghost_clusters, weights: Optional[Dict[str, np.ndarray]] = None):
    """x (B,1,S,D) -> (B, K*D): fused 1x1 assignment conv + VladPooling (one launch)."""
    if aggregation not in ("vlad", "gvlad"):
        return x
    G = ghost_clusters if aggregation == "gvlad" else 0
    D = x.shape[-1]
    kg = vlad_clusters + G
    if weights is None:
        rng = np.random.RandomState(1234)
        weights = {aggregation + "_center_assignment/kernel": rng.randn(1, 1, D, kg) / np.sqrt(D),
                   aggregation + "_center_assignment/bias": np.zeros(kg),
                   aggregation + "_pool/centers": rng.randn(kg, D) / np.sqrt(D)}
    dev = x.device
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
    return ops.vlad(x.reshape(x.shape[0], -1, D).contiguous(),
                    t(weights[aggregation + "_center_assignment/kernel"].reshape(D, kg)),
                    t(weights[aggregation + "_center_assignment/bias"]),
                    t(weights[aggregation + "_pool/centers"]), vlad_clusters, G)


# =========================
#         AR Module         (model.py:118-167)
# =========================
def integration(x, hidden_dim=256, mto="avg", vlad_clusters=8, ghost_clusters=2, weights=None):
    if mto == "avg":
        return ops.avgpool(x.contiguous())
    if mto == "bigru":
        layer = BIGRU(hidden_dim, seq=False, name="AR_MERGE")
        if weights:
            layer.set_weights_dict({k[len("AR_MERGE/"):]: v for k, v in weights.items() if k.startswith("AR_MERGE/")})
        return layer(x)
    if mto in ("vlad", "gvlad"):
        return vlad(EXPAND(axis=1)(x), aggregation=mto, vlad_clusters=vlad_clusters, ghost_clusters=ghost_clusters,
                    weights=weights)
    print("Please specify avg/bigru/vlad/gvlad ..")
    raise SystemExit(1)


def disc_loss(x, accent_label, accent_classes, loss, margin, name, W: Optional[np.ndarray] = None):
    """model.py:142-167 on device tensors; returns the layer output (probabilities, or raw
    cosines for 'circleloss'), or None for an unknown loss (as the reference does)."""
    if loss not in ops.HEAD or loss in (None, "none", "circle_raw"):
        return None
    if W is None:
        lim = np.sqrt(6.0 / (x.shape[1] + accent_classes))
        W = np.random.RandomState(1234).uniform(-lim, lim, (x.shape[1], accent_classes))
    wd = torch.from_numpy(np.ascontiguousarray(W, dtype=np.float32)).to(x.device)
    y = None if accent_label is None else accent_label.contiguous().float()
    return ops.head(None, None, emb_d=x.contiguous(), wd=wd, onehot=y, n_classes=accent_classes,
                    head_kind=loss, margin=margin)["y_disc"]


# =========================
#           Model           (model.py:175-371)
# =========================
class LayerView:
    """What `model.get_layer(name)` returns: named access to one layer's weights."""

    def __init__(self, model: "SARModel", name: str):
        self.model, self.name = model, name
        keys = [k for k in model.weights if k == name or k.startswith(name + "/")]
        if not keys and name not in model.config.input_names() + model.config.output_names():
            raise ValueError("No such layer: " + name)
        self.keys = keys

    def get_weights(self):
        return [self.model.weights[k] for k in self.keys]

    def set_weights(self, ws):
        for k, v in zip(self.keys, ws):
            if tuple(v.shape) != tuple(self.model.weights[k].shape):
                raise ValueError("shape mismatch for %s" % k)
            self.model.weights[k] = np.ascontiguousarray(v, dtype=np.float32)
        self.model._engine = None


class SARModel:
    """The object SAR_Net returns (stands in for keras.models.Model on this path)."""

    def __init__(self, config: SARConfig, weights: Dict[str, np.ndarray], name=None, device="cuda",
                 outputs: Optional[List[str]] = None, gpus: int = 1):
        self.config, self.weights, self.name = config, weights, name or "model"
        self.device = device
        self.gpus = gpus
        self._outputs = outputs or config.output_names()
        self._engine: Optional[SARNetEngine] = None
        self._pinned: Dict[str, torch.Tensor] = {}
        self._pin_events: Dict[str, torch.cuda.Event] = {}
        self._pinned_out: Optional[torch.Tensor] = None
        self.use_graph = True          # replay a CUDA graph of the step (captured once per batch shape)
        # concurrent micro-batch lanes per step (engine.forward_lanes); 0 = automatic (see _lanes_for)
        self.lanes = int(os.environ.get("SAR_LANES", "0"))
        self._pipe = None              # predict_generator's staging slots (see _pipe_state)
        self.lr = 0.01                 # compile(): Adam(lr, decay=2e-4), model.py:197
        self._trainer = None           # training.HeadTrainer over the whole model, built by the first training call

    # -- Keras-like surface
    @property
    def input_names(self):
        return self.config.input_names()

    @property
    def output_names(self):
        return list(self._outputs)

    def engine(self) -> SARNetEngine:
        if self._engine is None:
            self._engine = SARNetEngine(self.config, self.weights, self.device)
        return self._engine

    def summary(self, print_fn=print):
        tot = 0
        print_fn('Model: "%s"' % self.name)
        for k, v in self.weights.items():
            print_fn("%-48s %-22s %10d" % (k, tuple(v.shape), v.size))
            tot += v.size
        print_fn("Total params: %d" % tot)

    def get_layer(self, name):
        return LayerView(self, name)

    def save_weights(self, path):
        """Keras `Model.save_weights`: `*.h5` -> Keras HDF5 layout (readable by the reference's load_weights), else npz."""
        _weights.save_weights(path, self.weights, cfg=self.config)

    def save(self, path):
        """Keras `Model.save` (train.py:35 `model.save("%s/%03d.h5")`): the weights under `/model_weights`.  No
        optimizer state or model_config is written (forward-only path)."""
        _weights.save_weights(path, self.weights, cfg=self.config, full_model=True)

    def load_weights(self, path, by_name=True, skip_mismatch=True):
        """model.py:181-183 semantics: load by name, silently skip shape mismatches.  An `.npz` keyed by KERAS weight
        names (`conv2d_1/kernel:0`, ...: `np.savez(path, **{w.name: v ...})` on the TF side) is mapped
        to the canonical names first (weights.keras_weight_names)."""
        loaded = _weights.load_weights(path)                  # .h5 (Keras HDF5, parsed by h5lite) or .npz
        if any(k.endswith(":0") for k in loaded):
            loaded = _weights.from_keras_named(self.config, loaded)
        for k, v in loaded.items():
            if k in self.weights and tuple(self.weights[k].shape) == tuple(v.shape):
                self.weights[k] = np.ascontiguousarray(v, dtype=np.float32)
            elif k in self.weights and not skip_mismatch:
                raise ValueError("shape mismatch for %s" % k)
        self._engine = None

    # -- execution
    def _to_host_tensor(self, key: str, v) -> torch.Tensor:
        """numpy input -> host tensor of the dtype the kernels take, in PAGE-LOCKED memory: arrays that already
        live in pinned memory (utils.pinned_like -- what a loader's ring buffer would be) are used in place, any
        other array is staged through a cached pinned buffer (guarded by an event: the previous asynchronous
        H2D out of that buffer must have finished before it is overwritten)."""
        a = np.ascontiguousarray(v)
        want = np.int32 if key in ("x_ctc_in_len", "x_ctc_out_len") else np.float32
        a = a.astype(want, copy=False)
        th = torch.from_numpy(a)
        if a.nbytes >= (1 << 16) and th.is_pinned():
            return th
        pin = self._pinned.get(key)
        if pin is None or pin.shape != th.shape or pin.dtype != th.dtype:
            pin = torch.empty(th.shape, dtype=th.dtype, pin_memory=True)
            self._pinned[key] = pin
        ev = self._pin_events.get(key)
        if ev is not None:
            ev.synchronize()
        pin.copy_(th)
        return pin

    def _mark_h2d(self, keys):
        for k in keys:
            if k in self._pinned:
                ev = self._pin_events.get(k)
                if ev is None:
                    ev = self._pin_events[k] = torch.cuda.Event()
                ev.record()

    def _to_device(self, key: str, v: ArrayLike) -> torch.Tensor:
        if isinstance(v, torch.Tensor):
            t = v.to(self.device, non_blocking=True)
        else:
            t = self._to_host_tensor(key, v).to(self.device, non_blocking=True)
            self._mark_h2d([key])
        if key in ("x_ctc_in_len", "x_ctc_out_len"):
            return t.to(torch.int32)
        return t.float() if t.dtype != torch.float32 else t

    def _as_dict(self, x) -> Dict[str, ArrayLike]:
        if isinstance(x, dict):
            return x
        if isinstance(x, (list, tuple)):
            return dict(zip(self.config.input_names(), x))
        return {"x_data": x}

    @property
    def decode_only(self) -> bool:
        """A sub-model whose only output is `ctc_pred` (sub_model(model, 'x_data', 'ctc_pred'), model.py:380-383):
        encoder + ASR branch, x_data is its only input."""
        return list(self._outputs) == ["ctc_pred"]

    def required_inputs(self) -> List[str]:
        return ["x_data"] if self.decode_only else self.config.input_names()

    def _lanes_for(self, B: int) -> int:
        """Micro-batch lanes of a graphed step: batches too small to fill the GPU with one kernel at a time are
        split so that independent kernels of the lanes overlap (measured: profiles/r1_lanes.md)."""
        if self.lanes > 0:
            return min(self.lanes, B)
        return AUTO_LANES if B >= 32 else 1

    def forward_device(self, x, want_intermediates=False, graph: bool = None) -> Dict[str, torch.Tensor]:
        """One forward on the device.  `graph` (default: self.use_graph) replays a captured CUDA graph
        of the step; the returned tensors are then the graph's static outputs."""
        xd = self._as_dict(x)
        missing = [k for k in self.required_inputs() if k not in xd]
        if missing:
            raise ValueError("missing model inputs: %s" % missing)
        use_graph = self.use_graph if graph is None else graph
        if self.decode_only:
            # decode-only inference: no labels, no CTC loss, no accent branch (the reference predicts ctc_pred on
            # the x_data -> ctc_pred sub-model, model.py:380-389)
            xin = {"x_data": self._to_device("x_data", xd["x_data"])}
            eng = self.engine()
            out = (eng.forward_graphed(xin, tag="decode", decode_only=True) if use_graph
                   else eng.forward(xin, decode_only=True))
            out = dict(out)
            out["ctc_pred"] = ops.softmax_rows(out["__ctc_logits"], classes=self.config.bpe_classes)
            return out
        if use_graph and not want_intermediates:
            # host inputs go from page-locked memory STRAIGHT into the captured graph's static input buffers
            # (one asynchronous H2D per input, no intermediate device tensor)
            src = {}
            for k, v in xd.items():
                if isinstance(v, torch.Tensor):
                    src[k] = self._to_device(k, v)
                else:
                    src[k] = self._to_host_tensor(k, v)
            out = self.engine().forward_lanes(src, self._lanes_for(len(xd["x_data"])))
            self._mark_h2d([k for k, v in xd.items() if not isinstance(v, torch.Tensor)])
        else:
            dev_in = {k: self._to_device(k, v) for k, v in xd.items()}
            out = self.engine().forward(dev_in, want_intermediates=want_intermediates)
        if self.config.ctc_enable and bool((out["ctc_status"] != 0).any()):
            # tf.nn.ctc_loss raises on infeasible / out-of-range labels
            raise SarnetError("CTC: infeasible or out-of-range label sequence in batch "
                              "(Not enough time for target transition sequence)")
        return out

    def predict(self, x, batch_size=32, verbose=0):
        """Keras Model.predict: list of host arrays in output order (single array if one output).
        Device-tensor inputs return device tensors (no host round trip)."""
        xd = self._as_dict(x)
        n = len(xd["x_data"])
        on_device = isinstance(xd["x_data"], torch.Tensor) and xd["x_data"].is_cuda
        if not on_device and n > batch_size and self.use_graph and not self.decode_only:
            # several chunks: the pipelined path (copies and consecutive chunks overlap), same outputs
            return self.predict_generator({k: v[b0:b0 + batch_size] for k, v in xd.items()} for b0 in range(0, n, batch_size))
        chunks: List[List] = [[] for _ in self._outputs]
        for b0 in range(0, n, batch_size):
            sl = {k: v[b0:b0 + batch_size] for k, v in xd.items()}
            out = self.forward_device(sl)
            if on_device:
                for i, name in enumerate(self._outputs):
                    chunks[i].append(out[name].clone())
                continue
            # all outputs of the chunk -> one pinned buffer, asynchronously, then ONE stream synchronisation
            outs = [out[name] for name in self._outputs]
            tot = sum(o.numel() for o in outs)
            if self._pinned_out is None or self._pinned_out.numel() < tot:
                self._pinned_out = torch.empty((max(tot, 4096),), dtype=torch.float32, pin_memory=True)
            off = 0
            for o in outs:
                self._pinned_out[off:off + o.numel()].view(o.shape).copy_(o, non_blocking=True)
                off += o.numel()
            torch.cuda.current_stream().synchronize()
            off = 0
            for i, o in enumerate(outs):
                chunks[i].append(self._finite(self._pinned_out[off:off + o.numel()].view(o.shape).numpy().copy()))
                off += o.numel()
        cat = (lambda c: torch.cat(c, 0)) if on_device else (lambda c: np.concatenate(c, 0))
        res = [cat(c) for c in chunks]
        return res[0] if len(res) == 1 else res

    # -- pipelined prediction over a batch generator
    PIPE_DEPTH = 3

    def _pipe_state(self):
        if self._pipe is None:
            dev = torch.device(self.device)
            D = self.PIPE_DEPTH
            self._pipe = {
                "copy": torch.cuda.Stream(device=dev),
                "h2d": [torch.cuda.Event() for _ in range(D)], "free": [None] * D,
                "done": [torch.cuda.Event() for _ in range(D)],
                "stage": [dict() for _ in range(D)], "pin_in": [dict() for _ in range(D)], "pin_out": [None] * D,
            }
        return self._pipe

    def _submit(self, x, slot: int):
        """Enqueue one batch without waiting for it: H2D on the copy stream into staging slot `slot`, the graphed
        step and the D2H of its outputs on the compute stream.  Returns the handle _collect() waits on."""
        ps = self._pipe_state()
        xd = self._as_dict(x)
        missing = [k for k in self.required_inputs() if k not in xd]
        if missing:
            raise ValueError("missing model inputs: %s" % missing)
        B = len(xd["x_data"])
        cur = torch.cuda.current_stream()
        stage, pins = ps["stage"][slot], ps["pin_in"][slot]
        with torch.cuda.stream(ps["copy"]):
            if ps["free"][slot] is not None:
                ps["copy"].wait_event(ps["free"][slot])      # the step that last read this slot has consumed it
            for k, v in xd.items():
                want = torch.int32 if k in ("x_ctc_in_len", "x_ctc_out_len") else torch.float32
                if isinstance(v, torch.Tensor):
                    src = v
                    if v.is_cuda and getattr(xd, "ready", None) is not None:
                        ps["copy"].wait_event(xd.ready)      # produced on a PinnedRing worker's stream
                        v.record_stream(ps["copy"])
                    elif v.is_cuda:
                        # produced by kernels on the CALLER's stream (utils.data_loader / fbank_batch hand their
                        # outputs over without a host round trip): the staging copy must run after them, and the
                        # caching allocator must not recycle the tensor while the copy stream still reads it
                        ps["copy"].wait_stream(cur)
                        v.record_stream(ps["copy"])
                else:
                    a = np.ascontiguousarray(v).astype(np.int32 if want == torch.int32 else np.float32, copy=False)
                    src = torch.from_numpy(a)
                    if not src.is_pinned():                  # pageable: through this slot's own pinned buffer (the
                        pin = pins.get(k)                    # wait on `free` above also covers its previous H2D)
                        if pin is None or pin.shape != src.shape or pin.dtype != src.dtype:
                            pin = pins[k] = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
                        if ps["free"][slot] is not None:
                            ps["free"][slot].synchronize()
                        pin.copy_(src)
                        src = pin
                dst = stage.get(k)
                if dst is None or dst.shape != src.shape:
                    dst = stage[k] = torch.empty(src.shape, dtype=want, device=self.device)
                dst.copy_(src, non_blocking=True)
            ps["h2d"][slot].record(ps["copy"])
        lanes = self._lanes_for(B)
        if self.use_graph and lanes == 1:
            # the whole step runs on the slot's own stream / graph / activation buffers: consecutive batches overlap
            # on the device as well (engine.forward_slot), not only their copies
            comp = self.engine().slot_stream(slot)
            comp.wait_stream(cur)
            comp.wait_event(ps["h2d"][slot])
            out, _ = self.engine().forward_slot({k: stage[k] for k in xd}, slot)
        else:
            comp = cur
            cur.wait_event(ps["h2d"][slot])
            if self.use_graph:
                out = self.engine().forward_lanes({k: stage[k] for k in xd}, lanes)
            else:
                out = self.engine().forward({k: stage[k] for k in xd})
        if ps["free"][slot] is None:
            ps["free"][slot] = torch.cuda.Event()
        ps["free"][slot].record(comp)
        outs = [out[name] for name in self._outputs]
        if self.config.ctc_enable:
            outs.append(out["ctc_status"])
        tot = sum(o.numel() for o in outs)
        pin = ps["pin_out"][slot]
        if pin is None or pin.numel() < tot:
            pin = ps["pin_out"][slot] = torch.empty((max(tot, 4096),), dtype=torch.float32, pin_memory=True)
        off = 0
        views = []
        with torch.cuda.stream(comp):
            for o in outs:
                dst = pin[off:off + o.numel()]
                dst = dst.view(torch.int32) if o.dtype == torch.int32 else dst
                dst.view(o.shape).copy_(o, non_blocking=True)
                views.append(dst.view(o.shape))
                off += o.numel()
            ps["done"][slot].record(comp)
        return slot, views

    def _collect(self, handle) -> List[np.ndarray]:
        slot, views = handle
        self._pipe_state()["done"][slot].synchronize()
        if self.config.ctc_enable and bool((views[-1] != 0).any()):
            raise SarnetError("CTC: infeasible or out-of-range label sequence in batch "
                              "(Not enough time for target transition sequence)")
        return [self._finite(v.numpy().copy()) for v in views[:len(self._outputs)]]

    @staticmethod
    def _finite(a: np.ndarray) -> np.ndarray:
        """The tensor-core operand planes are fp16 hi/lo pairs: an activation with |x| >= 65504 becomes inf in the hi plane
        and NaN one layer later, and reaches the outputs as NaN.  The hot kernels carry no range check; the (tiny) host
        copy of the outputs does, so a saturated batch fails loudly instead of returning garbage."""
        if a.size <= (1 << 20) and a.dtype.kind == "f" and not np.isfinite(a).all():
            raise SarnetError("non-finite model output: the inputs are non-finite, or an activation left the fp16 range of the "
                              "tensor-core operand planes (|x| >= 65504); rescale the features (utils.feat_norm yields [0, 1])")
        return a

    def predict_generator(self, generator, steps=None, max_queue_size=10, workers=1, use_multiprocessing=False, verbose=0,
                          prefetch: bool = False):
        """Keras `Model.predict_generator` (the forward-only twin of the `fit_generator(generator, max_queue_size=20)`
        loop the reference trains with, train.py:38-44): `generator` yields one batch per step -- an input dict /
        list as utils.data_loader builds it (utils.py:102-116), or the (inputs, targets) tuple data_generator
        yields.  Batches are pipelined PIPE_DEPTH deep, each in flight on a stream, CUDA graph and buffer set of its own:
        while step i runs, step i+1's inputs are DMA'd from pinned host memory by the copy engine, its first kernels
        fill the SMs that step i's latency-bound tail (Bi-GRU, VLAD, head) leaves idle, and step i-1's outputs are
        read back.  Returns the outputs concatenated over steps, as predict() does.
        `prefetch=True` runs the generator on a worker thread like Keras does (`workers`, `max_queue_size`): a
        utils.PinnedRing keeps max_queue_size batches staged in page-locked memory ahead of the device."""
        from collections import deque
        if self.decode_only:                      # x_data -> ctc_pred sub-model: plain per-batch predict
            outs = [self.predict(x[0] if isinstance(x, tuple) else x, batch_size=1 << 30)
                    for i, x in enumerate(generator) if steps is None or i < steps]
            return np.concatenate(outs, 0) if outs else np.zeros((0,), np.float32)
        ring = None
        if workers and workers >= 1 and prefetch and not isinstance(generator, _utils.PinnedRing):
            # Keras runs the generator on a worker thread and queues max_queue_size batches; here that thread also
            # stages host arrays into a ring of pinned buffers, so this thread only launches
            ring = generator = _utils.PinnedRing(generator, max_queue_size=max_queue_size, keep=self.PIPE_DEPTH + 2,
                                                 device=torch.device(self.device))
        it = iter(generator)
        pending = deque()
        chunks: List[List] = [[] for _ in self._outputs]

        def drain():
            for i, a in enumerate(self._collect(pending.popleft())):
                chunks[i].append(a)
        n = 0
        try:
            while steps is None or n < steps:
                try:
                    x = next(it)
                except StopIteration:
                    break
                if isinstance(x, tuple):
                    x = x[0]
                if len(pending) >= self.PIPE_DEPTH:
                    drain()
                pending.append(self._submit(x, n % self.PIPE_DEPTH))
                n += 1
            while pending:
                drain()
        finally:
            if ring is not None:
                ring.close()
        res = [np.concatenate(c, 0) for c in chunks] if n else [np.zeros((0,), np.float32) for _ in chunks]
        return res[0] if len(res) == 1 else res

    # -- training surface (train.py:38-44): the whole model, nothing frozen
    def trainer(self, group=None):
        """The training.HeadTrainer behind train_on_batch / fit_generator: every layer trains (ResNet with batch-statistic
        BatchNormalization, CRNN, accent branch, and the CTC branch when ctc_enable), Adam(self.lr, decay=2e-4)."""
        if self._trainer is None:
            cfg = self.config
            if not cfg.ar_enable or cfg.bn_dim:
                raise SarnetError("training is built for ar_enable=True and bn_dim=0 (got ar_enable=%s, bn_dim=%s)"
                                  % (cfg.ar_enable, cfg.bn_dim))
            from .training import HeadTrainer
            self._trainer = HeadTrainer(self, lr=self.lr, group=group, train_resnet=True, train_ctc=bool(cfg.ctc_enable))
        return self._trainer

    def train_on_batch(self, x, y=None):
        """Keras Model.train_on_batch: one optimisation step; returns {'loss', 'loss_accent', 'loss_disc'[, 'loss_ctc']}.
        The inference weights follow at the end of fit_generator (or trainer().sync_to_model())."""
        return self.trainer().train_on_batch(x, y)

    def fit_generator(self, generator, steps_per_epoch, epochs=1, verbose=1, callbacks=None, validation_data=None,
                      max_queue_size=10, workers=1, use_multiprocessing=False, initial_epoch=0, **kw):
        """Keras Model.fit_generator as train.py:38-44 calls it: `epochs - initial_epoch` epochs of `steps_per_epoch` batches
        of (inputs, targets) from `generator` (utils.data_generator).  After every epoch the trained weights are written
        back into the model (so a callback's `model.save(...)`, train.py:31-35, stores them), `validation_data` =
        (inputs, targets) is evaluated with the inference engine (`val_*` entries), and every object in `callbacks` that
        has `on_epoch_end(epoch, logs)` is called (duck-typed: Keras itself is not a dependency).  A callback may stop the
        run by setting `model.stop_training = True` (what EarlyStopping does).  Returns the list of per-epoch logs."""
        tr = self.trainer()
        it = iter(generator)
        history = []
        self.stop_training = False
        for epoch in range(int(initial_epoch), int(epochs)):
            acc = []
            for _ in range(int(steps_per_epoch)):
                item = next(it)
                xb, yb = item if isinstance(item, tuple) else (item, None)
                acc.append(tr.train_on_batch(xb, yb))
            logs = {k: float(np.mean([a[k] for a in acc])) for k in acc[0]}
            tr.sync_to_model()
            if validation_data is not None:
                vx, vy = validation_data[0], (validation_data[1] if len(validation_data) > 1 else None)
                for k, v in self.evaluate(vx, vy).items():
                    logs["val_" + k] = float(v)
            history.append(logs)
            if verbose:
                print("Epoch %d/%d - %s" % (epoch + 1, epochs, " - ".join("%s: %.4f" % kv for kv in logs.items())))
            for cb in (callbacks or []):
                if hasattr(cb, "on_epoch_end"):
                    cb.on_epoch_end(epoch, logs)
            if self.stop_training:
                break
        return history

    def evaluate(self, x, y: Optional[Dict[str, ArrayLike]] = None, batch_size=32, group=None):
        """Keras-style evaluate: batch-mean losses, accuracies and the weighted total
        (model.py:344-367).  With torch.distributed initialised the 8-float loss vector is
        all-reduced (SUM) across ranks -- the single collective of this path."""
        from . import dist as _dist
        xd = dict(self._as_dict(x))
        if y is not None and "y_accent" in y and "x_accent" not in xd:
            xd["y_true"] = y["y_accent"]
        n = len(xd["x_data"])
        total = None
        for b0 in range(0, n, batch_size):
            sl = {k: v[b0:b0 + batch_size] for k, v in xd.items()}
            if "y_true" in sl:
                sl["y_true"] = self._to_device("y_true", sl["y_true"])
            vec = self.forward_device(sl)["loss_vector"]
            total = vec.clone() if total is None else total + vec     # vec may be a graph-static buffer
        total = _dist.all_reduce_loss_vector(total, group)
        return _dist.loss_vector_to_metrics(total.cpu().numpy(), self.config)


def build(inputs, outputs, raw=None, name="model"):
    """model.py:175-184.  `inputs` = SARConfig, `outputs` = weights dict on this path."""
    model = SARModel(inputs, outputs, name=name)
    if raw:
        print("===== init weights from:%s =====" % raw)
        model.load_weights(raw, by_name=True, skip_mismatch=True)
    return model


def compile(model, gpus, lr=None, loss=None, loss_weights=None, metrics=None):
    """model.py:187-201: Adam(lr, decay=2e-4) is what fit_generator / train_on_batch use (training.HeadTrainer).  gpus>1
    returns the data-parallel view (one process per GPU under torchrun; batch sharded by rank; loss vector and, in
    training, the gradients all-reduced)."""
    if lr is not None:
        model.lr = float(lr)
    if gpus > 1:
        from .dist import DataParallelModel
        return DataParallelModel(model, gpus)
    return model


def SAR_Net(input_shape, ctc_enable=False, ar_enable=True, disc_enable=False, res_type="res18",
            res_filters=64, hidden_dim=256, bn_dim=0, bpe_classes=1000, accent_classes=8, max_ctc_len=72,
            mto=None, vlad_clusters=8, ghost_clusters=2, metric_loss="cosface", margin=0.3, raw_model=None,
            lr=0.01, gpus=1, mode="train", name=None, weights: Optional[Dict[str, np.ndarray]] = None,
            seed: int = 1234, device="cuda"):
    """Signature of model.py:204-224 (+ `weights`/`seed`/`device` extensions, keyword-only in
    practice).  Returns (model, train_model); `train_model is model` when gpus == 1."""
    if mode != "train":
        # model.py:229-232 builds Input([None, D, 1]) here, which _shortcut cannot handle
        # (resnet.py:75 divides None) -- the reference cannot build this graph either.
        raise SarnetError("mode != 'train' (variable-length Input) cannot be built by the reference either "
                          "(resnet.py:75 on a None dimension); use fixed-T zero-padded inputs")
    cfg = SARConfig(tuple(int(v) for v in input_shape), ctc_enable, ar_enable, disc_enable, res_type, res_filters,
                    hidden_dim, bn_dim, bpe_classes, accent_classes, max_ctc_len, mto, vlad_clusters,
                    ghost_clusters, metric_loss, margin)
    if res_type not in ("res18", "res34", "res50", "res101", "res152"):
        print("======= ERROR: please specify cnn in res-[18,34,50,101,152] ======")
        raise ValueError(res_type)
    cfg.plan()                                   # raises NotImplementedError for res50/101/152 (Q1)
    if ar_enable and mto not in ("avg", "bigru", "vlad", "gvlad"):
        print("Please specify avg/bigru/vlad/gvlad ..")      # model.py:136-138
        raise SystemExit(1)
    if hidden_dim != 256:
        raise SarnetError("this build's Bi-GRU kernel supports hidden_dim=256 only")
    w = weights if weights is not None else _weights.init_weights(cfg, seed)
    expect = _weights.weight_shapes(cfg)
    for k, shp in expect.items():
        if k not in w or tuple(w[k].shape) != tuple(shp):
            raise ValueError("weights: %s missing or wrong shape (want %s)" % (k, shp))
    model = SARModel(cfg, {k: np.ascontiguousarray(w[k], dtype=np.float32) for k in expect}, name=name,
                     device=device, gpus=gpus)
    if raw_model:
        print("===== init weights from:%s =====" % raw_model)
        model.load_weights(raw_model, by_name=True, skip_mismatch=True)
    train_model = compile(model, gpus, lr=lr, loss_weights=cfg.loss_weights())
    print(cfg.loss_weights())                    # model.py:370
    return model, train_model


# ======================
#         OTHER            (model.py:380-389)
# ======================
def sub_model(model: SARModel, input_name, output_name):
    """Model(inputs=get_layer(input_name).input, outputs=get_layer(output_name).output)."""
    valid = model.config.output_names() + ["embedding", "y_accent_logits", "y_disc_logits"]
    if model.config.ctc_enable:
        valid.append("ctc_pred")           # x_data -> posteriors (B, S, bpe_classes): decode-only, needs no labels
    if output_name == "AR_BN2":
        output_name = "embedding"
    if output_name not in valid:
        raise ValueError("sub_model: output %s is not exposed (have %s)" % (output_name, valid))
    sm = SARModel(model.config, model.weights, name=model.name, device=model.device, outputs=[output_name])
    sm._engine = model._engine
    return sm


def ctc_pred(model: SARModel, x, batch_size, input_len):
    """model.py:385-389: `K.ctc_decode(model.predict(x), [input_len]*n, greedy=True)[0][0]` -- the greedy CTC decode of
    the ctc_pred posteriors (first maximum per frame, repeats merged, blank = C-1 dropped), dense (n, Lmax) int64 padded
    with -1.  On device: the graphed step leaves the pre-softmax ctc_pred logits in HBM and sar_ctc_greedy_fwd decodes
    them (argmax of the softmax = argmax of the logits); only the decoded ids travel to the host."""
    if not model.config.ctc_enable:
        raise SarnetError("ctc_pred needs a model built with ctc_enable=True")
    xd = model._as_dict(x)
    n = len(xd["x_data"])
    S = model.config.plan().seq_len
    T = max(0, min(int(input_len), S))
    rows = []
    # x_data is the only input the decode needs: the step runs as the x_data -> ctc_pred sub-model whatever model
    # object was passed (label inputs, if present, are ignored -- dummy labels must not trip the infeasible-CTC check)
    dm = model if model.decode_only else sub_model(model, "x_data", "ctc_pred")
    for b0 in range(0, n, batch_size):
        o = dm.forward_device({"x_data": xd["x_data"][b0:b0 + batch_size]})
        dec, dec_len = ops.ctc_greedy(o["__ctc_logits"], fixed_len=T, classes=model.config.bpe_classes)
        dec, dec_len = dec.cpu().numpy(), dec_len.cpu().numpy()
        rows += [dec[i, :dec_len[i]] for i in range(len(dec))]
    L = max([len(r) for r in rows] + [1])
    out = -np.ones((len(rows), L), dtype=np.int64)       # K.ctc_decode pads with -1
    for i, r in enumerate(rows):
        out[i, :len(r)] = r
    return out
