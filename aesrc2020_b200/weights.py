"""Canonical-name weight container for the SAR-Net hot path.

Layouts are the reference's Keras layouts: conv kernels HWIO,
Dense kernels (in, out), CuDNNGRU kernel (Din, 3u) / recurrent_kernel (u, 3u) /
bias (6u,) in gate order z|r|h, BN gamma/beta/moving_mean/moving_variance,
LayerNormalization gamma/beta, VladPooling `centers` (K+G, D) (VLAD.py:16-19),
margin-head `W` (D, n_classes) (losses.py:22-26).

Model-level layer names are the reference's explicit Keras names (model.py:252-322,
95-106).  ResNet layers have no stable Keras names (resnet.py passes no `name=`), so
they get canonical names resnet/s{stage}b{block}/{bn1,conv1,bn2,conv2,short}.

Files are `.npz`.  h5py is not available in this image, so the HDF5 container of a Keras `.h5`
checkpoint is not parsed here; its NAMING half is: `keras_weight_names(cfg)` gives the
Keras weight name (`conv2d_7/kernel:0`, `CRNN/forward_cu_dnngru_1/bias:0`, ...) of every canonical weight,
and `from_keras_named` / `SARModel.load_weights` accept an `.npz` keyed by those names, which one line
on the TF side produces.  The tensor layouts need no conversion (see above).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Tuple

import numpy as np

from .config import SARConfig


def weight_shapes(cfg: SARConfig) -> "OrderedDict[str, Tuple[int, ...]]":
    """Every weight tensor of the forward graph SAR_Net(cfg) builds, in creation order."""
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()

    def bn(name, c):
        for k in ("gamma", "beta", "moving_mean", "moving_variance"):
            s["%s/%s" % (name, k)] = (c,)

    def dense(name, i, o, bias=True):
        s[name + "/kernel"] = (i, o)
        if bias:
            s[name + "/bias"] = (o,)

    def ln(name, c):
        s[name + "/gamma"] = (c,)
        s[name + "/beta"] = (c,)

    def bigru(name, din, u):
        for d in ("forward", "backward"):
            s["%s/%s/kernel" % (name, d)] = (din, 3 * u)
            s["%s/%s/recurrent_kernel" % (name, d)] = (u, 3 * u)
            s["%s/%s/bias" % (name, d)] = (6 * u,)

    def conv(c):
        s[c.name + "/kernel"] = (c.kh, c.kw, c.cin, c.cout)
        s[c.name + "/bias"] = (c.cout,)

    plan = cfg.plan()
    conv(plan.stem)
    bn("resnet/stem_bn", plan.stem.cout)
    for b in plan.blocks:
        if b.conv1.pre_bn:
            bn(b.conv1.pre_bn, b.conv1.cin)
        conv(b.conv1)
        bn(b.conv2.pre_bn, b.conv2.cin)
        conv(b.conv2)
        if b.short:
            conv(b.short)
    bn(plan.final_bn, plan.cout)

    H = cfg.hidden_dim
    dense("CNN_LIN", plan.cout, H)
    ln("CNN_LIN_LN", H)
    bigru("CRNN", H, H)
    ln("CRNN_LN", 2 * H)
    if cfg.ctc_enable:
        bigru("CTC_BIGRU", 2 * H, H)
        ln("CTC_BIGRU_LN", 2 * H)
        dense("CTC_DS", 2 * H, H)
        ln("CTC_DS_LN", H)
        dense("ctc_pred", H, cfg.bpe_classes)
    if cfg.ar_enable:
        dense("AR_DS", 2 * H, H)
        ln("AR_DS_LN", H)
        if cfg.mto == "bigru":
            bigru("AR_MERGE", H, H)
        elif cfg.mto in ("vlad", "gvlad"):
            kg = cfg.vlad_clusters + (cfg.ghost_clusters if cfg.mto == "gvlad" else 0)
            s[cfg.mto + "_center_assignment/kernel"] = (1, 1, H, kg)
            s[cfg.mto + "_center_assignment/bias"] = (kg,)
            s[cfg.mto + "_pool/centers"] = (kg, H)
        d_int = cfg.integration_dim()
        bn("AR_BN1", d_int)
        dense("AR_EMBEDDING", d_int, H)
        bn("AR_BN2", H)
        dense("AR_CF_DS1", H, 64)
        dense("AR_CF_DS2", 64, 64)
        dense("y_accent", 64, cfg.accent_classes)

        def disc(name, din):
            if cfg.metric_loss in ("sphereface", "cosface", "arcface"):
                s[name + "/W"] = (din, cfg.accent_classes)
            elif cfg.metric_loss in ("softmax", "circleloss"):
                s[name + "/kernel"] = (din, cfg.accent_classes)

        if cfg.disc_enable:
            disc("y_disc", H)
        if cfg.disc_enable and cfg.bn_dim:
            dense("AR_BN_DS", H, 64)
            bn("AR_BN3", 64)
            dense("bottleneck", 64, cfg.bn_dim)
            bn("AR_BN4", cfg.bn_dim)
            disc("y_disc_bn", cfg.bn_dim)
    return s


def init_weights(cfg: SARConfig, seed: int = 1234, degenerate: bool = False) -> Dict[str, np.ndarray]:
    """Seeded synthetic weights.  With degenerate=False (default) BN statistics, affine
    terms and biases are random so that no term of the forward is hidden by a Keras
    zero/one initialiser; scales follow the reference's initialisers (he_normal convs and
    Dense, glorot/orthogonal-scale GRU) so activations stay O(1) through the network."""
    rng = np.random.RandomState(seed)
    out: Dict[str, np.ndarray] = {}
    # A trained network's BN moving statistics track its activations.  The un-normalised
    # residual stream of a pre-activation ResNet gains roughly one unit of variance per block, so
    # the statistics of block j's leading BN (and of the final BN) are centred on that estimate --
    # otherwise the synthetic stream grows ~2x per block and every downstream tanh saturates.
    stream_var = {}
    for j, b in enumerate(cfg.plan().blocks):
        if b.conv1.pre_bn:
            stream_var[b.conv1.pre_bn] = 1.0 + 0.6 * j
    stream_var[cfg.plan().final_bn] = 1.0 + 0.6 * len(cfg.plan().blocks)
    for name, shp in weight_shapes(cfg).items():
        leaf = name.rsplit("/", 1)[1]
        base_var = stream_var.get(name.rsplit("/", 1)[0], 1.0)
        if leaf == "kernel" and len(shp) == 4:
            fan_in = shp[0] * shp[1] * shp[2]
            w = rng.randn(*shp) * np.sqrt(2.0 / fan_in)
        elif leaf == "kernel":
            if "/forward/" in name or "/backward/" in name:
                lim = np.sqrt(6.0 / (shp[0] + shp[1]))
                w = rng.uniform(-lim, lim, size=shp)
            else:
                w = rng.randn(*shp) * np.sqrt(2.0 / shp[0])
        elif leaf == "recurrent_kernel":
            w = rng.randn(*shp) / np.sqrt(shp[0])
        elif leaf == "W":
            lim = np.sqrt(6.0 / (shp[0] + shp[1]))
            w = rng.uniform(-lim, lim, size=shp)
        elif leaf == "centers":
            w = rng.randn(*shp) / np.sqrt(shp[1])
        elif leaf == "bias":
            w = np.zeros(shp) if degenerate else rng.randn(*shp) * 0.1
        elif leaf == "gamma":
            w = np.ones(shp) if degenerate else rng.uniform(0.7, 1.3, size=shp)
        elif leaf == "beta":
            w = np.zeros(shp) if degenerate else rng.randn(*shp) * 0.1
        elif leaf == "moving_mean":
            w = np.zeros(shp) if degenerate else rng.randn(*shp) * 0.1 * np.sqrt(base_var)
        elif leaf == "moving_variance":
            w = np.ones(shp) if degenerate else rng.uniform(0.6, 1.4, size=shp) * base_var
        else:
            raise KeyError(name)
        out[name] = np.ascontiguousarray(w, dtype=np.float32)
    return out


def _is_h5_path(path) -> bool:
    return isinstance(path, str) and path.lower().endswith((".h5", ".hdf5", ".keras.h5"))


def save_weights(path: str, weights: Dict[str, np.ndarray], cfg: "SARConfig" = None, full_model: bool = False) -> None:
    """`model.save_weights(path)` / `model.save(path)` (train.py:35, model.py:416-417).  A `.h5` / `.hdf5` path writes
    the Keras HDF5 layout (h5lite.write_keras_weights: `layer_names` / `weight_names` attributes, one dataset per
    weight under Keras' own weight names -- the file the reference's `load_weights(..., by_name=True)` reads); any
    other path writes an `.npz` of the canonical names AT EXACTLY THAT PATH (np.savez would append '.npz')."""
    if _is_h5_path(path):
        if cfg is None:
            raise ValueError("save_weights('%s'): the Keras HDF5 layout needs the model config (weight names)" % path)
        from . import h5lite
        knames = keras_weight_names(cfg)
        layers: "OrderedDict[str, OrderedDict[str, np.ndarray]]" = OrderedDict()
        for canon, kname in knames.items():
            if canon not in weights:
                continue
            layers.setdefault(kname.split("/")[0], OrderedDict())[kname] = np.asarray(weights[canon], dtype=np.float32)
        h5lite.write_keras_weights(path, layers, full_model=full_model)
        return
    with open(path, "wb") as f:
        np.savez(f, **{k.replace("/", "|"): np.asarray(v) for k, v in weights.items()})


def load_weights(path: str) -> Dict[str, np.ndarray]:
    """-> {name: array}.  HDF5 files (sniffed by signature, whatever the extension) give KERAS weight names
    (`conv2d_1/kernel:0`; map them with from_keras_named), `.npz` files whatever names they were saved under."""
    import os
    if not os.path.exists(path) and os.path.exists(path + ".npz"):
        path = path + ".npz"                               # files written by np.savez(path_without_extension)
    with open(path, "rb") as f:
        magic = f.read(8)
    if magic == b"\x89HDF\r\n\x1a\n":
        from . import h5lite
        return dict(h5lite.read_keras_weights(path))
    with np.load(path) as z:
        return {k.replace("|", "/"): z[k] for k in z.files}


def keras_weight_names(cfg: SARConfig) -> "OrderedDict[str, str]":
    """canonical weight name -> the name Keras gives that weight in a FRESH session (`layer.weights[i].name`, the
    `weight_names` attribute of a `.h5` checkpoint), for the graph model.SAR_Net(cfg) builds.

    Layers the reference names keep their names (model.py:252-322).  The unnamed ones are called
    `<snake_case(class)>_<uid>` with the uid counted per class from 1 in CREATION order (keras Layer.__init__): the ResNet's
    Conv2D / BatchNormalization (resnet.py passes no `name=`; per block: bn_a, conv_a, bn_b, conv_b, then the shortcut
    conv -- resnet.py:111-123,82) and the CuDNNGRU inside every Bidirectional, whose two copies Bidirectional renames
    `forward_<name>` / `backward_<name>`.  weight_shapes(cfg) is in creation order, so walking it reproduces the uids;
    pinned by tests/golden/keras_names.json (the reference's own model.py executed on minikeras)."""
    out: "OrderedDict[str, str]" = OrderedDict()
    uid = {"conv2d": 0, "batch_normalization": 0, "cu_dnngru": 0}
    layer_of: Dict[str, str] = {}
    for name in weight_shapes(cfg):
        parts = name.split("/")
        weight = parts[-1]
        if len(parts) >= 3 and parts[-2] in ("forward", "backward"):
            owner = "/".join(parts[:-2])                                  # the Bidirectional layer (named)
            if owner not in layer_of:
                uid["cu_dnngru"] += 1
                layer_of[owner] = "cu_dnngru_%d" % uid["cu_dnngru"]
            out[name] = "%s/%s_%s/%s:0" % (owner, parts[-2], layer_of[owner], weight)
            continue
        layer = "/".join(parts[:-1])
        if layer.startswith("resnet/"):
            if layer not in layer_of:
                kind = "batch_normalization" if weight in ("gamma", "beta", "moving_mean", "moving_variance") else "conv2d"
                uid[kind] += 1
                layer_of[layer] = "%s_%d" % (kind, uid[kind])
            out[name] = "%s/%s:0" % (layer_of[layer], weight)
        else:
            out[name] = "%s/%s:0" % (layer, weight)
    return out


def from_keras_named(cfg: SARConfig, named) -> Dict[str, np.ndarray]:
    """Canonical weights from a mapping keyed by Keras weight names (`conv2d_1/kernel:0`; the HDF5 spelling
    `conv2d_1/conv2d_1/kernel:0` and names without the `:0` suffix are accepted too).  Missing or mis-shaped entries are
    skipped like load_weights(by_name=True, skip_mismatch=True) does (model.py:181-183); layouts are Keras' own."""
    def norm(k):
        k = k[:-2] if k.endswith(":0") else k
        p = k.split("/")
        if len(p) >= 3 and p[0] == p[1]:
            p = p[1:]
        return "/".join(p)
    src = {norm(k): v for k, v in named.items()}
    shapes = weight_shapes(cfg)
    out = {}
    for canon, kname in keras_weight_names(cfg).items():
        v = src.get(norm(kname))
        if v is not None and tuple(np.shape(v)) == tuple(shapes[canon]):
            out[canon] = np.ascontiguousarray(v, dtype=np.float32)
    return out
