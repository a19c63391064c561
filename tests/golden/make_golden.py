"""Generate tests/golden/*.npz by executing the reference's OWN source files.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference, which does not
exist on the GPU box):

    python tests/golden/make_golden.py            # rewrites tests/golden/sarnet_*.npz

How: `minikeras` installs eager numpy stand-ins for the `keras`, `keras_layer_normalization`
and `tensorflow` modules the reference imports; then `/root/reference/{model,resnet,VLAD,
losses}.py` are imported UNMODIFIED and `model.SAR_Net(...)` is called exactly as train.py:10-37
calls it.  Building the graph IS running it (eager), so the outputs are whatever the reference's
own wiring and hand-written arithmetic (VladPooling.call, SphereFace/CosFace/ArcFace.call,
circle_loss, the loss / loss_weights dicts of model.py:344-367) produce.

Weights are not stored (thin-ResNet34 is ~50 MB): they are `aesrc2020_b200.weights.init_weights(cfg,
seed)` -- a pure numpy RandomState function -- handed to the reference graph in layer CREATION
order with shape and name checks (minikeras.reset(queue=...)).  A test regenerates the same
weights from (cfg, seed); `weights_l1` in each fixture guards against RNG drift.

What these fixtures pin: graph wiring and the reference-authored arithmetic.  The third-party primitives underneath
(minikeras) are asserted against PyTorch's own conv / batch-norm / pool / GRU / layer-norm / CTC kernels by
check_minikeras.assert_primitives_match_torch() before anything is written (and again by tests/test_golden.py).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import minikeras as mk                                   # noqa: E402  (installs the stand-in modules)

sys.path.insert(0, REF)
import model as ref_model                                # noqa: E402  /root/reference/model.py
import losses as ref_losses                              # noqa: E402  /root/reference/losses.py
import VLAD as ref_vlad                                  # noqa: E402  /root/reference/VLAD.py

from aesrc2020_b200.config import SARConfig              # noqa: E402
from aesrc2020_b200 import weights as W, utils as us     # noqa: E402

assert ref_model.__file__.startswith(REF) and ref_losses.__file__.startswith(REF) and ref_vlad.__file__.startswith(REF)

# name -> (T, B, seed, lengths, SAR_Net kwargs).  cfg1..cfg5 are BASELINE.json's configs at
# fixture-sized T/B; the rest cover the remaining heads, merge modes and the bottleneck branch.
CASES = {
    "cfg1_res18_avg_softmax": (300, 1, 11, None, dict(res_type="res18", res_filters=64, mto="avg")),
    "cfg2_gvlad_arcface": (200, 2, 12, None, dict(disc_enable=True, res_type="res34", res_filters=32, mto="gvlad",
                                                  vlad_clusters=64, ghost_clusters=8, metric_loss="arcface", margin=0.3)),
    "cfg3_ctc_circle_bigru": (260, 3, 13, [100, 183, 260], dict(ctc_enable=True, disc_enable=True, res_type="res34",
                                                                 res_filters=32, mto="bigru", metric_loss="circleloss",
                                                                 margin=0.2)),
    "cfg4_vlad_cosface": (200, 2, 14, None, dict(disc_enable=True, res_type="res34", res_filters=32, mto="vlad",
                                                 vlad_clusters=64, metric_loss="cosface", margin=0.3)),
    "cfg5_gvlad_circle_ctc": (200, 2, 15, None, dict(ctc_enable=True, disc_enable=True, res_type="res34", res_filters=32,
                                                     mto="gvlad", vlad_clusters=64, ghost_clusters=8,
                                                     metric_loss="circleloss", margin=0.2)),
    "sphereface_bn": (200, 2, 16, None, dict(ctc_enable=True, disc_enable=True, res_type="res34", res_filters=32,
                                             mto="gvlad", vlad_clusters=8, ghost_clusters=2, bn_dim=32,
                                             metric_loss="sphereface", margin=1.35)),
    "softmax_head_res18_thin": (200, 2, 17, None, dict(disc_enable=True, res_type="res18", res_filters=32, mto="avg",
                                                       metric_loss="softmax")),
    "ctc_only_res18": (200, 2, 18, None, dict(ctc_enable=True, ar_enable=False, res_type="res18", res_filters=16)),
}

INTERMEDIATES = {"CNN2SEQ": "resnet_seq", "CNN_LIN_LN": "cnn_lin", "CRNN_LN": "crnn", "AR_DS_LN": "ar_ds",
                 "AR_BN2": "embedding", "ctc_pred": "ctc_pred", "vlad_pool": "integration", "gvlad_pool": "integration",
                 "AR_MERGE": "integration"}


def keras_categorical_crossentropy(y, p):
    """[KERAS-SEMANTICS] K.categorical_crossentropy on probabilities (third-party; restated)."""
    p = p / p.sum(-1, keepdims=True)
    p = np.clip(p, 1e-7, 1 - 1e-7)
    return -(y * np.log(p)).sum(-1)


def run_case(name):
    T, B, seed, lengths, kw = CASES[name]
    cfg = SARConfig(input_shape=(T, 80, 1), **kw)
    weights = W.init_weights(cfg, seed=seed)
    x, y = us.synthetic_batch(cfg, B, seed=seed + 100, lengths=lengths, label_len_range=(2, 6))
    if cfg.bn_dim and cfg.disc_enable:
        y["y_disc_bn"] = y["y_disc"]            # utils.py:113-114
    mk.reset(queue=list(weights.items()))
    mk.FEED.update(x)
    model, train_model = ref_model.SAR_Net((T, 80, 1), **kw)          # the reference's own graph
    assert train_model is model
    assert not mk.QUEUE, "unused weights: %r" % [n for n, _ in mk.QUEUE][:5]
    assert [n for n, _ in mk.USED] == list(weights), "creation order differs from weight_shapes()"

    out = {}
    for oname, t in zip(cfg.output_names(), model.outputs):
        out[oname] = t.v
    for lname, key in INTERMEDIATES.items():
        try:
            v = model.get_layer(lname).output.v
            out[key] = v if key == "embedding" else v.astype(np.float32)      # big ones at fp32 (fixture size)
        except ValueError:
            pass
    # losses exactly as compile() received them (model.py:344-367): string => Keras CE, else the lambda
    comp = model.compiled
    total = 0.0
    for oname in cfg.output_names():
        fn = comp["loss"][oname]
        tgt = y[oname].reshape(B, -1) if oname != "y_ctc_loss" else y[oname]
        if fn == "categorical_crossentropy":
            per = keras_categorical_crossentropy(np.asarray(tgt, np.float64), out[oname])
        else:
            per = fn(mk.KT(tgt), mk.KT(out[oname])).v
        per = np.asarray(per).reshape(B, -1).mean(-1)
        out["loss/" + oname] = np.float64(per.mean())
        out["loss_weight/" + oname] = np.float64(comp["loss_weights"][oname])
        total += comp["loss_weights"][oname] * per.mean()
        if oname in comp["metrics"]:
            out["acc/" + oname] = np.float64((out[oname].argmax(-1) == y[oname].argmax(-1)).mean())
    out["loss/total"] = np.float64(total)
    out["weights_l1"] = np.float64(sum(float(np.abs(v.astype(np.float64)).sum()) for v in weights.values()))
    meta = dict(T=T, B=B, seed=seed, input_seed=seed + 100, lengths=lengths, kwargs=kw, label_len_range=[2, 6])
    out["meta"] = np.array(json.dumps(meta))
    for k, v in x.items():                      # x_data is regenerated from input_seed; its checksum is kept
        if k == "x_data":
            out["in_l1/x_data"] = np.float64(np.abs(v.astype(np.float64)).sum())
        else:
            out["in/" + k] = v
    for k, v in y.items():
        out["tgt/" + k] = v
    return out


def layer_cases():
    """VladPooling / margin heads / circle_loss called directly at odd sizes (no SAR_Net)."""
    rng = np.random.RandomState(77)
    out = {}
    # *256: hidden_dim-wide descriptors (the width the CUDA kernel is built for) at odd K / G / S
    for mode, K, G, S, D in (("vlad", 5, 0, 7, 12), ("gvlad", 6, 3, 9, 16), ("vlad256", 5, 0, 7, 256), ("gvlad256", 6, 3, 9, 256)):
        feat = rng.randn(3, 1, S, D)
        score = rng.randn(3, 1, S, K + G) * 2
        cen = rng.randn(K + G, D).astype(np.float32)
        mk.reset(queue=[("p/centers", cen)])
        lay = ref_vlad.VladPooling(mode=mode.replace("256", ""), k_centers=K, g_centers=G, name="p")
        res = lay([mk.KT(feat), mk.KT(score)])
        assert lay.compute_output_shape([feat.shape, score.shape]) == (3, K * D)
        out.update({"%s/feat" % mode: feat, "%s/score" % mode: score, "%s/centers" % mode: cen, "%s/out" % mode: res.v})
    x = rng.randn(5, 24)
    yl = rng.randint(0, 8, size=5)
    y = np.eye(8)[yl]
    out["face/x"], out["face/y"] = x, y
    for cls, m in (("SphereFace", 1.35), ("CosFace", 0.35), ("ArcFace", 0.5), ("ArcFace", 0.3)):
        w = rng.uniform(-0.4, 0.4, size=(24, 8)).astype(np.float32)
        mk.reset(queue=[("h/W", w)])
        lay = getattr(ref_losses, cls)(n_classes=8, m=m, name="h")
        res = lay([mk.KT(x), mk.KT(y)])
        out["face/%s_%g/W" % (cls, m)] = w
        out["face/%s_%g/out" % (cls, m)] = res.v
    cos = np.tanh(rng.randn(6, 8))
    yc = np.eye(8)[rng.randint(0, 8, size=6)]
    out["circle/cos"], out["circle/y"] = cos, yc
    for g, m in ((256, 0.25), (256, 0.2), (64, 0.4)):
        out["circle/g%d_m%g" % (g, m)] = ref_losses.circle_loss(mk.KT(yc), mk.KT(cos), gamma=g, margin=m).v
    return out


def main():
    only = sys.argv[1:]
    # the stand-in primitives are checked against PyTorch's own kernels BEFORE any fixture is written, so the fixtures
    # are tied to an implementation that is independent of minikeras and of the oracle (check_minikeras.py)
    from check_minikeras import assert_primitives_match_torch
    assert_primitives_match_torch(verbose=True)
    for name in CASES:
        if only and name not in only:
            continue
        res = run_case(name)
        path = os.path.join(HERE, "sarnet_%s.npz" % name)
        np.savez_compressed(path, **{k.replace("/", "|"): v for k, v in res.items()})
        print("%-28s %7.1f KB  outputs: %s" % (name, os.path.getsize(path) / 1024,
                                               {k: np.asarray(v).shape for k, v in res.items() if k.startswith("y_")}))
    if not only or "layers" in only:
        res = layer_cases()
        path = os.path.join(HERE, "layers.npz")
        np.savez_compressed(path, **{k.replace("/", "|"): v for k, v in res.items()})
        print("layers  %.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
