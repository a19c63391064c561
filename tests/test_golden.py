"""Golden-vector parity: tests/golden/*.npz hold outputs of the reference's OWN source
(/root/reference model.py / resnet.py / VLAD.py / losses.py) executed on the minikeras eager
stand-in by tests/golden/make_golden.py.  Nothing here reads /root/reference.

 * not gpu: the float64 oracle must reproduce every fixture (pins the oracle's graph wiring and
   its restatement of VladPooling / margin heads / circle loss / loss weights);
 * gpu: the CUDA path through the C ABI must reproduce the same fixtures within north_star's
   1e-3 relative tolerance (1e-6 floor).
"""
import glob
import json
import os

import numpy as np
import pytest
import torch

from helpers import rel_err, norm_err, REL_TOL
from oracle import sarnet_oracle as O
from aesrc2020_b200.config import SARConfig
from aesrc2020_b200 import weights as W, utils as us

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[len("sarnet_"):-4] for p in glob.glob(os.path.join(GOLD, "sarnet_*.npz")))


def load_case(name):
    with np.load(os.path.join(GOLD, "sarnet_%s.npz" % name)) as z:
        g = {k.replace("|", "/"): z[k] for k in z.files}
    meta = json.loads(str(g["meta"]))
    cfg = SARConfig(input_shape=(meta["T"], 80, 1), **meta["kwargs"])
    weights = W.init_weights(cfg, seed=meta["seed"])
    l1 = sum(float(np.abs(v.astype(np.float64)).sum()) for v in weights.values())
    assert abs(l1 - float(g["weights_l1"])) <= 1e-9 * l1, "init_weights drifted from the fixture generator"
    x, y = us.synthetic_batch(cfg, meta["B"], seed=meta["input_seed"], lengths=meta["lengths"],
                              label_len_range=tuple(meta["label_len_range"]))
    assert abs(float(np.abs(x["x_data"].astype(np.float64)).sum()) - float(g["in_l1/x_data"])) < 1e-6
    for k in x:
        if k != "x_data":
            assert np.array_equal(x[k], g["in/" + k]), k
    return cfg, meta, weights, x, y, g


def test_fixture_inventory():
    assert len(CASES) == 8 and os.path.exists(os.path.join(GOLD, "layers.npz"))
    assert {"cfg1_res18_avg_softmax", "cfg2_gvlad_arcface", "cfg3_ctc_circle_bigru", "cfg4_vlad_cosface",
            "cfg5_gvlad_circle_ctc"} <= set(CASES)


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference_graph(name):
    cfg, meta, weights, x, y, g = load_case(name)
    ref = O.sar_net_forward(weights, x, **cfg.model_kwargs(), return_intermediates=True)
    for n in cfg.output_names():
        assert tuple(ref[n].shape) == g[n].shape
        assert rel_err(ref[n], g[n], floor=1e-12) < 1e-8, n
    B = meta["B"]
    inter = {"resnet_seq": ref["resnet"].reshape(B, -1, ref["resnet"].shape[-1]), "cnn_lin": ref["cnn_lin"],
             "crnn": ref["crnn"]}
    for k in ("ar_ds", "integration", "embedding", "ctc_pred"):
        if k in ref:
            inter[k] = ref[k]
    for k, v in inter.items():
        assert k in g, k
        assert norm_err(v, g[k]) < (1e-9 if g[k].dtype == np.float64 else 1e-6), k
    # compile()'d losses, loss weights and accuracies as the reference's dicts define them
    if cfg.ar_enable:
        tgt = torch.as_tensor(y["y_accent"], dtype=torch.float64)
        res = O.sar_net_losses(ref, tgt, ctc_enable=cfg.ctc_enable, ar_enable=cfg.ar_enable,
                               disc_enable=cfg.disc_enable, bn_dim=cfg.bn_dim, metric_loss=cfg.metric_loss,
                               margin=cfg.margin)
        for n in cfg.output_names():
            assert abs(float(res[n + "_loss"]) - float(g["loss/" + n])) <= 1e-8 * max(1.0, abs(float(g["loss/" + n]))), n
            if "acc/" + n in g:
                assert float(res[n + "_acc"]) == float(g["acc/" + n]) if n + "_acc" in res else True
        assert abs(float(res["loss"]) - float(g["loss/total"])) <= 1e-8 * max(1.0, abs(float(g["loss/total"])))
    lw = O.loss_weights(cfg.ctc_enable, cfg.ar_enable, cfg.disc_enable, cfg.bn_dim)
    assert lw == cfg.loss_weights()
    for n in cfg.output_names():
        assert lw[n] == float(g["loss_weight/" + n]), n


def _layers():
    with np.load(os.path.join(GOLD, "layers.npz")) as z:
        return {k.replace("|", "/"): z[k] for k in z.files}


def test_oracle_layers_match_reference_source():
    g = _layers()
    t = lambda a: torch.as_tensor(a, dtype=torch.float64)
    for mode, K in (("vlad", 5), ("gvlad", 6), ("vlad256", 5), ("gvlad256", 6)):
        got = O.vlad_pooling(t(g[mode + "/feat"]), t(g[mode + "/score"]), t(g[mode + "/centers"]),
                             mode.replace("256", ""), K)
        assert rel_err(got, g[mode + "/out"], floor=1e-12) < 1e-9, mode
    for key, kind, m in (("SphereFace_1.35", "sphereface", 1.35), ("CosFace_0.35", "cosface", 0.35),
                         ("ArcFace_0.5", "arcface", 0.5), ("ArcFace_0.3", "arcface", 0.3)):
        logits = O.face_logits(t(g["face/x"]), t(g["face/%s/W" % key]), t(g["face/y"]), kind, m)
        assert rel_err(torch.softmax(logits, -1), g["face/%s/out" % key], floor=1e-12) < 1e-9, key
    for gm, m in ((256, 0.25), (256, 0.2), (64, 0.4)):
        got = O.circle_loss(t(g["circle/y"]), t(g["circle/cos"]), float(gm), m)
        assert rel_err(got, g["circle/g%d_m%g" % (gm, m)], floor=1e-9) < 1e-9


# ------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_path_reproduces_reference_graph(cuda_device, name):
    from aesrc2020_b200 import model as mdl
    cfg, meta, weights, x, y, g = load_case(name)
    model, _ = mdl.SAR_Net((meta["T"], 80, 1), **meta["kwargs"], weights=weights)
    outs = model.predict(x, batch_size=meta["B"])
    outs = outs if isinstance(outs, list) else [outs]
    for n, got in zip(cfg.output_names(), outs):
        assert got.shape == g[n].shape
        assert rel_err(got, g[n]) < REL_TOL, n
    dev_out = model.forward_device(x, want_intermediates=True)
    for k in ("cnn_lin", "crnn", "ar_ds", "integration", "embedding", "ctc_pred"):
        if k in g and k in dev_out:
            assert norm_err(dev_out[k], g[k]) < 1e-4, k
    if cfg.ar_enable:
        got = model.evaluate(x, y, batch_size=meta["B"])
        for n in cfg.output_names():
            want = float(g["loss/" + n])
            assert abs(got[n + "_loss"] - want) <= REL_TOL * max(abs(want), 1e-3), (n, got[n + "_loss"], want)
        want = float(g["loss/total"])
        assert abs(got["loss"] - want) <= REL_TOL * max(abs(want), 1e-3)


@pytest.mark.gpu
def test_cuda_layers_match_reference_source(cuda_device):
    """VladPooling / margin-head call surfaces of the mirror vs the reference source's outputs."""
    from aesrc2020_b200 import VLAD as vd, losses as ls
    g = _layers()
    c = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float32)).cuda()
    for mode, K, G in (("vlad256", 5, 0), ("gvlad256", 6, 3)):      # vlad.cu is built for D == hidden_dim == 256
        lay = vd.VladPooling(mode=mode.replace("256", ""), k_centers=K, g_centers=G, name="p")
        feat, score = c(g[mode + "/feat"]), c(g[mode + "/score"])
        lay.build([tuple(feat.shape), tuple(score.shape)])
        lay.set_weights([g[mode + "/centers"]])
        got = lay([feat, score])
        assert tuple(got.shape) == lay.compute_output_shape([tuple(feat.shape), tuple(score.shape)])
        assert rel_err(got, g[mode + "/out"], floor=1e-4) < REL_TOL, mode
    for key, cls, m in (("SphereFace_1.35", ls.SphereFace, 1.35), ("CosFace_0.35", ls.CosFace, 0.35),
                        ("ArcFace_0.5", ls.ArcFace, 0.5), ("ArcFace_0.3", ls.ArcFace, 0.3)):
        lay = cls(n_classes=8, m=m, name="h")
        x, y = c(g["face/x"]), c(g["face/y"])
        lay.build([tuple(x.shape), tuple(y.shape)])
        lay.set_weights([g["face/%s/W" % key]])
        assert rel_err(lay([x, y]), g["face/%s/out" % key]) < REL_TOL, key
    for gm, m in ((256, 0.25), (256, 0.2), (64, 0.4)):
        got = ls.circle_loss(c(g["circle/y"]), c(g["circle/cos"]), gamma=gm, margin=m)
        assert rel_err(got, g["circle/g%d_m%g" % (gm, m)], floor=1e-3) < REL_TOL


# ---------------------------------------------------------------------------------------------------------------------
# utils.data_loader: fixture made by the reference's own utils.py (tests/golden/make_golden_utils.py)
def _utils_fixture():
    z = np.load(os.path.join(GOLD, "utils_data_loader.npz"))
    lst = [str(u) for u in z["in/lst"]]
    fo = np.concatenate([[0], np.cumsum(z["in/frames"])])
    to = np.concatenate([[0], np.cumsum(z["in/trans_len"])])
    data = {u: z["in/feats"][fo[i]:fo[i + 1]] for i, u in enumerate(lst)}
    acc = {u: int(z["in/accent"][i]) for i, u in enumerate(lst)}
    trans = {u: [int(v) for v in z["in/trans"][to[i]:to[i + 1]]] for i, u in enumerate(lst)}
    kw = {k[3:]: int(z[k]) for k in z.files if k.startswith("kw/")}
    return z, lst, data, acc, trans, kw


def test_oracle_data_loader_reproduces_the_reference_utils_fixture():
    """oracle/fbank_oracle.data_loader == the reference's utils.data_loader (utils.py:71-117, real sklearn MinMaxScaler)
    on ragged pickled features: x_data to float32 rounding, every label tensor exactly, dtypes and shapes."""
    from oracle import fbank_oracle as FO
    z, lst, data, acc, trans, kw = _utils_fixture()
    x, y = FO.data_loader(lst, True, True, True, data, acc, trans, **kw)
    assert {"x/" + k for k in x} | {"y/" + k for k in y} == {k for k in z.files if k[:2] in ("x/", "y/")}
    for k, v in list(x.items()) + list(y.items()):
        ref = z[("x/" if k in x else "y/") + k]
        assert v.shape == ref.shape and v.dtype == ref.dtype, k
        if k == "x_data":
            assert np.abs(v - ref).max() < 3e-7
        else:
            assert np.array_equal(v, ref), k
    # the helper known answers of the same fixture
    from aesrc2020_b200 import utils as us
    assert us.text_ids_norm(list(range(10, 30)), 8) == z["text_ids_norm/long"].tolist()
    assert us.text_ids_norm([5, 6], 8) == z["text_ids_norm/short"].tolist()
    assert np.array_equal(us.feat_reshape(np.arange(12.0).reshape(3, 4), 5), z["feat_reshape/pad"])
    assert np.array_equal(us.feat_reshape(np.arange(12.0).reshape(3, 4), 2), z["feat_reshape/cut"])
    assert us.cal_descriptors(1200, 80) == int(z["cal_descriptors_1200_80"]) == 114


@pytest.mark.gpu
def test_device_data_loader_reproduces_the_reference_utils_fixture(cuda_device):
    """aesrc2020_b200.utils.data_loader (sar_feat_batch_fwd + sar_labels_pack_fwd) == the reference's utils.data_loader
    fixture: x_data within fp32 arithmetic of the float64 MinMaxScaler, labels exactly."""
    from aesrc2020_b200 import utils as us
    z, lst, data, acc, trans, kw = _utils_fixture()
    x, y = us.data_loader(lst, True, True, True, data, acc, trans, **kw)
    for k, v in list(x.items()) + list(y.items()):
        ref = z[("x/" if k in x else "y/") + k]
        got = v.cpu().numpy() if hasattr(v, "cpu") else np.asarray(v)
        assert got.shape == ref.shape and got.dtype == ref.dtype, k
        if k == "x_data":
            assert np.abs(got - ref).max() < 2e-6
        else:
            assert np.array_equal(got, ref), k


# ---------------------------------------------------------------------------------------------------------------------
# Keras weight names (the naming half of the .h5 importer)
def test_keras_weight_names_match_the_reference_creation_order():
    """weights.keras_weight_names == the names obtained by executing the reference's model.py / resnet.py on minikeras
    with Keras' layer auto-naming rule (tests/golden/make_golden_names.py), for all 8 fixture graphs; and SURVEY
    Appendix A's spot values (shortcut convs are created last in a block: conv2d_4, conv2d_11)."""
    with open(os.path.join(GOLD, "keras_names.json")) as f:
        fx = json.load(f)
    assert len(fx) == 8
    for case, rec in fx.items():
        cfg = SARConfig(input_shape=(rec["T"], 80, 1), **rec["kwargs"])
        got = W.keras_weight_names(cfg)
        assert list(got) == list(W.weight_shapes(cfg))
        assert dict(got) == rec["names"], case
    n = fx["cfg2_gvlad_arcface"]["names"]
    assert n["resnet/s1b1/short/kernel"] == "conv2d_4/kernel:0" and n["resnet/s2b1/short/kernel"] == "conv2d_11/kernel:0"
    assert n["CRNN/backward/recurrent_kernel"] == "CRNN/backward_cu_dnngru_1/recurrent_kernel:0"


def test_keras_named_npz_loads_into_canonical_weights(tmp_path):
    """An .npz keyed by Keras weight names (what `np.savez(path, **{w.name: v})` writes on the TF side), in either the
    `layer/weight:0` or the HDF5 `layer/layer/weight:0` spelling, maps back to the canonical container; entries with a
    wrong shape or unknown name are skipped (load_weights(by_name=True, skip_mismatch=True), model.py:181-183)."""
    cfg = SARConfig(input_shape=(200, 80, 1), ctc_enable=True, disc_enable=True, res_type="res18", res_filters=16, mto="bigru",
                    metric_loss="cosface")
    w = W.init_weights(cfg, seed=5)
    names = W.keras_weight_names(cfg)
    assert sum(1 for v in names.values() if "cu_dnngru_3" in v) == 6          # CRNN, CTC_BIGRU, AR_MERGE in creation order
    keras = {names[k]: v for k, v in w.items()}
    back = W.from_keras_named(cfg, keras)
    assert set(back) == set(w) and all(np.array_equal(back[k], w[k]) for k in w)
    h5 = {("%s/%s" % (k.split("/")[0], k)): v for k, v in keras.items()}
    back = W.from_keras_named(cfg, h5)
    assert set(back) == set(w)
    keras["conv2d_1/kernel:0"] = np.zeros((3, 3, 1, 16), np.float32)             # wrong shape -> skipped
    keras["not_a_layer/kernel:0"] = np.zeros(3, np.float32)
    back = W.from_keras_named(cfg, keras)
    assert "resnet/stem/kernel" not in back and len(back) == len(w) - 1
    p = str(tmp_path / "keras_named.npz")
    np.savez(p, **{names[k]: v for k, v in w.items()})
    loaded = W.load_weights(p)
    assert any(k.endswith(":0") for k in loaded)
    back = W.from_keras_named(cfg, loaded)
    assert all(np.array_equal(back[k], w[k]) for k in w)


def test_minikeras_primitives_match_torch():
    """The eager Keras/TF stand-in the fixtures were generated on agrees with PyTorch's own kernels (F.conv2d with explicit
    TF-SAME pads, F.batch_norm, F.max_pool2d, nn.GRU with permuted gates, F.layer_norm(eps=1e-14), F.ctc_loss, F.normalize)
    to 1e-10: the golden files are tied to an implementation independent of minikeras and of the oracle."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    mods = {k: sys.modules.get(k) for k in ("keras", "tensorflow", "keras_layer_normalization")}
    try:
        from check_minikeras import assert_primitives_match_torch
        res = assert_primitives_match_torch()
        assert set(res) == {"conv2d", "bn/pool/dense", "gru/ln", "ctc/l2n"} and max(res.values()) < 1e-10
    finally:                                   # minikeras installs stand-in `keras` / `tensorflow` modules: remove them
        for k in list(sys.modules):
            if k.split(".")[0] in ("keras", "tensorflow", "keras_layer_normalization", "minikeras", "check_minikeras"):
                if mods.get(k) is None:
                    sys.modules.pop(k, None)
