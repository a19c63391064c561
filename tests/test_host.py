"""CPU tests of the host logic: shape plan, weight container, C-ABI surface, SAR_Net argument
behaviour, tap tables of the tensor-core path.  No compute calls (no GPU here)."""
import os
import re

import numpy as np
import pytest

from aesrc2020_b200 import _shim, config, weights as W, tc, utils as us, fbank as fb
from aesrc2020_b200.config import SARConfig, resnet_plan, same_pad

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_header_symbol():
    hdr = open(os.path.join(ROOT, "include", "sarnet.h")).read()
    declared = set(re.findall(r"\b(sar_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"sar_status"}
    lib = _shim.load_library()
    assert declared == set(_shim.SIGNATURES), declared ^ set(_shim.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.sar_version() >= 100 and lib.sar_compiled_arch() == 100


def test_product_path_fails_loudly_without_gpu_or_library(tmp_path):
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(_shim.SarnetError):
            _shim.lib()
    with pytest.raises(_shim.SarnetError):
        _shim.load_library(str(tmp_path / "missing.so"))


def test_plan_matches_survey_appendix_a():
    p = resnet_plan("res34", 32, 500)
    convs = p.convs()
    assert len(convs) == 37 and (p.hout, p.wout, p.cout, p.seq_len) == (16, 3, 256, 48)
    assert abs(p.flops_per_utt() / 1e9 - 1.606) < 1e-3
    s2 = p.blocks[3]
    assert (s2.conv1.stride, s2.conv1.hout, s2.conv1.wout, s2.conv1.pad_t, s2.conv1.pad_l) == (2, 63, 10, 1, 0)
    s4 = p.blocks[13]
    assert (s4.conv1.hout, s4.conv1.wout, s4.conv1.pad_t, s4.conv1.pad_l) == (16, 3, 0, 1)
    assert p.blocks[0].short is not None and p.blocks[0].short.stride == 1      # Q2: 64 -> 32 projection, stride 1
    assert p.blocks[0].conv1.pre_bn is None                                     # resnet.py:111-117
    p18 = resnet_plan("res18", 64, 500)
    assert len(p18.convs()) == 20 and p18.cout == 512 and p18.blocks[0].short is None
    assert abs(p18.flops_per_utt() / 1e9 - 2.937) < 2e-3
    for name in ("res50", "res101", "res152"):
        with pytest.raises(NotImplementedError):
            resnet_plan(name, 64, 500)


def test_weight_container_shapes_and_roundtrip(tmp_path):
    cfg = SARConfig(input_shape=(500, 80, 1), ctc_enable=True, disc_enable=True, res_type="res34", res_filters=32,
                    mto="gvlad", vlad_clusters=64, ghost_clusters=8, metric_loss="arcface")
    shp = W.weight_shapes(cfg)
    assert shp["gvlad_pool/centers"] == (72, 256) and shp["gvlad_center_assignment/kernel"] == (1, 1, 256, 72)
    assert shp["AR_EMBEDDING/kernel"] == (16384, 256) and shp["CRNN/forward/bias"] == (1536,)
    assert shp["CTC_BIGRU/forward/kernel"] == (512, 768) and shp["y_disc/W"] == (256, 8)
    nonres = sum(int(np.prod(s)) for k, s in shp.items() if not k.startswith("resnet/") and
                 not (k.endswith("/bias") and "/forward/" not in k and "/backward/" not in k))
    assert nonres == 6880256                                                     # SURVEY Appendix B total
    w = W.init_weights(cfg, 7)
    p = str(tmp_path / "w.npz")
    W.save_weights(p, w)
    w2 = W.load_weights(p)
    assert set(w) == set(w2) and all(np.array_equal(w[k], w2[k]) for k in w)
    cfg2 = SARConfig(input_shape=(500, 80, 1), disc_enable=True, mto="bigru", metric_loss="circleloss")
    assert "AR_MERGE/forward/kernel" in W.weight_shapes(cfg2) and "y_disc/kernel" in W.weight_shapes(cfg2)


def test_sar_net_argument_behaviour_without_compute():
    from aesrc2020_b200 import model as mdl
    model, train_model = mdl.SAR_Net((300, 80, 1), mto="avg")
    assert train_model is model                                                  # model.py:195-196
    assert model.input_names == ["x_data"] and model.output_names == ["y_accent"]
    m2, _ = mdl.SAR_Net((500, 80, 1), ctc_enable=True, disc_enable=True, bn_dim=32, mto="vlad", metric_loss="cosface")
    assert m2.input_names == ["x_data", "x_accent", "x_ctc_label", "x_ctc_in_len", "x_ctc_out_len"]   # model.py:327-335
    assert m2.output_names == ["y_accent", "y_disc", "y_ctc_loss", "y_disc_bn"]                       # model.py:328-338
    assert m2.get_layer("vlad_pool").get_weights()[0].shape == (8, 256)
    with pytest.raises(SystemExit):
        mdl.SAR_Net((300, 80, 1))                                                # mto=None -> exit(1), model.py:136-138
    with pytest.raises(NotImplementedError):
        mdl.SAR_Net((300, 80, 1), res_type="res50", mto="avg")
    with pytest.raises(_shim.SarnetError):
        mdl.SAR_Net((300, 80, 1), mto="avg", mode="test")                        # Q3
    with pytest.raises(ValueError):
        mdl.SAR_Net((300, 80, 1), mto="avg", weights={})
    assert model.config.loss_weights() == {"y_accent": 1.0}
    assert m2.config.loss_weights() == {"y_accent": 0.01, "y_disc": 0.6, "y_ctc_loss": 0.01, "y_disc_bn": 0.1}


def test_tap_tables_reproduce_tf_same_index_arithmetic():
    """Flat-pad planes: verify the (phase, shift) per tap against explicit TF-SAME indices on a
    numpy emulation of the planes layout (stride 1 and stride 2, even and odd sizes)."""
    rng = np.random.RandomState(0)
    for H, W_, stride in ((7, 6, 1), (7, 6, 2), (8, 5, 2), (5, 3, 2), (16, 3, 1)):
        x = rng.randn(H, W_)
        Ho, pt, _ = same_pad(H, 3, stride)
        Wo, pl, _ = same_pad(W_, 3, stride)
        P = Wo + 1
        if stride == 1:
            planes = np.zeros((1, (Ho + 1) * P + 2 * P + 2))
            base = P + 1                                                        # room for negative rows (TMA zero fill)
            for h in range(H):
                planes[0, base + h * P: base + h * P + W_] = x[h]
        else:
            planes = np.zeros((4, (Ho + 1) * P + 2 * P + 2))
            base = P + 1
            for h in range(H):
                for w in range(W_):
                    planes[(h & 1) * 2 + (w & 1), base + (h >> 1) * P + (w >> 1)] = x[h, w]
        offs, pls = tc.tap_table(3, 3, stride, pt, pl, Wo)
        for ho in range(Ho):
            for wo in range(Wo):
                q = ho * P + wo
                for t, (r, s) in enumerate([(r, s) for r in range(3) for s in range(3)]):
                    hi, wi = ho * stride - pt + r, wo * stride - pl + s
                    want = x[hi, wi] if (0 <= hi < H and 0 <= wi < W_) else 0.0
                    got = planes[pls[t] // 2, base + q + offs[t]]
                    assert got == want, (H, W_, stride, ho, wo, r, s)


def test_pack_weights_layout_and_precision():
    rng = np.random.RandomState(1)
    k = rng.randn(3, 3, 32, 64).astype(np.float32) * 0.05
    s = rng.randn(1, 1, 64, 64).astype(np.float32) * 0.1
    p = tc.pack_weights(k, s)
    assert p.shape == (2, 64, 9 * 32 + 64) and p.dtype == np.float16
    rec = p[0].astype(np.float32) + p[1].astype(np.float32) / 2048.0
    want = np.concatenate([k.reshape(288, 64), s.reshape(64, 64)], 0).T
    assert np.max(np.abs(rec - want)) < 2.0 ** -22 * np.max(np.abs(want)) * 4
    assert rec[5, 2 * 32 + 7] == pytest.approx(k[0, 2, 7, 5], rel=1e-6)          # k = tap*Cin + ci


def test_synthetic_batch_contract_and_utils():
    cfg = SARConfig(input_shape=(500, 80, 1), ctc_enable=True, disc_enable=True, res_type="res34", res_filters=32,
                    mto="gvlad", vlad_clusters=64, ghost_clusters=8, metric_loss="circleloss")
    x, y = us.synthetic_batch(cfg, 5, seed=3, lengths=np.array([200, 300, 500, 250, 499]))
    assert x["x_data"].shape == (5, 500, 80, 1) and x["x_data"].dtype == np.float32        # utils.py:102
    assert (x["x_data"][0, 200:] == 0).all() and x["x_data"][0, :200].max() <= 1.0
    assert x["x_ctc_label"].dtype == np.float32 and x["x_ctc_label"].shape == (5, 72)      # utils.py:107
    assert x["x_ctc_in_len"].dtype == np.int32 and (x["x_ctc_in_len"] == 48).all()         # utils.py:96,105
    assert x["x_ctc_label"].max() <= 998                                                    # blank = 999 (Q8)
    assert np.array_equal(x["x_accent"], y["y_accent"]) and set(y) == {"y_accent", "y_disc", "y_ctc_loss"}
    assert us.text_ids_norm([5, 6, 7], 5) == [5, 6, 7, 2, 2] and us.text_ids_norm(list(range(9)), 4) == [0, 1, 2, 3]
    assert us.cal_descriptors(1200, 80) == 114                                              # utils.py:193
    assert fb.num_frames(16000 * 5) == 499 and fb.mel_filterbank().shape == (80, 257)


def test_model_load_weights_accepts_a_keras_named_npz(tmp_path, capsys):
    """SARModel.load_weights maps an .npz keyed by Keras weight names (`conv2d_1/kernel:0`, ...) to the canonical names."""
    from aesrc2020_b200 import model as mdl, weights as W
    kw = dict(disc_enable=True, res_type="res18", res_filters=16, mto="avg", metric_loss="softmax")
    m, _ = mdl.SAR_Net((200, 80, 1), seed=3, **kw)
    o, _ = mdl.SAR_Net((200, 80, 1), seed=4, **kw)
    capsys.readouterr()
    names = W.keras_weight_names(m.config)
    p = str(tmp_path / "keras_named.npz")
    np.savez(p, **{names[k]: v for k, v in m.weights.items()})
    assert not np.array_equal(o.weights["resnet/stem/kernel"], m.weights["resnet/stem/kernel"])
    o.load_weights(p)
    assert all(np.array_equal(o.weights[k], m.weights[k]) for k in m.weights)


def test_fit_generator_host_loop_epochs_callbacks_and_early_stop():
    """model.fit_generator's host logic (train.py:38-44) with the device trainer stubbed out: `epochs - initial_epoch` epochs of
    `steps_per_epoch` batches, (inputs, targets) tuples split, per-epoch weight sync BEFORE the callbacks run (so a callback's
    model.save stores trained weights), mean logs, duck-typed callbacks, `model.stop_training` ends the run."""
    from aesrc2020_b200 import model as mdl, weights as W
    from aesrc2020_b200.config import SARConfig
    cfg = SARConfig(input_shape=(100, 80, 1), ctc_enable=False, ar_enable=True, disc_enable=True, res_type="res18", res_filters=32,
                    mto="bigru", metric_loss="softmax")
    model = mdl.SARModel(cfg, W.init_weights(cfg, 1))
    events = []

    class FakeTrainer:
        def __init__(self):
            self.n = 0

        def train_on_batch(self, x, y):
            self.n += 1
            events.append(("step", x["x_data"], None if y is None else y["y_accent"]))
            return {"loss": 10.0 / self.n, "loss_accent": 1.0}

        def sync_to_model(self):
            events.append(("sync",))

    model._trainer = FakeTrainer()

    def gen():
        i = 0
        while True:
            i += 1
            yield ({"x_data": i}, {"y_accent": -i}) if i % 2 else {"x_data": i}          # tuples and bare input dicts

    class Stopper:
        def on_epoch_end(self, epoch, logs):
            events.append(("cb", epoch, round(logs["loss"], 4)))
            if epoch == 2:
                model.stop_training = True

    hist = model.fit_generator(gen(), steps_per_epoch=3, epochs=5, callbacks=[Stopper(), object()], initial_epoch=1, verbose=0)
    assert len(hist) == 2                                                    # epochs 1 and 2, then stopped
    steps = [e for e in events if e[0] == "step"]
    assert [s[1] for s in steps] == [1, 2, 3, 4, 5, 6] and [s[2] for s in steps] == [-1, None, -3, None, -5, None]
    assert abs(hist[0]["loss"] - np.mean([10.0, 5.0, 10.0 / 3])) < 1e-12 and hist[0]["loss_accent"] == 1.0
    kinds = [e[0] for e in events]
    assert kinds == ["step"] * 3 + ["sync", "cb"] + ["step"] * 3 + ["sync", "cb"]          # sync precedes the callbacks
    assert [e[1] for e in events if e[0] == "cb"] == [1, 2]
    # compile() keeps the learning rate for the optimiser (model.py:187-201)
    assert mdl.compile(model, 1, lr=0.003) is model and model.lr == 0.003


def test_resnet_trainer_parameter_inventory_matches_the_weights():
    """training_resnet.ResNetTrainer.param_keys (host logic): every ResNet weight of weights.init_weights is either trainable or
    a BN moving statistic, exactly once; the l2 regulariser sits on the conv kernels only (resnet.py:36,56,86,116); the
    first block of thin-ResNet34 carries its projection shortcut; HeadTrainer's key sets for the other slices exist in the weights."""
    from aesrc2020_b200 import weights as W
    from aesrc2020_b200.config import SARConfig
    from aesrc2020_b200.training_resnet import ResNetTrainer
    from oracle import train_oracle as TO
    for res_type, f in (("res18", 64), ("res34", 32)):
        cfg = SARConfig(input_shape=(200, 80, 1), ctc_enable=True, ar_enable=True, disc_enable=True, res_type=res_type, res_filters=f,
                        mto="bigru", metric_loss="softmax")
        w = W.init_weights(cfg, 1)
        keys, l2, stats = ResNetTrainer.param_keys(cfg)
        res = sorted(k for k in w if k.startswith("resnet/"))
        assert sorted(keys + stats) == res and len(set(keys + stats)) == len(keys + stats)
        assert all(k.endswith("/kernel") for k in l2) and set(l2) == {k for k in keys if k.endswith("/kernel")}
        assert all(k.endswith(("/moving_mean", "/moving_variance")) for k in stats)
        assert ("resnet/s1b1/short/kernel" in keys) == (res_type == "res34")
        for k in TO.DS_KEYS + TO.CRNN_KEYS + TO.CTC_KEYS + TO.pool_keys("bigru"):
            assert k in w, k
