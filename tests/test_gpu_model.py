"""GPU parity of the whole forward (SAR_Net(...).predict through the C ABI) against the float64
CPU oracle, for the BASELINE.json configurations at oracle-sized batches, plus the
size-independent properties used at full size (batch-split invariance, padding semantics)."""
import os

import numpy as np
import pytest
import torch

from helpers import rel_err, norm_err, REL_TOL, dev
from oracle import sarnet_oracle as O

pytestmark = pytest.mark.gpu

CONFIGS = {
    # BASELINE.json configs[0]: single 300-frame utterance, res18/64 + AvgPool + Softmax
    "cfg1_res18_avg_softmax": dict(T=300, B=1, kw=dict(res_type="res18", res_filters=64, mto="avg")),
    # configs[1]: 500 frames, ResNet+Bi-GRU+GhostVLAD(64c/8g)+ArcFace
    "cfg2_gvlad_arcface": dict(T=500, B=3, kw=dict(disc_enable=True, res_type="res34", res_filters=32, mto="gvlad",
                                                   vlad_clusters=64, ghost_clusters=8, metric_loss="arcface", margin=0.3)),
    # configs[2]: variable 200-800 frames zero-padded, CTC + Circle-Loss, bigru merge
    "cfg3_ctc_circle_bigru": dict(T=800, B=3, lengths=[200, 517, 800],
                                  kw=dict(ctc_enable=True, disc_enable=True, res_type="res34", res_filters=32, mto="bigru",
                                          metric_loss="circleloss", margin=0.2)),
    # configs[3]: NetVLAD(64c)+CosFace
    "cfg4_vlad_cosface": dict(T=500, B=2, kw=dict(disc_enable=True, res_type="res34", res_filters=32, mto="vlad",
                                                  vlad_clusters=64, metric_loss="cosface", margin=0.3)),
    # configs[4]: full CRNN+GhostVLAD+Circle-Loss+CTC
    "cfg5_gvlad_circle_ctc": dict(T=500, B=2, kw=dict(ctc_enable=True, disc_enable=True, res_type="res34", res_filters=32,
                                                      mto="gvlad", vlad_clusters=64, ghost_clusters=8,
                                                      metric_loss="circleloss", margin=0.2)),
    # remaining heads / bottleneck branch / train.py's hard-coded 8c+2g (Q10) at the reference's T=1200 -> S=114
    "sphereface_bn_T1200": dict(T=1200, B=1, kw=dict(ctc_enable=True, disc_enable=True, res_type="res34", res_filters=32,
                                                     mto="gvlad", vlad_clusters=8, ghost_clusters=2, bn_dim=32,
                                                     metric_loss="sphereface", margin=1.35)),
    "softmax_head_res18_thin": dict(T=200, B=2, kw=dict(disc_enable=True, res_type="res18", res_filters=32, mto="avg",
                                                        metric_loss="softmax")),
}


def build(name):
    from aesrc2020_b200 import model as mdl, utils as us
    c = CONFIGS[name]
    model, train_model = mdl.SAR_Net((c["T"], 80, 1), **c["kw"])
    assert train_model is model                        # gpus == 1 (model.py:195-196)
    x, y = us.synthetic_batch(model.config, c["B"], seed=len(name), lengths=c.get("lengths"))
    return model, x, y


@pytest.mark.parametrize("name", list(CONFIGS))
def test_forward_matches_oracle(cuda_device, name):
    model, x, y = build(name)
    cfg = model.config
    ref = O.sar_net_forward(model.weights, x, **cfg.model_kwargs())
    outs = model.predict(x, batch_size=CONFIGS[name]["B"])
    outs = outs if isinstance(outs, list) else [outs]
    assert [o.shape for o in outs] == [tuple(ref[n].shape) for n in cfg.output_names()]
    for n, got in zip(cfg.output_names(), outs):
        assert rel_err(got, ref[n]) < REL_TOL, n
    # pre-softmax logits and the embedding, through the device dict
    dev_out = model.forward_device(x)
    assert norm_err(dev_out["embedding"], ref["embedding"]) < 1e-4
    assert rel_err(dev_out["y_accent_logits"], ref["y_accent_logits"], floor=1e-2) < REL_TOL
    if cfg.disc_enable:
        assert rel_err(dev_out["y_disc_logits"], ref["y_disc_logits"], floor=1e-2) < REL_TOL
    # Keras-style losses/metrics from the reduced 8-float vector
    tgt = torch.as_tensor(y["y_accent"], dtype=torch.float64)
    want = O.sar_net_losses(ref, tgt, ctc_enable=cfg.ctc_enable, ar_enable=cfg.ar_enable, disc_enable=cfg.disc_enable,
                            bn_dim=cfg.bn_dim, metric_loss=cfg.metric_loss, margin=cfg.margin)
    got = model.evaluate(x, y, batch_size=CONFIGS[name]["B"])
    for k, v in want.items():
        assert abs(got[k] - float(v)) <= REL_TOL * max(abs(float(v)), 1e-3), (k, got[k], float(v))


def test_resnet_surface_and_intermediates(cuda_device):
    """resnet34_(input, filters) call surface + stage-by-stage agreement of the encoder."""
    from aesrc2020_b200 import resnet as rn, model as mdl, utils as us
    model, x, _ = build("cfg5_gvlad_circle_ctc")
    w64 = {k: torch.as_tensor(v, dtype=torch.float64) for k, v in model.weights.items()}
    ref = O.sar_net_forward(model.weights, x, **model.config.model_kwargs(), return_intermediates=True)
    fmap = rn.resnet34_(torch.from_numpy(x["x_data"]).cuda(), filters=32, weights=model.weights)
    assert tuple(fmap.shape) == tuple(ref["resnet"].shape) == (2, 16, 3, 256)
    assert norm_err(fmap, ref["resnet"]) < 1e-4
    out = model.forward_device(x, want_intermediates=True)
    for k in ("cnn_lin", "crnn", "ar_ds", "integration", "ctc_pred"):
        assert norm_err(out[k], ref[k]) < 1e-4, k
    with pytest.raises(NotImplementedError):
        rn.resnet50_(torch.zeros(1, 200, 80, 1, device="cuda"))


def test_batch_split_invariance_and_padding(cuda_device):
    """Size-independent properties used at BASELINE sizes: per-utterance outputs do not depend
    on batch composition (bitwise), and zero-padded frames are COMPUTED ON, not masked (Q4)."""
    from aesrc2020_b200 import model as mdl, utils as us
    kw = dict(ctc_enable=True, disc_enable=True, res_type="res34", res_filters=32, mto="gvlad", vlad_clusters=64,
              ghost_clusters=8, metric_loss="arcface")
    model, _ = mdl.SAR_Net((500, 80, 1), **kw)
    x, _ = us.synthetic_batch(model.config, 64, seed=5)
    full = model.predict(x, batch_size=64)
    halves = model.predict(x, batch_size=32)
    ragged = model.predict(x, batch_size=7)
    for a, b, c in zip(full, halves, ragged):
        assert np.array_equal(a, b) and np.array_equal(a, c)
    assert all(np.isfinite(a).all() for a in full)
    assert np.allclose(full[0].sum(-1), 1.0, atol=1e-5) and np.allclose(full[1].sum(-1), 1.0, atol=1e-5)
    # padding semantics: zeroing the tail of an utterance changes its output (no masking) ...
    x2 = {k: v.copy() for k, v in x.items()}
    x2["x_data"][0, 300:] = 0.0
    o2 = model.predict(x2, batch_size=64)
    assert not np.array_equal(o2[0][0], full[0][0])
    # ... but leaves every other utterance bit-identical (utterances are independent)
    assert np.array_equal(o2[0][1:], full[0][1:])


def test_weights_roundtrip_and_layers(cuda_device, tmp_path):
    from aesrc2020_b200 import model as mdl, utils as us
    model, x, _ = build("cfg4_vlad_cosface")
    a = model.predict(x)
    p = str(tmp_path / "demo.npz")
    model.save_weights(p)                                  # model.py:416-417 round trip
    other, _ = mdl.SAR_Net((500, 80, 1), **CONFIGS["cfg4_vlad_cosface"]["kw"], seed=99)
    assert not np.array_equal(other.predict(x)[0], a[0])
    other.load_weights(p)
    assert all(np.array_equal(u, v) for u, v in zip(other.predict(x), a))
    lay = model.get_layer("vlad_pool")
    assert lay.get_weights()[0].shape == (64, 256)         # VLAD.py:17-19
    sub = mdl.sub_model(model, "x_data", "y_accent")       # model.py:415
    assert np.array_equal(sub.predict(x), a[0])


def test_device_tensor_inputs_are_zero_copy(cuda_device):
    model, x, _ = build("cfg2_gvlad_arcface")
    xd = {k: torch.from_numpy(v).cuda() for k, v in x.items()}
    outs = model.predict(xd, batch_size=8)
    assert all(isinstance(o, torch.Tensor) and o.is_cuda for o in outs)
    host = model.predict(x, batch_size=8)
    assert all(np.array_equal(o.cpu().numpy(), h) for o, h in zip(outs, host))


def test_ctc_infeasible_raises(cuda_device):
    from aesrc2020_b200 import model as mdl, utils as us, _shim
    model, x, _ = build("cfg5_gvlad_circle_ctc")
    x["x_ctc_out_len"][:] = 40                             # 40 labels cannot fit in S=48 with repeats forced below
    x["x_ctc_label"][:, :40] = 7.0
    with pytest.raises(_shim.SarnetError):
        model.predict(x)


def test_predict_host_paths_agree_bitwise(cuda_device):
    """model.predict() with pageable numpy inputs (staged through an internal pinned buffer), with inputs that
    already live in pinned memory (utils.pinned_like: DMA'd in place) and with device tensors gives identical
    bits; repeated calls reuse the staging buffer safely."""
    from aesrc2020_b200 import model as mdl, utils as us
    kw = dict(ctc_enable=True, disc_enable=True, res_type="res34", res_filters=32, mto="gvlad", vlad_clusters=8,
              ghost_clusters=2, metric_loss="arcface", margin=0.3)
    model, _ = mdl.SAR_Net((200, 80, 1), **kw)
    xs = [us.synthetic_batch(model.config, 4, seed=s)[0] for s in (1, 2)]
    ref = [model.predict(x, batch_size=4) for x in xs]
    for x, want in zip(xs, ref):
        got_pin = model.predict(us.pinned_like(x), batch_size=4)
        got_dev = model.predict({k: model._to_device(k, v) for k, v in x.items()}, batch_size=4)
        got_chunks = model.predict(x, batch_size=3)              # ragged second chunk, staging buffer re-used
        for a, b, c, d in zip(want, got_pin, got_dev, got_chunks):
            assert np.array_equal(a, b)
            assert np.array_equal(a, c.cpu().numpy())
            assert np.array_equal(a, d)
    assert all(np.array_equal(a, b) for a, b in zip(ref[0], model.predict(xs[0], batch_size=4)))


def test_stage_chain_launch_is_bitwise_the_layer_by_layer_path(cuda_device):
    """sar_conv_tc_chain_fwd (all stride-1 3x3 layers of a stage in one persistent launch, tiles synchronised by
    per-M-tile counters) vs one sar_conv_tc_fwd launch per layer: identical bits, also on repeated calls (the
    counters are self-cleaning) and with every stage chained."""
    from aesrc2020_b200 import model as mdl, utils as us
    kw = dict(disc_enable=True, res_type="res34", res_filters=32, mto="gvlad", vlad_clusters=8, ghost_clusters=2,
              metric_loss="arcface", margin=0.3)
    model, _ = mdl.SAR_Net((500, 80, 1), **kw)
    model.use_graph = False
    x, _ = us.synthetic_batch(model.config, 6, seed=11)
    rn = model.engine().resnet
    rn.chain_stages = set()
    want = model.predict(x, batch_size=6)
    for stages in ({2, 3, 4}, {1, 2, 3, 4}, {3}):
        rn.chain_stages = stages
        for _ in range(2):
            got = model.predict(x, batch_size=6)
            for a, b in zip(want, got):
                assert np.array_equal(a, b), stages


@pytest.mark.gpu
@pytest.mark.parametrize("ctc", [False, True])
def test_micro_batch_lanes_are_bitwise_the_single_stream_step(cuda_device, ctc):
    """engine.forward_lanes: the step as 2 / 3 / 4 concurrent micro-batch graphs on separate streams returns bitwise
    the outputs AND the batch loss vector of the single-graph step (utterances are independent; the loss reduce runs
    once over all rows after the lanes join), step after step with changing inputs, for host and device inputs."""
    from aesrc2020_b200 import model as mdl, utils as us
    kw = dict(ctc_enable=ctc, disc_enable=True, res_type="res34", res_filters=32, mto="gvlad", vlad_clusters=64,
              ghost_clusters=8, metric_loss="circleloss" if ctc else "arcface")
    model, _ = mdl.SAR_Net((500, 80, 1), **kw)
    eng = model.engine()
    batches = [us.synthetic_batch(model.config, 50, seed=40 + i)[0] for i in range(3)]   # 50: lanes of unequal size
    model.lanes = 1
    ref = [model.predict(x, batch_size=50) for x in batches]
    dev = [{k: model._to_device(k, v).clone() for k, v in x.items()} for x in batches]
    ref_vec = [eng.forward_graphed(d)["loss_vector"].cpu().numpy().copy() for d in dev]
    for lanes in (2, 3, 4):
        model.lanes = lanes
        for rep in range(2):
            for x, d, r, rv in zip(batches, dev, ref, ref_vec):
                got = model.predict(x, batch_size=50)
                assert all(np.array_equal(a, b) for a, b in zip(got, r)), (lanes, rep)
                out = eng.forward_lanes(d, lanes)
                assert np.array_equal(out["loss_vector"].cpu().numpy(), rv), (lanes, rep)
                assert np.array_equal(out["y_accent"].cpu().numpy(), r[0])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cfg2_gvlad_arcface", "cfg5_gvlad_circle_ctc"])
def test_predict_generator_pipeline_is_bitwise_predict(cuda_device, name):
    """Keras-style predict_generator (H2D of step i+1 and D2H of step i-1 under step i's kernels): bitwise the
    per-batch predict() outputs, for (inputs, targets) tuples, plain dicts, pinned and pageable arrays, a ragged last
    batch and `steps` shorter than the generator."""
    from aesrc2020_b200 import utils as us
    model, _, _ = build(name)
    sizes = [8, 8, 8, 8, 5]
    batches = [us.synthetic_batch(model.config, b, seed=70 + i) for i, b in enumerate(sizes)]
    want = [model.predict(x, batch_size=len(x["x_data"])) for x, _ in batches]
    want = [w if isinstance(w, list) else [w] for w in want]
    cat = [np.concatenate([w[i] for w in want], 0) for i in range(len(want[0]))]

    def gen(pinned, tuples):
        for x, y in batches:
            xx = us.pinned_like(x) if pinned else x
            yield (xx, y) if tuples else xx
    for pinned in (False, True):
        got = model.predict_generator(gen(pinned, tuples=pinned))
        got = got if isinstance(got, list) else [got]
        assert all(np.array_equal(a, b) for a, b in zip(got, cat)), pinned
    got = model.predict_generator(gen(True, True), steps=3)
    got = got if isinstance(got, list) else [got]
    assert all(np.array_equal(a, b[:24]) for a, b in zip(got, cat))


def test_predict_generator_raises_on_infeasible_ctc(cuda_device):
    from aesrc2020_b200 import _shim
    model, x, _ = build("cfg5_gvlad_circle_ctc")
    bad = {k: v.copy() for k, v in x.items()}
    bad["x_ctc_out_len"][:] = 40
    bad["x_ctc_label"][:, :40] = 7.0
    with pytest.raises(_shim.SarnetError):
        model.predict_generator(iter([x, bad, x]))
    assert np.array_equal(model.predict_generator(iter([x]))[0], model.predict(x)[0])     # the pipeline survives the error


def test_pipelined_slots_are_bitwise_the_single_stream_step(cuda_device):
    """engine.forward_slot: consecutive batches in flight on 3 streams / graphs / buffer sets (what bench.py's value
    loop and predict_generator do) give bitwise the single-stream outputs, batch after batch."""
    from aesrc2020_b200 import model as mdl, utils as us
    kw = dict(ctc_enable=True, disc_enable=True, res_type="res34", res_filters=32, mto="gvlad", vlad_clusters=64,
              ghost_clusters=8, metric_loss="arcface")
    model, _ = mdl.SAR_Net((500, 80, 1), **kw)
    eng = model.engine()
    names = ["y_accent", "y_disc", "y_ctc_loss", "loss_vector"]
    dev = [{k: model._to_device(k, v).clone() for k, v in us.synthetic_batch(model.config, 24, seed=90 + i)[0].items()}
           for i in range(7)]
    ref = []
    for d in dev:
        out = eng.forward_graphed(d)
        ref.append({k: out[k].clone() for k in names})
    torch.cuda.synchronize()
    for rep in range(2):
        got = []
        for i, d in enumerate(dev):
            out, st = eng.forward_slot(d, i % 3)
            with torch.cuda.stream(st):
                got.append({k: out[k].clone() for k in names})      # before the slot's next replay overwrites them
        torch.cuda.synchronize()
        for g, r in zip(got, ref):
            assert all(torch.equal(g[k], r[k]) for k in names), rep


def test_ctc_pred_greedy_decode_matches_posteriors(cuda_device):
    """mdl.ctc_pred (model.py:385-389) through the graphed step + sar_ctc_greedy_fwd = the oracle's greedy decode of the
    SAME device posteriors (the `ctc_pred` intermediate), for full and truncated input_len."""
    from aesrc2020_b200 import model as mdl
    model, x, _ = build("cfg5_gvlad_circle_ctc")
    probs = model.forward_device(x, want_intermediates=True)["ctc_pred"].cpu().numpy()
    S = probs.shape[1]
    for T in (S, S // 2):
        got = mdl.ctc_pred(model, x, batch_size=2, input_len=T)
        assert np.array_equal(got, O.ctc_greedy_decode(probs, T))


def test_data_loader_feeds_predict_without_host_round_trip(cuda_device):
    """utils.data_loader's device tensors go straight into model.predict (zero-copy path) and give what the host-assembled
    batch (oracle data_loader -> numpy -> predict) gives."""
    from aesrc2020_b200 import utils as us
    from oracle import fbank_oracle as FO
    model, _, _ = build("cfg5_gvlad_circle_ctc")
    rng = np.random.RandomState(8)
    lst = ["a", "b", "c"]
    data = {u: rng.rand(n, 80).astype(np.float32) * 9 for u, n in zip(lst, [480, 500, 650])}
    acc = {"a": 1, "b": 7, "c": 0}
    trans = {u: [int(v) for v in rng.randint(3, 998, size=n)] for u, n in zip(lst, [5, 9, 3])}
    kw = dict(max_input_len=500, max_ctc_len=72, encoder_len=model.config.plan().seq_len, accent_classes=8)
    xd, _ = us.data_loader(lst, True, True, True, data, acc, trans, **kw)
    xh, _ = FO.data_loader(lst, True, True, True, data, acc, trans, **kw)
    got = model.predict(xd, batch_size=3)
    want = model.predict(xh, batch_size=3)
    assert all(isinstance(g, torch.Tensor) and g.is_cuda for g in got)
    for g, w in zip(got, want):
        assert np.allclose(g.cpu().numpy(), w, rtol=2e-4, atol=1e-6)


def test_ctc_pred_is_decode_only_and_sub_model_exposes_the_posteriors(cuda_device):
    """the reference decodes on sub_model(model, 'x_data', 'ctc_pred') with x as the only input
    (model.py:380-389).  ctc_pred() must not need label inputs, dummy labels must not trip the infeasible-CTC check, and
    the sub-model's predict() returns the (B, S, bpe_classes) posteriors."""
    from aesrc2020_b200 import model as mdl
    model, x, _ = build("cfg5_gvlad_circle_ctc")
    S = model.config.plan().seq_len
    full = model.forward_device(x, want_intermediates=True)["ctc_pred"].cpu().numpy()
    only_x = {"x_data": x["x_data"]}
    got = mdl.ctc_pred(model, only_x, batch_size=2, input_len=S)
    assert np.array_equal(got, O.ctc_greedy_decode(full, S))
    bad = dict(x)
    bad["x_ctc_out_len"] = np.full_like(np.asarray(x["x_ctc_out_len"]), 71)      # infeasible for the loss: ignored here
    assert np.array_equal(mdl.ctc_pred(model, bad, batch_size=2, input_len=S), got)
    sm = mdl.sub_model(model, "x_data", "ctc_pred")
    assert sm.required_inputs() == ["x_data"]
    probs = sm.predict(only_x, batch_size=2)
    assert probs.shape == full.shape
    assert rel_err(probs, full) < 1e-5 and np.allclose(probs.sum(-1), 1.0, atol=1e-5)
    assert np.array_equal(mdl.ctc_pred(sm, x["x_data"], batch_size=1, input_len=S), got)
    with pytest.raises(ValueError):
        mdl.sub_model(mdl.SAR_Net((200, 80, 1), res_type="res18", res_filters=32, mto="avg")[0], "x_data", "ctc_pred")


def test_predict_generator_accepts_device_batches_from_data_loader(cuda_device):
    """predict_generator's staging copy runs on a private copy stream; device tensors produced on the caller's
    stream (utils.data_loader kernels) must be ordered before it.  Many rounds with fresh tensors every step (so the
    caching allocator recycles them) against per-batch predict."""
    from aesrc2020_b200 import utils as us
    model, _, _ = build("cfg5_gvlad_circle_ctc")
    rng = np.random.RandomState(11)
    lst = ["a", "b", "c", "d"]
    kw = dict(max_input_len=500, max_ctc_len=72, encoder_len=model.config.plan().seq_len, accent_classes=8)
    datas = []
    for r in range(6):
        data = {u: rng.rand(n, 80).astype(np.float32) * 9 for u, n in zip(lst, rng.randint(300, 700, size=4))}
        acc = {u: int(rng.randint(0, 8)) for u in lst}
        trans = {u: [int(v) for v in rng.randint(3, 998, size=rng.randint(3, 9))] for u in lst}
        datas.append((data, acc, trans))

    def gen():
        for data, acc, trans in datas:
            yield us.data_loader(lst, True, True, True, data, acc, trans, **kw)      # (inputs, targets), device tensors
    want = [model.predict(us.data_loader(lst, True, True, True, d, a, t, **kw)[0], batch_size=4) for d, a, t in datas]
    got = model.predict_generator(gen())
    for i in range(len(got)):
        w = np.concatenate([(o[i].cpu().numpy() if isinstance(o[i], torch.Tensor) else o[i]) for o in want], 0)
        assert np.allclose(got[i], w, rtol=2e-4, atol=1e-6), i


def test_ds_softmax_wider_than_the_head_kernel(cuda_device):
    """DS(1000, 'softmax') (the reference's ctc_pred layer, model.py:268) through the layer surface."""
    from aesrc2020_b200 import model as mdl
    rng = np.random.RandomState(3)
    x = rng.randn(7, 5, 64).astype(np.float32)
    for n in (8, 33, 1000, 1500):
        layer = mdl.DS(n, "softmax", name="d")
        w = {"kernel": rng.randn(64, n).astype(np.float32) * 0.3, "bias": rng.randn(n).astype(np.float32)}
        layer.set_weights_dict(w)
        got = layer(dev(x)).cpu().numpy()
        z = x.astype(np.float64) @ w["kernel"].astype(np.float64) + w["bias"]
        e = np.exp(z - z.max(-1, keepdims=True))
        assert got.shape == (7, 5, n) and rel_err(got, e / e.sum(-1, keepdims=True)) < 2e-4


def test_model_saves_and_loads_keras_h5(cuda_device, tmp_path):
    """model.save_weights('x.h5') / model.save('x.h5') (train.py:35) write the Keras HDF5 layout at exactly that path and
    load_weights / raw_model read it back by name (model.py:181-183)."""
    from aesrc2020_b200 import model as mdl, h5lite
    model, x, _ = build("cfg2_gvlad_arcface")
    kw = CONFIGS["cfg2_gvlad_arcface"]["kw"]
    want = model.predict(x, batch_size=8)
    for fn, name in ((model.save_weights, "w.h5"), (model.save, "m.h5")):
        p = str(tmp_path / name)
        fn(p)
        assert os.path.exists(p) and not os.path.exists(p + ".npz")
        names = list(h5lite.read_keras_weights(p))
        assert "conv2d_1/kernel:0" in names and "gvlad_pool/centers:0" in names
        m2, _ = mdl.SAR_Net((model.config.input_shape[0], 80, 1), seed=999, raw_model=p, **kw)
        got = m2.predict(x, batch_size=8)
        for g, w in zip(got, want):
            assert np.array_equal(g, w)


FULL_SIZE = {
    # BASELINE.json configs at their REAL per-GPU size on the device; the float64 oracle runs on a 32-utterance slice of
    # the same batch (utterances are independent on this path, so the slice's reference is the batch's reference)
    "configs[1] B=64 T=500": dict(cfg="cfg2_gvlad_arcface", B=64, T=500, lo=16, n=32, slot=False),
    "configs[2] B=256 T=800 (slice of 32)": dict(cfg="cfg3_ctc_circle_bigru", B=256, T=800, lo=101, n=32, slot=True,
                                                  lengths=True),
    "configs[4] shard B=512 T=500 (slice of 32)": dict(cfg="cfg5_gvlad_circle_ctc", B=512, T=500, lo=333, n=32, slot=True),
}


@pytest.mark.parametrize("name", list(FULL_SIZE))
def test_full_size_batch_matches_oracle_on_a_slice(cuda_device, name):
    """oracle comparisons at the sizes the bench runs -- many-tile persistent conv loops, the
    32-utterance Bi-GRU clusters (`bigru_tc_kernel<32>`), CTC shared-memory sizing at S = 75, the multi-item VLAD CTAs --
    not only at B = 1-3.  `slot`: through engine.forward_slot (the kernel choices the headline `value` is measured on)."""
    from aesrc2020_b200 import model as mdl, utils as us
    c = FULL_SIZE[name]
    kw = CONFIGS[c["cfg"]]["kw"]
    model, _ = mdl.SAR_Net((c["T"], 80, 1), **kw)
    cfg = model.config
    lengths = np.random.RandomState(7).randint(200, 801, size=c["B"]) if c.get("lengths") else None
    x, _ = us.synthetic_batch(cfg, c["B"], seed=77, lengths=lengths)
    sl = slice(c["lo"], c["lo"] + c["n"])
    ref = O.sar_net_forward(model.weights, {k: v[sl] for k, v in x.items()}, **cfg.model_kwargs())
    if c["slot"]:
        dev_in = {k: model._to_device(k, v) for k, v in x.items()}
        out, st = model.engine().forward_slot(dev_in, 1)
        st.synchronize()
        got = {n: out[n][sl].cpu().numpy() for n in cfg.output_names()}
        assert int(out["ctc_status"].abs().max()) == 0
    else:
        outs = model.predict(x, batch_size=c["B"])
        got = {n: o[sl] for n, o in zip(cfg.output_names(), outs)}
    for n in cfg.output_names():
        # Circle-Loss's y_disc are RAW COSINES in [-1, 1] (model.py:161-163), many of them near zero: 1e-3 relative with
        # the floor at 1e-2 of the range (the same floor the pre-softmax logits get above); probabilities and the CTC
        # loss keep the 1e-6 floor
        floor = 1e-2 if (n == "y_disc" and cfg.metric_loss == "circleloss") else 1e-6
        assert rel_err(got[n], ref[n], floor=floor) < REL_TOL, (name, n, rel_err(got[n], ref[n], floor=floor))


def test_predict_generator_prefetch_ring_matches_predict(cuda_device):
    """predict_generator(prefetch=True) -- a utils.PinnedRing worker thread stages pageable host batches into
    a ring of pinned buffers (and runs device-side generators under its own stream + ready events) -- returns bitwise what
    per-batch predict returns, for host batches and for utils.data_generator's device batches."""
    from aesrc2020_b200 import utils as us
    model, _, _ = build("cfg5_gvlad_circle_ctc")
    host = [us.synthetic_batch(model.config, 6, seed=200 + i)[0] for i in range(9)]
    want = [model.predict(h, batch_size=6) for h in host]
    got = model.predict_generator(iter(host), prefetch=True, max_queue_size=3)
    for i in range(len(got)):
        assert np.array_equal(got[i], np.concatenate([w[i] for w in want], 0)), i
    # device batches produced on the ring's worker thread (data_generator -> data_loader kernels)
    rng = np.random.RandomState(4)
    lst = ["u%d" % i for i in range(12)]
    data = {u: rng.rand(int(n), 80).astype(np.float32) * 7 for u, n in zip(lst, rng.randint(300, 700, size=12))}
    acc = {u: int(rng.randint(0, 8)) for u in lst}
    trans = {u: [int(v) for v in rng.randint(3, 998, size=rng.randint(3, 9))] for u in lst}
    kw = dict(ctc_enable=True, ar_enable=True, disc_enable=True, batch_size=4, data_dct=data, accent_dct=acc, trans_dct=trans,
              max_input_len=500, max_ctc_len=72, encoder_len=model.config.plan().seq_len, accent_classes=8)
    ref_batches = []
    g = us.data_generator(lst, seed=9, **kw)
    for _ in range(6):
        xin, _ = next(g)
        ref_batches.append([o.cpu().numpy() for o in model.predict(xin, batch_size=4)])
    got = model.predict_generator(us.data_generator(lst, seed=9, **kw), steps=6, prefetch=True, max_queue_size=2)
    for i in range(len(got)):
        assert np.allclose(got[i], np.concatenate([r[i] for r in ref_batches], 0), rtol=2e-4, atol=1e-6), i


def test_activation_overflow_fails_loudly(cuda_device):
    """The tensor-core operand planes are fp16 hi/lo pairs (|x| < 65504).  Features far outside the [0, 1] range
    utils.feat_norm produces saturate them; the outputs turn NaN and predict / predict_generator raise instead of
    returning them (the in-range batch before and after is unaffected)."""
    from aesrc2020_b200 import _shim
    model, x, _ = build("cfg2_gvlad_arcface")
    good = model.predict(x, batch_size=len(x["x_data"]))
    bad = dict(x, x_data=x["x_data"] * np.float32(3e6))
    with pytest.raises(_shim.SarnetError, match="non-finite"):
        model.predict(bad, batch_size=len(x["x_data"]))
    with pytest.raises(_shim.SarnetError, match="non-finite"):
        model.predict_generator(iter([x, bad, x]))
    again = model.predict(x, batch_size=len(x["x_data"]))
    for a, b in zip(good, again):
        assert np.array_equal(a, b)
