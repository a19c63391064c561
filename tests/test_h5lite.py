"""h5lite: the HDF5 subset of Keras weight files (reference call sites train.py:35, model.py:183,416-417).

h5py is not in the image, so the reader is pinned three ways: (1) round trips through the matching writer, which emits
the structures libhdf5 writes for such files (superblock v0, B-tree v1 + local heap + SNOD groups, object header v1,
contiguous datasets, v1 attributes); (2) byte layouts built by hand from the HDF5 file-format specification for the
structures the writer never emits (object-header continuation blocks, two-level group B-trees, attribute messages v2/v3,
compact and chunked datasets, superblock v1); (3) model-level: save_weights('x.h5') -> load_weights('x.h5') restores every
weight of the thin-ResNet34 configuration under the Keras names pinned by tests/golden/keras_names.json."""
import io
import os
import struct
from collections import OrderedDict

import numpy as np
import pytest

from aesrc2020_b200 import h5lite as H
from aesrc2020_b200 import weights as W
from aesrc2020_b200.config import SARConfig

UNDEF = H.UNDEF


def _layers(n=70, seed=0):
    rng = np.random.RandomState(seed)
    layers = OrderedDict()
    for i in range(n):
        nm = "conv2d_%d" % (i + 1)
        layers[nm] = OrderedDict([("%s/kernel:0" % nm, rng.randn(3, 3, 4, 8).astype(np.float32)),
                                  ("%s/bias:0" % nm, rng.randn(8).astype(np.float32))])
    layers["CRNN"] = OrderedDict([("CRNN/forward_cu_dnngru_1/kernel:0", rng.randn(16, 48).astype(np.float32)),
                                  ("CRNN/backward_cu_dnngru_1/kernel:0", rng.randn(16, 48).astype(np.float32))])
    return layers


@pytest.mark.parametrize("full", [False, True])
def test_round_trip_keras_layout(full):
    layers = _layers()
    bio = io.BytesIO()
    H.write_keras_weights(bio, layers, full_model=full, extra_layers=["activation_1", "add_1"])
    got = H.read_keras_weights(bio.getvalue())
    want = OrderedDict((wn, a) for ws in layers.values() for wn, a in ws.items())
    assert list(got) == list(want)                      # file order = layer_names / weight_names order
    for k in want:
        assert got[k].dtype == np.float32 and np.array_equal(got[k], want[k])
    f = H.H5File(bio.getvalue())
    g = f["model_weights"] if full else f
    assert len(g.keys()) == 73                          # 71 weighted + 2 weightless layers, across several SNODs
    assert g["conv2d_7"]["conv2d_7/kernel:0"].shape == (3, 3, 4, 8)
    assert g["activation_1"].keys() == []
    assert bytes(g.attrs["backend"]) == b"tensorflow"
    with pytest.raises(KeyError):
        g["nope"]


def test_empty_and_scalar_and_dtypes():
    root = H.H5Group()
    root.create_dataset("a/empty", np.zeros((0, 3), np.float32))
    root.create_dataset("a/scalar", np.float64(2.5))
    root.create_dataset("i32", np.arange(-3, 4, dtype=np.int32))
    root.create_dataset("u8", np.arange(7, dtype=np.uint8))
    root.create_dataset("f16", np.linspace(-1, 1, 9).astype(np.float16))
    root.attrs["ints"] = np.array([1, 2, 3], dtype=np.int64)
    root.attrs["name"] = np.bytes_(b"xyz")
    bio = io.BytesIO()
    H.write_h5(bio, root)
    f = H.H5File(bio.getvalue())
    assert f["a/empty"].read().shape == (0, 3)
    assert f["a"]["scalar"].read() == 2.5 and f["a/scalar"].shape == ()
    assert np.array_equal(f["i32"].read(), np.arange(-3, 4)) and f["i32"].dtype == np.int32
    assert np.array_equal(f["u8"].read(), np.arange(7)) and f["u8"].dtype == np.uint8
    assert f["f16"].dtype == np.float16
    assert np.array_equal(f.attrs["ints"], [1, 2, 3]) and bytes(f.attrs["name"]) == b"xyz"
    assert sorted(p for p, _ in f.visit_datasets()) == ["a/empty", "a/scalar", "f16", "i32", "u8"]


def test_not_hdf5_and_unsupported():
    with pytest.raises(H.H5Error):
        H.H5File(b"PK\x03\x04" + b"\0" * 100)
    with pytest.raises(H.H5Error):                      # vlen datatype class is refused for datasets
        dt, _ = H._parse_datatype(bytes([0x19, 0, 0, 0]) + struct.pack("<I", 16) + bytes([0x10, 0, 0, 0]) +
                                  struct.pack("<I", 1) + struct.pack("<HH", 0, 8))
        assert dt.vlen
        raise H.H5Error("vlen")


# ------------------------------------------------------------------ hand-built layouts (HDF5 spec, version 1.8+)
class _Img:
    def __init__(self):
        self.b = bytearray(b"\0" * 2048)                # room for a superblock at 0

    def add(self, data: bytes) -> int:
        a = (len(self.b) + 7) & ~7
        self.b.extend(b"\0" * (a - len(self.b)))
        self.b.extend(data)
        return a

    def superblock(self, root_oh, btree, heap, version=0):
        sb = H.SIG + bytes([version, 0, 0, 0, 0, 8, 8, 0]) + struct.pack("<HH", 4, 16)
        sb += struct.pack("<I", 0)
        if version == 1:
            sb += struct.pack("<HH", 32, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.b), UNDEF)
        sb += struct.pack("<QQII", 0, root_oh, 1, 0) + struct.pack("<QQ", btree, heap)
        self.b[0:len(sb)] = sb
        return bytes(self.b)


def _msg(t, body, flags=0):
    body = body + b"\0" * (-len(body) % 8)
    return struct.pack("<HHBBBB", t, len(body), flags, 0, 0, 0) + body


def _oh(img, msgs, nmsgs=None):
    blob = b"".join(msgs)
    return img.add(struct.pack("<BBHII", 1, 0, nmsgs if nmsgs is not None else len(msgs), 1, len(blob)) + b"\0" * 4 + blob)


def _heap(img, names):
    data = bytearray(b"\0" * 8)
    offs = {}
    for n in names:
        offs[n] = len(data)
        e = n.encode() + b"\0"
        data += e + b"\0" * (-len(e) % 8)
    da = img.add(bytes(data))
    ha = img.add(b"HEAP" + bytes(4) + struct.pack("<QQQ", len(data), UNDEF, da))
    return ha, offs


def _snod(img, entries, offs):
    body = b"".join(struct.pack("<QQII", offs[n], a, 0, 0) + bytes(16) for n, a in entries)
    return img.add(b"SNOD" + bytes([1, 0]) + struct.pack("<H", len(entries)) + body + bytes(40 * (8 - len(entries))))


def _tree(img, level, children, keys):
    body = struct.pack("<Q", keys[0])
    for c, k in zip(children, keys[1:]):
        body += struct.pack("<QQ", c, k)
    return img.add(b"TREE" + bytes([0, level]) + struct.pack("<H", len(children)) + struct.pack("<QQ", UNDEF, UNDEF) + body)


def _f32_dataset(img, a, layout="contiguous"):
    a = np.ascontiguousarray(a, np.float32)
    ds, dt = _msg(1, H._ds_message(a.shape)), _msg(3, H._dt_message(a.dtype))
    if layout == "contiguous":
        da = img.add(a.tobytes())
        lay = _msg(8, bytes([3, 1]) + struct.pack("<QQ", da, a.nbytes))
    elif layout == "compact":
        lay = _msg(8, bytes([3, 0]) + struct.pack("<H", a.nbytes) + a.tobytes())
    elif layout == "v1-contiguous":
        da = img.add(a.tobytes())
        lay = _msg(8, bytes([1, a.ndim, 1, 0, 0, 0, 0, 0]) + struct.pack("<Q", da) + b"".join(struct.pack("<I", d) for d in a.shape))
    else:
        raise ValueError(layout)
    return ds, dt, lay


def test_continuation_block_two_level_btree_and_superblock_v1():
    img = _Img()
    rng = np.random.RandomState(1)
    arrs = OrderedDict(("w%02d" % i, rng.randn(2, 3).astype(np.float32)) for i in range(10))
    names = sorted(arrs)
    ha, offs = _heap(img, names)
    addr = {}
    for i, n in enumerate(names):
        ds, dt, lay = _f32_dataset(img, arrs[n], ["contiguous", "compact", "v1-contiguous"][i % 3])
        if i % 2 == 0:                                   # the layout message lives in a continuation block
            cont = img.add(lay + _msg(0, b""))
            addr[n] = _oh(img, [ds, dt, _msg(0x10, struct.pack("<QQ", cont, len(lay) + 8))], nmsgs=5)
        else:
            addr[n] = _oh(img, [ds, dt, lay])
    # leaves of 3/3/4 names under two level-0 nodes under one level-1 root
    groups = [names[0:3], names[3:6], names[6:10]]
    snods = [_snod(img, [(n, addr[n]) for n in g], offs) for g in groups]
    t0 = _tree(img, 0, snods[:2], [0, offs[groups[0][-1]], offs[groups[1][-1]]])
    t1 = _tree(img, 0, snods[2:], [offs[groups[1][-1]], offs[groups[2][-1]]])
    root_tree = _tree(img, 1, [t0, t1], [0, offs[groups[1][-1]], offs[groups[2][-1]]])
    attr_v3 = struct.pack("<BBHHHB", 3, 0, 3, 8, 16, 0) + b"n3\0" + H._dt_message(np.dtype("S4")) + H._ds_message((2,)) + b"ab\0\0cd\0\0"
    attr_v2 = struct.pack("<BBHHH", 2, 0, 3, 20, 8) + b"n2\0" + H._dt_message(np.dtype(np.float32)) + H._ds_message(()) + struct.pack("<f", 1.5)
    root = _oh(img, [_msg(0x11, struct.pack("<QQ", root_tree, ha)), _msg(0x0C, attr_v3), _msg(0x0C, attr_v2),
                     _msg(0x12, struct.pack("<BxxxI", 1, 0))])
    f = H.H5File(img.superblock(root, root_tree, ha, version=1))
    assert f.sb_version == 1
    assert f.keys() == names
    for n in names:
        assert np.array_equal(f[n].read(), arrs[n]), n
    assert [bytes(x) for x in f.attrs["n3"]] == [b"ab", b"cd"]
    assert float(f.attrs["n2"]) == 1.5


def test_chunked_dataset_without_filters():
    img = _Img()
    a = np.arange(5 * 7, dtype=np.float32).reshape(5, 7)
    cd = (2, 4)
    kids, keys = [], []
    for i0 in range(0, 5, 2):
        for j0 in range(0, 7, 4):
            ch = np.zeros(cd, np.float32)
            blk = a[i0:i0 + 2, j0:j0 + 4]
            ch[:blk.shape[0], :blk.shape[1]] = blk
            kids.append(img.add(ch.tobytes()))
            keys.append(struct.pack("<II", ch.nbytes, 0) + struct.pack("<QQQ", i0, j0, 0))
    keys.append(struct.pack("<II", 0, 0) + struct.pack("<QQQ", 6, 8, 0))
    body = b"".join(k + struct.pack("<Q", c) for k, c in zip(keys, kids)) + keys[-1]
    bt = img.add(b"TREE" + bytes([1, 0]) + struct.pack("<H", len(kids)) + struct.pack("<QQ", UNDEF, UNDEF) + body)
    lay = _msg(8, bytes([3, 2, 3]) + struct.pack("<Q", bt) + struct.pack("<III", 2, 4, 4))
    ds = _oh(img, [_msg(1, H._ds_message(a.shape)), _msg(3, H._dt_message(a.dtype)), lay])
    ha, offs = _heap(img, ["x"])
    sn = _snod(img, [("x", ds)], offs)
    t = _tree(img, 0, [sn], [0, offs["x"]])
    root = _oh(img, [_msg(0x11, struct.pack("<QQ", t, ha))])
    f = H.H5File(img.superblock(root, t, ha))
    assert np.array_equal(f["x"].read(), a)


def test_layer_names_split_into_chunks():
    """keras save_attributes_to_hdf5_group splits attributes above 64 KB into name0, name1, ..."""
    root = H.H5Group()
    root.attrs["layer_names0"] = np.array([b"a", b"b"])
    root.attrs["layer_names1"] = np.array([b"c"])
    for n in "abc":
        g = root.require_group(n)
        g.attrs["weight_names"] = np.array([("%s/kernel:0" % n).encode()])
        g.create_dataset("%s/kernel:0" % n, np.full((2,), ord(n), np.float32))
    bio = io.BytesIO()
    H.write_h5(bio, root)
    got = H.read_keras_weights(bio.getvalue())
    assert list(got) == ["a/kernel:0", "b/kernel:0", "c/kernel:0"] and got["c/kernel:0"][0] == ord("c")


# ------------------------------------------------------------------ model level
def test_model_h5_round_trip(tmp_path):
    cfg = SARConfig(input_shape=(200, 80, 1), ctc_enable=True, disc_enable=True, res_type="res34", res_filters=32,
                    mto="gvlad", vlad_clusters=8, ghost_clusters=2, metric_loss="arcface", bn_dim=64)
    w = W.init_weights(cfg, 5)
    for full in (False, True):
        p = str(tmp_path / ("m%d.h5" % full))
        W.save_weights(p, w, cfg=cfg, full_model=full)
        assert os.path.exists(p) and not os.path.exists(p + ".npz")
        raw = W.load_weights(p)
        assert all(k.endswith(":0") for k in raw)
        assert "conv2d_1/kernel:0" in raw and "CRNN/forward_cu_dnngru_1/recurrent_kernel:0" in raw
        back = W.from_keras_named(cfg, raw)
        assert set(back) == set(w)
        for k in w:
            assert np.array_equal(back[k], w[k]), k


def test_npz_path_is_honoured(tmp_path):
    cfg = SARConfig(input_shape=(200, 80, 1), res_type="res18", res_filters=32, mto="avg")
    w = W.init_weights(cfg, 1)
    p = str(tmp_path / "weights.bin")                   # no '.npz' suffix: np.savez(path) would have appended one
    W.save_weights(p, w)
    assert os.path.exists(p) and not os.path.exists(p + ".npz")
    back = W.load_weights(p)
    assert set(back) == set(w) and all(np.array_equal(back[k], w[k]) for k in w)
    np.savez(str(tmp_path / "legacy"), **{k.replace("/", "|"): v for k, v in w.items()})
    assert set(W.load_weights(str(tmp_path / "legacy"))) == set(w)      # falls back to legacy + '.npz'
