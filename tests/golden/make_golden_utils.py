"""Generate tests/golden/utils_data_loader.npz by executing the reference's OWN utils.py.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python tests/golden/make_golden_utils.py

`/root/reference/utils.py` is imported UNMODIFIED.  Its third-party imports are the real `sklearn` MinMaxScaler (in this
image) and `keras.utils.np_utils.to_categorical`, for which a 6-line numpy stand-in with Keras 2.2's documented behaviour
(float32 one-hot, last axis = classes, scalar input -> vector) is installed.  The script pickles seeded synthetic feature
matrices the way the reference's feature store does (utils.py:14-22), calls `utils.data_loader(...)` (utils.py:71-117)
exactly as data_generator does (utils.py:120-154), and stores its inputs and outputs.
"""
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def _install_np_utils():
    def to_categorical(y, num_classes=None, dtype="float32"):
        y = np.array(y, dtype="int")
        shape = y.shape
        if shape and shape[-1] == 1 and len(shape) > 1:
            shape = tuple(shape[:-1])
        y = y.ravel()
        if not num_classes:
            num_classes = int(np.max(y)) + 1
        out = np.zeros((y.shape[0], num_classes), dtype=dtype)
        out[np.arange(y.shape[0]), y] = 1
        return out.reshape(shape + (num_classes,))
    for name in ("keras", "keras.utils", "keras.utils.np_utils"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["keras.utils.np_utils"].to_categorical = to_categorical
    sys.modules["keras.utils"].np_utils = sys.modules["keras.utils.np_utils"]
    sys.modules["keras"].utils = sys.modules["keras.utils"]


def main():
    _install_np_utils()
    sys.path.insert(0, REF)
    import utils as ref_utils                                 # /root/reference/utils.py
    assert ref_utils.__file__.startswith(REF)
    rng = np.random.RandomState(20201017)
    lst = ["utt%02d" % i for i in range(9)]
    frames = [37, 120, 64, 333, 1, 100, 99, 101, 250]         # 1 frame ... longer than max_input_len
    # float32-exact values (stored as float32, handed to the reference as float64 like unpickled psf.fbank output)
    feats = {u: (rng.rand(n, 80) * rng.uniform(1, 60) - rng.uniform(0, 5)).astype(np.float32).astype(np.float64)
             for u, n in zip(lst, frames)}
    feats["utt02"][:, 7] = 3.25                               # constant column: MinMaxScaler zero range
    feats["utt06"][:, 0] = 0.0
    accent = {u: str((3 * i + 1) % 8) for i, u in enumerate(lst)}          # the list files hold strings (utils.py:165-170)
    trans = {u: [int(v) for v in rng.randint(3, 999, size=n)] for u, n in zip(lst, [4, 90, 1, 72, 10, 73, 30, 0, 71])}
    kw = dict(max_input_len=100, max_ctc_len=72, encoder_len=13, accent_classes=8, bn=1)
    with tempfile.TemporaryDirectory() as tmp:
        paths = {}
        for u in lst:
            paths[u] = os.path.join(tmp, u + ".pkl")
            ref_utils.save(paths[u], feats[u])                # utils.py:14-17
        x, y = ref_utils.data_loader(lst, ctc_enable=True, ar_enable=True, disc_enable=True, data_dct=paths,
                                     accent_dct=accent, trans_dct=trans, **kw)
    out = {"in/lst": np.array(lst), "in/frames": np.array(frames, dtype=np.int64),
           "in/feats": np.concatenate([feats[u] for u in lst], 0).astype(np.float32),
           "in/accent": np.array([int(accent[u]) for u in lst], dtype=np.int32),
           "in/trans_len": np.array([len(trans[u]) for u in lst], dtype=np.int64),
           "in/trans": np.array([t for u in lst for t in trans[u]], dtype=np.int32)}
    out.update({"kw/" + k: np.array(v) for k, v in kw.items()})
    out.update({"x/" + k: np.asarray(v) for k, v in x.items()})
    out.update({"y/" + k: np.asarray(v) for k, v in y.items()})
    # the small helpers, called directly
    out["text_ids_norm/long"] = np.array(ref_utils.text_ids_norm(list(range(10, 30)), 8))
    out["text_ids_norm/short"] = np.array(ref_utils.text_ids_norm([5, 6], 8))
    out["feat_reshape/pad"] = ref_utils.feat_reshape(np.arange(12.0).reshape(3, 4), 5)
    out["feat_reshape/cut"] = ref_utils.feat_reshape(np.arange(12.0).reshape(3, 4), 2)
    out["cal_descriptors_1200_80"] = np.array(ref_utils.cal_descriptors(1200, 80))
    path = os.path.join(HERE, "utils_data_loader.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if k.startswith(("x/", "y/"))})


if __name__ == "__main__":
    main()
