"""CPU oracle for the SAR-Net forward path (TEST INFRASTRUCTURE ONLY).

This package is the float64 CPU restatement of the reference's Keras forward
(`/root/reference` model.py / resnet.py / VLAD.py / losses.py / utils.py /
local/make_fbank.py).  It is the *checker*: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import it.  The product package `aesrc2020_b200` never does.

PARITY STATUS: "parity partially pinned".  The reference ships no tests, no
golden vectors and cannot be executed here (tensorflow / keras /
keras_layer_normalization / python_speech_features are not installed, and
CuDNNGRU has no CPU kernel).  What pins this oracle:
  * the reference's only numeric known-answer, S(1200, 80) = 114
    (train.py:69, utils.py:156-159);
  * `tests/golden/*.npz`: outputs of the reference's OWN source files
    (VLAD.py, losses.py, resnet.py, model.py) executed in this container on
    top of a minimal eager stand-in for the Keras/TF primitives
    (`tests/golden/minikeras/`, script `tests/golden/make_golden.py`) -- this
    pins the graph wiring and the layer arithmetic the reference itself
    writes, NOT the third-party primitives (Conv2D SAME padding, BN eps,
    CuDNNGRU gate layout, LayerNormalization eps, ctc_batch_cost), which
    remain restated from the libraries' documented behaviour;
  * independent cross-checks of those primitives against torch (nn.GRU,
    F.ctc_loss, F.conv2d), brute-force CTC path enumeration and closed forms.
"""
