"""Mirror of the reference's resnet.py call surface on device tensors.

    resnet18_(input, filters=64) / resnet34_(input, filters=64)        resnet.py:170,188
    (B,T,80,1) NHWC in -> (B,H',W',8*filters) NHWC out, after the final BN->ReLU.

`input` is either a CUDA float32 tensor (eager call: runs the kernels and returns the
feature map) or a `ResNetSpec` request built by model.SAR_Net.  Weights come from the
keyword `weights` (canonical-name dict, see weights.py) -- the Keras version creates them
inside; here a seeded synthetic set is created when none is given.

resnet50_/resnet101_/resnet152_ are accepted names but raise NotImplementedError: in the
reference they return keras Model objects (resnet.py:217,233,249) that SAR_Net cannot
consume (model.py:252), i.e. they are unreachable there too.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from .config import SARConfig, resnet_plan
from .engine import ResNetDevice, ResNetTC
from . import weights as _w


def _run(res_type: str, input: torch.Tensor, filters: int, weights: Optional[Dict[str, np.ndarray]], seed: int):
    if input.dim() != 4 or input.shape[-1] != 1:
        raise ValueError("expected (B,T,D,1) NHWC input, got %s" % (tuple(input.shape),))
    T, D = int(input.shape[1]), int(input.shape[2])
    plan = resnet_plan(res_type, filters, T, D)
    if weights is None:
        cfg = SARConfig(input_shape=(T, D, 1), res_type=res_type, res_filters=filters, mto="avg")
        weights = {k: v for k, v in _w.init_weights(cfg, seed).items() if k.startswith("resnet/")}
    tc_ok = all(c.cin % 32 == 0 and c.cout % 32 == 0 for c in plan.convs()[1:])
    dev = (ResNetTC if tc_ok else ResNetDevice)(plan, weights, input.device)
    return dev.forward(input.contiguous())


def resnet18_(input, filters=64, weights=None, seed=1234):
    return _run("res18", input, filters, weights, seed)


def resnet34_(input, filters=64, weights=None, seed=1234):
    return _run("res34", input, filters, weights, seed)


def _unreachable(name):
    def f(input, filters=64, **kw):
        resnet_plan(name, filters, 1200)     # raises NotImplementedError with the explanation
    f.__name__ = name.replace("res", "resnet") + "_"
    return f


resnet50_ = _unreachable("res50")
resnet101_ = _unreachable("res101")
resnet152_ = _unreachable("res152")
