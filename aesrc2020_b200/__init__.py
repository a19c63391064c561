"""aesrc2020_b200 -- B200-native (sm_100a) forward engine for the SAR-Net accent-recognition
path of pika-online/AESRC2020 (fbank -> ResNet -> Bi-GRU -> {Avg|Bi-GRU|NetVLAD|GhostVLAD}
-> margin-softmax head, + CTC auxiliary), behind the reference's Python call surface.

Modules mirror the reference's file names: model, resnet, VLAD, losses, utils (+ fbank for
local/make_fbank.py).  Compute goes through the C ABI of csrc/libsarnet_sm100.so
(include/sarnet.h); there is no CPU or PyTorch fallback.
"""
from .config import SARConfig, resnet_plan, encoder_len  # noqa: F401

__version__ = "0.1.0"
