"""ctypes binding of libsarnet_sm100.so (the C ABI declared in include/sarnet.h).

PyTorch is plumbing here: it owns device memory and streams; every tensor crosses the
boundary as a raw device pointer + sizes.  There is NO fallback: if the library is missing,
cannot be loaded, or no CUDA device is present, the product path raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libsarnet_sm100.so")

_lib = None

c_fp = C.c_void_p      # const float* / float* (device)
c_ip = C.c_void_p
c_int = C.c_int
c_f = C.c_float
c_ll = C.c_longlong
c_sz = C.c_size_t

# name -> (restype, argtypes); MUST list every symbol include/sarnet.h declares
SIGNATURES = {
    "sar_version": (c_int, []),
    "sar_last_error": (C.c_char_p, []),
    "sar_compiled_arch": (c_int, []),
    "sar_conv2d_fwd": (c_int, [c_fp] * 9 + [c_int] * 13 + [C.c_void_p]),
    "sar_planes_bytes": (c_sz, [c_int] * 5),
    "sar_planes_pack_fwd": (c_int, [c_fp, c_fp, c_fp, c_int, c_fp] + [c_int] * 5 + [C.c_void_p]),
    "sar_planes_unpack_fwd": (c_int, [c_fp, c_fp] + [c_int] * 5 + [C.c_void_p]),
    "sar_maxpool_planes_fwd": (c_int, [c_fp, c_fp] + [c_int] * 10 + [C.c_void_p]),
    "sar_conv_tc_fwd": (c_int, [C.c_void_p, C.c_void_p]),
    "sar_conv_tc_chain_workspace_bytes": (c_sz, [C.c_void_p, c_int]),
    "sar_conv_tc_chain_fwd": (c_int, [C.c_void_p, c_int, C.c_void_p, c_sz, C.c_void_p]),
    "sar_conv_tc_chain_grid_fwd": (c_int, [C.c_void_p, c_int, C.c_void_p, c_sz, c_int, C.c_void_p]),
    "sar_stem_pool_fwd": (c_int, [c_fp] * 6 + [c_int] * 4 + [C.c_void_p]),
    "sar_maxpool2d_fwd": (c_int, [c_fp, c_fp] + [c_int] * 10 + [C.c_void_p]),
    "sar_affine_relu_fwd": (c_int, [c_fp, c_fp, c_fp, c_fp, c_ll, c_int, c_int, C.c_void_p]),
    "sar_layernorm_fwd": (c_int, [c_fp, c_fp, c_fp, c_fp, c_ll, c_int, c_f, C.c_void_p]),
    "sar_layernorm_planes_fwd": (c_int, [c_fp, c_fp, c_fp, c_fp, c_fp, c_ll, c_int, c_ll, c_int, c_f, C.c_void_p]),
    "sar_bigru_fwd": (c_int, [c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_int, c_int, C.c_void_p]),
    "sar_bigru_nb_fwd": (c_int, [c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_int, c_int, c_int, C.c_void_p]),
    "sar_vlad_fwd": (c_int, [c_fp] * 6 + [c_int] * 5 + [C.c_void_p]),
    "sar_vlad_planes_fwd": (c_int, [c_fp] * 7 + [c_int] * 5 + [C.c_void_p]),
    "sar_vlad_tc_supported": (c_int, [c_int] * 5),
    "sar_vlad_tc_fwd": (c_int, [c_fp, c_ll, c_fp, c_fp, c_fp, c_fp, c_fp] + [c_int] * 5 + [C.c_void_p]),
    "sar_splitk_reduce_fwd": (c_int, [c_fp, c_fp, c_fp, c_int, c_int, c_int, C.c_void_p]),
    "sar_softmax_rows_fwd": (c_int, [c_fp, c_int, c_fp, c_ll, c_int, C.c_void_p]),
    "sar_avgpool_fwd": (c_int, [c_fp, c_fp, c_int, c_int, c_int, C.c_void_p]),
    "sar_gemm_splitk_workspace_bytes": (c_sz, [c_int, c_int, c_int]),
    "sar_gemm_splitk_fwd": (c_int, [c_fp] * 4 + [c_int] * 3 + [C.c_void_p, c_sz, C.c_void_p]),
    "sar_head_fwd": (c_int, [c_fp, c_int, c_fp, c_fp, c_int, c_fp, c_fp, c_int, c_fp, c_fp,
                             c_fp, c_int, c_fp, c_fp, c_int, c_int, c_f, c_f, c_f,
                             c_fp, c_fp, c_fp, c_fp, c_fp, c_int, C.c_void_p]),
    "sar_ctc_fwd": (c_int, [c_fp, c_fp, c_ip, c_ip, c_fp, c_fp, c_ip, c_int, c_int, c_int, c_int, C.c_void_p]),
    "sar_feat_batch_fwd": (c_int, [c_fp, c_ip, c_fp, c_int, c_int, c_int, C.c_void_p]),
    "sar_labels_pack_fwd": (c_int, [c_ip, c_int, c_fp, c_ip, c_ip, c_int, c_int, c_fp, c_ip, c_ip, c_ip, c_int, C.c_void_p]),
    "sar_ctc_greedy_fwd": (c_int, [c_fp, c_int, c_ip, c_int, c_ip, c_ip, c_int, c_int, c_int, C.c_void_p]),
    "sar_ctc_ld_fwd": (c_int, [c_fp, c_int, c_fp, c_ip, c_ip, c_fp, c_fp, c_ip, c_int, c_int, c_int, c_int, C.c_void_p]),
    "sar_ctc_grad_fwd": (c_int, [c_fp, c_int, c_fp, c_ip, c_ip, c_fp, c_fp, c_ip, c_int, c_int, c_int, c_int, C.c_float, C.c_void_p]),
    "sar_loss_reduce_fwd": (c_int, [c_fp, c_fp, c_fp, c_fp, c_int, C.c_void_p]),
    "sar_gemm_fwd": (c_int, [c_fp, c_fp, c_fp, c_int, c_int, c_int, c_int, c_int, c_f, c_f, C.c_void_p]),
    "sar_bn_train_fwd": (c_int, [c_fp] * 8 + [c_int, c_int, c_f, c_f, C.c_void_p]),
    "sar_bn_train_bwd": (c_int, [c_fp] * 8 + [c_int, c_int, C.c_void_p]),
    "sar_bias_act_fwd": (c_int, [c_fp, c_fp, c_fp, c_ll, c_int, c_int, C.c_void_p]),
    "sar_relu_bwd": (c_int, [c_fp, c_fp, c_fp, c_ll, C.c_void_p]),
    "sar_colsum_rows_fwd": (c_int, [c_fp, c_fp, c_int, c_int, c_int, c_fp, C.c_void_p]),
    "sar_bn_train_rows_fwd": (c_int, [c_fp] * 8 + [c_int, c_int, C.c_float, C.c_float, c_int, c_fp, C.c_void_p]),
    "sar_bn_train_rows_bwd": (c_int, [c_fp] * 8 + [c_int, c_int, c_int, c_fp, C.c_void_p]),
    "sar_conv2d_bwd_data": (c_int, [c_fp, c_fp, c_fp] + [c_int] * 12 + [C.c_float, C.c_void_p]),
    "sar_conv2d_bwd_weight": (c_int, [c_fp, c_fp, c_fp] + [c_int] * 13 + [C.c_void_p]),
    "sar_maxpool2d_bwd": (c_int, [c_fp, c_fp, c_fp] + [c_int] * 10 + [C.c_void_p]),
    "sar_axpy_fwd": (c_int, [c_fp, c_fp, C.c_float, c_ll, C.c_void_p]),
    "sar_gru_gate_fwd": (c_int, [c_fp] * 10 + [c_int] * 6 + [C.c_void_p]),
    "sar_gru_gate_bwd": (c_int, [c_fp] * 10 + [c_int] * 6 + [C.c_void_p]),
    "sar_colsum_fwd": (c_int, [c_fp, c_fp, c_int, c_int, C.c_void_p]),
    "sar_l2norm_fwd": (c_int, [c_fp, c_fp, c_fp, c_int, c_int, c_int, C.c_void_p]),
    "sar_l2norm_bwd": (c_int, [c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_int, c_f, C.c_void_p]),
    "sar_head_grad_fwd": (c_int, [c_fp, c_fp, c_fp, c_int, c_int, c_f, c_f, c_f, c_f, c_f, c_fp, c_fp, c_fp, c_int, C.c_void_p]),
    "sar_adam_fwd": (c_int, [c_fp, c_fp, c_fp, c_fp, c_ll, c_f, c_f, c_f, c_f, c_f, C.c_void_p]),
    "sar_adam_dev_fwd": (c_int, [c_fp, c_fp, c_fp, c_fp, c_ll, c_fp, c_f, c_f, c_f, c_f, C.c_void_p]),
    "sar_unit_norm_fwd": (c_int, [c_fp, c_int, c_int, C.c_void_p]),
    "sar_vlad_train_fwd": (c_int, [c_fp] * 7 + [c_int] * 5 + [C.c_void_p]),
    "sar_vlad_train_bwd": (c_int, [c_fp] * 8 + [c_int] * 5 + [C.c_void_p]),
    "sar_ln_train_bwd": (c_int, [c_fp] * 5 + [c_int, c_int, c_f, c_int, C.c_void_p]),
    "sar_fbank_fwd": (c_int, [c_fp, c_ip, c_fp, c_fp, c_fp, c_int, c_int, c_int, C.c_void_p]),
    "sar_fbank_pcm16_fwd": (c_int, [c_ip, c_ip, c_fp, c_fp, c_fp, c_int, c_int, c_int, C.c_void_p]),
}


class SarnetError(RuntimeError):
    pass


def load_library(path: Optional[str] = None):
    """dlopen the C-ABI library and attach prototypes.  Loading does not need a GPU."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise SarnetError(
            "libsarnet_sm100.so not found at %s -- build it with `python -m aesrc2020_b200.csrc.build` "
            "(or __graft_entry__.build()); there is no CPU/PyTorch fallback for this path" % p)
    lib = C.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def lib():
    """Library handle for compute calls: requires a CUDA device (no fallback)."""
    l = load_library()
    if not torch.cuda.is_available():
        raise SarnetError("aesrc2020_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return l


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load_library().sar_last_error()
        raise SarnetError("%s failed (rc=%d): %s" % (what or "sarnet call", rc, msg.decode() if msg else ""))


def ptr(t: Optional[torch.Tensor]):
    """Raw device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise SarnetError("expected a CUDA tensor at the C-ABI boundary, got %s" % t.device)
    if not t.is_contiguous():
        raise SarnetError("tensor crossing the C-ABI boundary must be contiguous")
    return t.data_ptr()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream
