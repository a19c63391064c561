"""Static description of the SAR-Net hot path: `SAR_Net` keyword set, TF-SAME shape
rules and the ResNet layer plan the device runner and the weight container share.

Reference: model.py:204-224 (kwargs), resnet.py:170-201 (layer order),
utils.py:156-159 / train.py:69 (S(1200,80)=114 shape known-answer).
"""
from __future__ import annotations

from dataclasses import dataclass, field, asdict
from typing import Dict, List, Optional, Tuple

RES_REPS = {"res18": [2, 2, 2, 2], "res34": [3, 4, 6, 3]}
MTO_CHOICES = ("avg", "bigru", "vlad", "gvlad")
METRIC_CHOICES = ("softmax", "sphereface", "cosface", "arcface", "circleloss")


def same_pad(n_in: int, k: int, s: int) -> Tuple[int, int, int]:
    """TF 'SAME' rule: out = ceil(in/s), pad_before = total//2 (extra goes after)."""
    n_out = -(-n_in // s)
    total = max((n_out - 1) * s + k - n_in, 0)
    return n_out, total // 2, total - total // 2


@dataclass
class ConvSpec:
    name: str                 # canonical weight prefix, e.g. "resnet/s2b1/conv1"
    kh: int
    kw: int
    stride: int
    cin: int
    cout: int
    hin: int
    win: int
    hout: int
    wout: int
    pad_t: int
    pad_l: int
    pre_bn: Optional[str] = None      # BN->ReLU applied to this conv's INPUT (resnet.py:47-65)
    post_bn: Optional[str] = None     # BN->ReLU applied to the OUTPUT (stem only, resnet.py:28-45)


@dataclass
class BlockSpec:
    name: str
    conv1: ConvSpec
    conv2: ConvSpec
    short: Optional[ConvSpec]         # 1x1 'valid' projection on the raw block input, or None


@dataclass
class ResNetPlan:
    res_type: str
    filters: int
    T: int
    D: int
    stem: ConvSpec
    pool_hout: int
    pool_wout: int
    pool_pad_t: int
    pool_pad_l: int
    blocks: List[BlockSpec]
    final_bn: str
    hout: int
    wout: int
    cout: int

    @property
    def seq_len(self) -> int:
        return self.hout * self.wout

    def convs(self) -> List[ConvSpec]:
        out = [self.stem]
        for b in self.blocks:
            out += [b.conv1, b.conv2] + ([b.short] if b.short else [])
        return out

    def flops_per_utt(self) -> float:
        return float(sum(2 * c.hout * c.wout * c.cout * c.cin * c.kh * c.kw for c in self.convs()))


def resnet_plan(res_type: str, filters: int, T: int, D: int = 80) -> ResNetPlan:
    """Layer list of resnet18_/resnet34_ (resnet.py:170-201) for a (T, D, 1) input."""
    if res_type in ("res50", "res101", "res152"):
        raise NotImplementedError(
            "%s: the reference's bottleneck ResNets return keras Model objects (resnet.py:217,233,249) "
            "which SAR_Net cannot reshape (model.py:252); only res18/res34 are reachable" % res_type)
    if res_type not in RES_REPS:
        raise ValueError("please specify cnn in res-[18,34,50,101,152]")
    f0 = filters if res_type == "res18" else 64           # resnet.py:173 vs :191
    h1, pt, _ = same_pad(T, 7, 2)
    w1, pl, _ = same_pad(D, 7, 2)
    stem = ConvSpec("resnet/stem", 7, 7, 2, 1, f0, T, D, h1, w1, pt, pl, post_bn="resnet/stem_bn")
    hp, ppt, _ = same_pad(h1, 3, 2)
    wp, ppl, _ = same_pad(w1, 3, 2)
    blocks: List[BlockSpec] = []
    h, w, c, f = hp, wp, f0, filters
    for i, reps in enumerate(RES_REPS[res_type]):
        for j in range(reps):
            name = "resnet/s%db%d" % (i + 1, j + 1)
            stride = 2 if (j == 0 and i != 0) else 1
            ho, pt1, _ = same_pad(h, 3, stride)
            wo, pl1, _ = same_pad(w, 3, stride)
            first = (i == 0 and j == 0)
            c1 = ConvSpec(name + "/conv1", 3, 3, stride, c, f, h, w, ho, wo, pt1, pl1,
                          pre_bn=None if first else name + "/bn1")
            c2 = ConvSpec(name + "/conv2", 3, 3, 1, f, f, ho, wo, ho, wo, 1, 1, pre_bn=name + "/bn2")
            sh = int(round(h / ho))
            sw = int(round(w / wo))
            short = None
            if sh > 1 or sw > 1 or c != f:
                assert sh == sw, "anisotropic shortcut stride is unreachable for (T,80) inputs"
                # 'valid' 1x1 with stride s: out = floor((in-1)/s)+1 == ceil(in/s)
                assert (h - 1) // sh + 1 == ho and (w - 1) // sw + 1 == wo
                short = ConvSpec(name + "/short", 1, 1, sh, c, f, h, w, ho, wo, 0, 0)
            blocks.append(BlockSpec(name, c1, c2, short))
            h, w, c = ho, wo, f
        f *= 2
    return ResNetPlan(res_type, filters, T, D, stem, hp, wp, ppt, ppl, blocks, "resnet/final_bn", h, w, c)


def encoder_len(T: int, D: int = 80, res_type: str = "res34", filters: int = 32) -> int:
    """S = H'*W' after the ResNet; equals the reference's cal_descriptors (utils.py:156-159)."""
    return resnet_plan(res_type, filters, T, D).seq_len


@dataclass
class SARConfig:
    """Mirror of SAR_Net's keyword arguments (model.py:204-224)."""
    input_shape: Tuple[int, int, int] = (1200, 80, 1)
    ctc_enable: bool = False
    ar_enable: bool = True
    disc_enable: bool = False
    res_type: str = "res18"
    res_filters: int = 64
    hidden_dim: int = 256
    bn_dim: int = 0
    bpe_classes: int = 1000
    accent_classes: int = 8
    max_ctc_len: int = 72
    mto: Optional[str] = None
    vlad_clusters: int = 8
    ghost_clusters: int = 2
    metric_loss: str = "cosface"
    margin: float = 0.3

    def model_kwargs(self) -> Dict:
        d = asdict(self)
        d.pop("input_shape")
        return d

    def plan(self) -> ResNetPlan:
        return resnet_plan(self.res_type, self.res_filters, int(self.input_shape[0]), int(self.input_shape[1]))

    def integration_dim(self) -> int:
        if self.mto == "avg":
            return self.hidden_dim
        if self.mto == "bigru":
            return 2 * self.hidden_dim
        if self.mto in ("vlad", "gvlad"):
            return self.vlad_clusters * self.hidden_dim
        raise ValueError("Please specify avg/bigru/vlad/gvlad ..")

    def input_names(self) -> List[str]:
        names = ["x_data"]                      # model.py:327-335
        if self.disc_enable:
            names.append("x_accent")
        if self.ctc_enable:
            names += ["x_ctc_label", "x_ctc_in_len", "x_ctc_out_len"]
        return names

    def output_names(self) -> List[str]:
        names = []                              # model.py:328-338
        if self.ar_enable:
            names.append("y_accent")
        if self.disc_enable:
            names.append("y_disc")
        if self.ctc_enable:
            names.append("y_ctc_loss")
        if self.bn_dim:
            names.append("y_disc_bn")
        return names

    def loss_weights(self) -> Dict[str, float]:
        """model.py:344-367, including the double assignment at :360-361 (second wins)."""
        alpha, beta = 0.4, 0.01
        w: Dict[str, float] = {}
        if self.ar_enable:
            w["y_accent"] = beta if self.disc_enable else 1.0
            if self.disc_enable:
                w["y_disc"] = 1 - alpha if self.ctc_enable else 1.0
        if self.ctc_enable:
            w["y_ctc_loss"] = alpha if self.disc_enable else 1.0
            w["y_ctc_loss"] = 1 - alpha if not self.disc_enable else beta
        if self.bn_dim:
            w["y_disc_bn"] = 0.1
        return w
