"""oracle/decoder.py — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU restatement (torch fp32, functional, eval-mode BatchNorm) of the 2-D decoder the reference wraps around the disparity
path: FeatUp (models/SemStereo.py:59-86), segmenthead (models/submodule.py:31-52), chal_* and the spx chain
(models/SemStereo.py:207-216, 256-271).  `p` is keyed by the reference's state_dict names.  Pinned against the reference's own
modules by oracle/make_golden_decoder.py -> tests/golden/decoder_us3d.npz.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .ops import _bn


def basic_conv2d(x, p, pre, deconv=False, relu=True, stride=1, pad=1):
    """BasicConv, 2-D flavour (models/submodule.py:89-116): conv (no bias) -> BN -> ReLU."""
    w = p[pre + ".conv.weight"]
    y = F.conv_transpose2d(x, w, None, stride=stride, padding=pad) if deconv else F.conv2d(x, w, None, stride=stride, padding=pad)
    y = _bn(y, p, pre + ".bn")
    return F.relu(y) if relu else y


def conv2x(x, rem, p, pre):
    """Conv2x(deconv=True, concat=True) (models/submodule.py:119-161): ConvTranspose2d k4 s2 p1 + BN + ReLU, cat with the skip,
    Conv2d 3x3 + BN + ReLU.  (The bilinear resize at :150-154 only triggers for sizes that are not multiples of 32.)"""
    x = basic_conv2d(x, p, pre + ".conv1", deconv=True, stride=2, pad=1)
    assert x.shape == rem.shape, "input sizes must be multiples of 32"
    return basic_conv2d(torch.cat((x, rem), 1), p, pre + ".conv2")


def feat_up(p, feat):
    """FeatUp.forward for one image (models/SemStereo.py:70-86; left and right share the weights)."""
    x2, x4, x8, x16, x32 = feat
    x16 = conv2x(x32, x16, p, "feature_up.deconv32_16")
    x8 = conv2x(x16, x8, p, "feature_up.deconv16_8")
    x4 = conv2x(x8, x4, p, "feature_up.deconv8_4")
    x2 = conv2x(x4, x2, p, "feature_up.deconv4_2")
    return [x2, x4, x8, x16, x32]


def segmenthead(x, p, pre):
    """segmenthead(scale_factor=2) (models/submodule.py:31-52)."""
    y = basic_conv2d(x, p, pre + ".conv1")
    y = F.conv2d(y, p[pre + ".conv2.weight"], p[pre + ".conv2.bias"])
    return F.interpolate(y, size=[y.shape[-2] * 2, y.shape[-1] * 2], mode="bilinear", align_corners=False)


def chal(x, p, i):
    """chal_i = Conv2d 1x1 (bias) + BatchNorm2d (models/SemStereo.py:207-211)."""
    return _bn(F.conv2d(x, p[f"chal_{i}.0.weight"], p[f"chal_{i}.0.bias"]), p, f"chal_{i}.1")


@torch.no_grad()
def forward(p, feat_l, feat_r, right_label=True):
    """SemStereo.forward:246-271: backbone pyramids -> what the disparity path consumes."""
    fl, fr = feat_up(p, feat_l), feat_up(p, feat_r)
    out = {"pred_label": segmenthead(fl[0], p, "head_l")}
    if right_label:
        out["pred_label_r"] = segmenthead(fr[0], p, "head_r")
    cl = [chal(fl[i], p, i) for i in range(5)]
    out.update(f4_l=cl[1], f8_l=cl[2], f4_r=chal(fr[1], p, 1), f8_r=chal(fr[2], p, 2))
    x = conv2x(cl[4], cl[3], p, "spx32_16")
    x = conv2x(x, cl[2], p, "spx16_8")
    x = conv2x(x, cl[1], p, "spx8_4")
    x = conv2x(x, cl[0], p, "spx4_2")
    out["spx_pred"] = F.conv_transpose2d(x, p["spx2.0.weight"], p["spx2.0.bias"], stride=2, padding=1)
    return out
