"""Parameter inventory and seeded synthetic data for the disparity hot path.

The key names and shapes are the reference's `state_dict` entries for the modules on the
path (`models/SemStereo.py:204-239`), so a reference checkpoint's
tensors can be handed to `semstereo_b200.hotpath.DisparityHotPath.load_state_dict` unchanged.
ConvTranspose3d weights keep PyTorch's (Cin, Cout, kD, kH, kW) layout.
"""
from __future__ import annotations

from collections import OrderedDict

import torch


def _bn(d, prefix, c):
    d[prefix + ".weight"] = (c,)
    d[prefix + ".bias"] = (c,)
    d[prefix + ".running_mean"] = (c,)
    d[prefix + ".running_var"] = (c,)


def _hourglass(d, hg, c=32):
    """hourglass / hourglass2 (`models/SemStereo.py:106-182`)."""
    for name, ci, co in (("conv1", c, 2 * c), ("conv2", 2 * c, 2 * c), ("conv3", 2 * c, 4 * c), ("conv4", 4 * c, 4 * c)):
        d[f"{hg}.{name}.0.0.weight"] = (co, ci, 3, 3, 3)
        _bn(d, f"{hg}.{name}.0.1", co)
    d[f"{hg}.attention_block.qkv_3d.weight"] = (12 * c, 4 * c)
    d[f"{hg}.attention_block.qkv_3d.bias"] = (12 * c,)
    d[f"{hg}.attention_block.final1x1.weight"] = (4 * c, 4 * c, 1, 1, 1)
    d[f"{hg}.attention_block.final1x1.bias"] = (4 * c,)
    d[f"{hg}.conv5.0.weight"] = (4 * c, 2 * c, 3, 3, 3)
    _bn(d, f"{hg}.conv5.1", 2 * c)
    d[f"{hg}.conv6.0.weight"] = (2 * c, c, 3, 3, 3)
    _bn(d, f"{hg}.conv6.1", c)
    d[f"{hg}.redir1.0.weight"] = (c, c, 1, 1, 1)
    _bn(d, f"{hg}.redir1.1", c)
    d[f"{hg}.redir2.0.weight"] = (2 * c, 2 * c, 1, 1, 1)
    _bn(d, f"{hg}.redir2.1", 2 * c)


def hotpath_param_shapes(num_classes: int = 6) -> "OrderedDict[str, tuple]":
    d: "OrderedDict[str, tuple]" = OrderedDict()
    d["gamma"] = (1,)
    d["beta"] = (1,)
    d["patch.weight"] = (32, 1, 1, 3, 3)
    for pre, cin in (("corr_feature_att_8", 256), ("concat_feature_att_4", 128)):
        d[pre + ".im_att.0.conv.weight"] = (cin // 2, cin, 1, 1)
        _bn(d, pre + ".im_att.0.bn", cin // 2)
        d[pre + ".im_att.1.weight"] = (32, cin // 2, 1, 1)
        d[pre + ".im_att.1.bias"] = (32,)
    _hourglass(d, "hourglass_att")
    _hourglass(d, "hourglass")
    for cl in ("classif_att_", "classif"):
        d[cl + ".0.0.weight"] = (32, 32, 3, 3, 3)
        _bn(d, cl + ".0.1", 32)
        d[cl + ".2.weight"] = (1, 32, 3, 3, 3)
    d["concat_stem.conv.weight"] = (32, 64, 3, 3, 3)
    _bn(d, "concat_stem.bn", 32)
    nc = num_classes
    _bn(d, "ssr_upsample.conv.0", 1)
    d["ssr_upsample.conv.1.weight"] = (nc, 1, 3, 3)
    d["ssr_upsample.conv.1.bias"] = (nc,)
    _bn(d, "ssr_upsample.conv.2", nc)
    for k in ("conv1", "conv2"):
        d[f"ssr_upsample.{k}.0.weight"] = (nc, nc, 1, 1)
        d[f"ssr_upsample.{k}.0.bias"] = (nc,)
        _bn(d, f"ssr_upsample.{k}.1", nc)
    d["ssr_upsample.conv3.weight"] = (1, nc, 1, 1)
    d["ssr_upsample.conv3.bias"] = (1,)
    # concat_feature (SemStereo.py:221-223): the 2-D convs that feed the sparse concat volume -- first step of the widening into
    # SURVEY section 8(f) rank 1.  Appended LAST so the seeded values of every tensor above are unchanged.
    d["concat_feature.0.conv.weight"] = (64, 128, 3, 3)
    _bn(d, "concat_feature.0.bn", 64)
    d["concat_feature.1.weight"] = (32, 64, 3, 3)
    return d


def make_params(seed: int = 1, peaked: float = 1.0, gamma: float = 0.05) -> "OrderedDict[str, torch.Tensor]":
    """Seeded random parameters (CPU fp32).  Conv/linear weights ~ U(-b, b), b = sqrt(3/fan_in)
    (unit-gain, keeps activations O(1) through the stack); BN statistics are deliberately
    non-trivial so that BN folding is exercised.  `peaked` scales the two 32->1 classifier heads:
    with peaked >> 1 the disparity softmaxes are sharp and top-k selection is well conditioned
    (SURVEY.md section 0.7)."""
    g = torch.Generator().manual_seed(seed)
    p: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, shape in hotpath_param_shapes().items():
        if name == "gamma":
            t = torch.full(shape, float(gamma))
        elif name == "beta":
            t = torch.full(shape, 2.0)
        elif name.endswith("running_var"):
            t = torch.rand(shape, generator=g) + 0.5
        elif name.endswith("running_mean"):
            t = 0.1 * torch.randn(shape, generator=g)
        elif len(shape) == 1 and name.endswith(".weight"):       # BN gamma
            t = torch.rand(shape, generator=g) + 0.5
        elif len(shape) == 1:                                     # biases
            t = 0.1 * torch.randn(shape, generator=g)
        else:
            if ".conv5.0." in name or ".conv6.0." in name:       # ConvTranspose3d (Cin,Cout,k,k,k)
                fan_in = shape[0] * 27 / 8.0                      # ~27/8 taps hit each output voxel
            else:
                fan_in = 1
                for s in shape[1:]:
                    fan_in *= s
            b = (3.0 / fan_in) ** 0.5
            t = (torch.rand(shape, generator=g) * 2 - 1) * b
        if name in ("classif_att_.2.weight", "classif.2.weight"):
            t = t * peaked
        p[name] = t.contiguous()
    return p


def make_inputs(seed: int, B: int, H: int, W: int, num_classes: int = 6, max_shift: int = 3):
    """Seeded synthetic inputs of the hot path for a (B,3,H,W) stereo pair (CPU fp32).
    Right features are the left ones displaced along x by a row-block-dependent shift plus noise,
    so the cost volumes have real structure.  H, W must be multiples of 32 (the 1/8-resolution attention branch halves twice and its
    transposed convs double back: H/8 must be a multiple of 4); a size that is not a multiple of 128 makes the hourglasses pad their
    4-wide attention windows at H/16 / H/32 (attention_block's padded branch, submodule_other.py:809-836)."""
    assert H % 32 == 0 and W % 32 == 0, "H and W must be multiples of 32"
    g = torch.Generator().manual_seed(seed)

    def pair(c, h, w, scale):
        l = torch.randn(B, c, h, w, generator=g)
        r = torch.empty_like(l)
        nblk = 4
        for i in range(nblk):
            s = (i % (2 * max_shift + 1)) - max_shift
            ys = slice(i * h // nblk, (i + 1) * h // nblk)
            r[:, :, ys] = torch.roll(l[:, :, ys], shifts=-s * scale, dims=-1)
        r = r + 0.25 * torch.randn(B, c, h, w, generator=g)
        return l.contiguous(), r.contiguous()

    f8_l, f8_r = pair(256, H // 8, W // 8, 1)
    f4_l, f4_r = pair(128, H // 4, W // 4, 2)
    cf_l, cf_r = pair(32, H // 4, W // 4, 2)
    spx = torch.randn(B, num_classes, H, W, generator=g)
    lab = 2.0 * torch.randn(B, num_classes, H, W, generator=g)
    return dict(f8_l=f8_l, f8_r=f8_r, f4_l=f4_l, f4_r=f4_r, cf_l=cf_l, cf_r=cf_r,
                spx_pred=spx.contiguous(), pred_label=lab.contiguous())


# ---------------------------------------------------------------------------------------------------------------------
# 2-D decoder around the path: FeatUp, segmentheads, chal_*, spx* (models/SemStereo.py:59-86, 196-216)
# ---------------------------------------------------------------------------------------------------------------------
BACKBONE_CHANS = (64, 128, 256, 384, 512)      # x2, x4, x8, x16, x32 of `Feature` (SemStereo.py:33-56)
CHANS = (128, 256, 512, 768, 512)              # after FeatUp (SemStereo.py:196)
CHANS2 = (64, 128, 256, 384, 256)              # after chal_* (SemStereo.py:197)


def _conv2x(d, pre, cin, cout):
    """Conv2x(cin, cout, deconv=True, concat=True) (models/submodule.py:119-146)."""
    d[pre + ".conv1.conv.weight"] = (cin, cout, 4, 4)            # ConvTranspose2d layout (Cin, Cout, kH, kW)
    _bn(d, pre + ".conv1.bn", cout)
    d[pre + ".conv2.conv.weight"] = (2 * cout, 2 * cout, 3, 3)
    _bn(d, pre + ".conv2.bn", 2 * cout)


def decoder_param_shapes(num_classes: int = 6) -> "OrderedDict[str, tuple]":
    d: "OrderedDict[str, tuple]" = OrderedDict()
    bc = BACKBONE_CHANS
    _conv2x(d, "feature_up.deconv32_16", bc[4], bc[3])
    _conv2x(d, "feature_up.deconv16_8", bc[3] * 2, bc[2])
    _conv2x(d, "feature_up.deconv8_4", bc[2] * 2, bc[1])
    _conv2x(d, "feature_up.deconv4_2", bc[1] * 2, bc[0])
    for h in ("head_l", "head_r"):
        d[h + ".conv1.conv.weight"] = (CHANS[0] // 4, CHANS[0], 3, 3)
        _bn(d, h + ".conv1.bn", CHANS[0] // 4)
        d[h + ".conv2.weight"] = (num_classes, CHANS[0] // 4, 1, 1)
        d[h + ".conv2.bias"] = (num_classes,)
    d["spx2.0.weight"] = (CHANS2[0] * 2, 6, 4, 4)
    d["spx2.0.bias"] = (6,)
    _conv2x(d, "spx4_2", CHANS2[1] * 2, CHANS2[0])
    _conv2x(d, "spx8_4", CHANS2[2] * 2, CHANS2[1])
    _conv2x(d, "spx16_8", CHANS2[3] * 2, CHANS2[2])
    _conv2x(d, "spx32_16", CHANS2[4], CHANS2[3])
    for i in range(5):
        d[f"chal_{i}.0.weight"] = (CHANS2[i], CHANS[i], 1, 1)
        d[f"chal_{i}.0.bias"] = (CHANS2[i],)
        _bn(d, f"chal_{i}.1", CHANS2[i])
    return d


def make_decoder_params(seed: int = 2) -> "OrderedDict[str, torch.Tensor]":
    """Seeded parameters of the 2-D decoder (same recipe as make_params: unit-gain weights, non-trivial BN statistics)."""
    g = torch.Generator().manual_seed(seed)
    p: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, shape in decoder_param_shapes().items():
        if name.endswith("running_var"):
            t = torch.rand(shape, generator=g) + 0.5
        elif name.endswith("running_mean"):
            t = 0.1 * torch.randn(shape, generator=g)
        elif len(shape) == 1 and name.endswith(".weight"):
            t = torch.rand(shape, generator=g) + 0.5
        elif len(shape) == 1:
            t = 0.1 * torch.randn(shape, generator=g)
        else:
            fan_in = shape[0] * 4 if shape[-1] == 4 else shape[1] * shape[2] * shape[3]      # k4 s2 deconv: 4 taps hit each output
            t = (torch.rand(shape, generator=g) * 2 - 1) * (3.0 / fan_in) ** 0.5
        p[name] = t.contiguous()
    return p


def make_backbone_features(seed: int, B: int, H: int, W: int, max_shift: int = 3):
    """Seeded stand-ins for `Feature` outputs [x2, x4, x8, x16, x32] of the left and right image (CPU fp32): right = left shifted
    along x by a row-block-dependent amount plus noise, like make_inputs."""
    assert H % 64 == 0 and W % 64 == 0, "H and W must be multiples of 64"
    g = torch.Generator().manual_seed(seed)
    fl, fr = [], []
    for c, s in zip(BACKBONE_CHANS, (2, 4, 8, 16, 32)):
        h, w = H // s, W // s
        l = torch.randn(B, c, h, w, generator=g)
        r = torch.empty_like(l)
        for i in range(4):
            sft = ((i % (2 * max_shift + 1)) - max_shift) * max(1, 8 // s)
            ys = slice(i * h // 4, (i + 1) * h // 4)
            r[:, :, ys] = torch.roll(l[:, :, ys], shifts=-sft, dims=-1)
        r = r + 0.25 * torch.randn(B, c, h, w, generator=g)
        fl.append(l.contiguous())
        fr.append(r.contiguous())
    return fl, fr


# ---------------------------------------------------------------------------------------------------------------------
# backbone: MobileViTv2-1.0, HuggingFace parameter names (semstereo_b200/backbone.py)
# ---------------------------------------------------------------------------------------------------------------------
def make_backbone_params(seed: int = 4) -> "OrderedDict[str, torch.Tensor]":
    """Seeded parameters of MobileViTv2Backbone (= a MobileViTV2Model state_dict): unit-gain conv weights, non-trivial
    BatchNorm statistics / GroupNorm affines / biases, so that every folded constant is exercised and activations stay O(1)."""
    from .backbone import MobileViTv2Backbone
    g = torch.Generator().manual_seed(seed)
    p: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, t in MobileViTv2Backbone().state_dict().items():
        shape = tuple(t.shape)
        if name.endswith("num_batches_tracked"):
            continue
        if name.endswith("running_var"):
            v = torch.rand(shape, generator=g) + 0.5
        elif name.endswith("running_mean"):
            v = 0.1 * torch.randn(shape, generator=g)
        elif len(shape) == 1 and name.endswith(".weight"):        # BatchNorm / GroupNorm gamma
            v = torch.rand(shape, generator=g) + 0.5
        elif len(shape) == 1:                                      # biases, betas
            v = 0.1 * torch.randn(shape, generator=g)
        else:                                                      # conv weights (Cout, Cin/groups, k, k)
            fan_in = shape[1] * shape[2] * shape[3]
            v = (torch.rand(shape, generator=g) * 2 - 1) * (3.0 / fan_in) ** 0.5 * 1.4
        p[name] = v.contiguous()
    return p


def make_images(seed: int, B: int, H: int, W: int, max_shift: int = 12):
    """Seeded synthetic stereo pair, ImageNet-normalised like datasets/data_io.py:6-13: smooth random texture, the right image is
    the left one displaced along x by a row-block-dependent shift plus noise (CPU fp32, (B,3,H,W) each)."""
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(B, 3, H // 4, W // 4, generator=g)
    left = torch.nn.functional.interpolate(base, size=(H, W), mode="bilinear", align_corners=False) + 0.1 * torch.rand(B, 3, H, W, generator=g)
    left = left.clamp(0, 1)
    right = torch.empty_like(left)
    for i in range(4):
        ys = slice(i * H // 4, (i + 1) * H // 4)
        right[:, :, ys] = torch.roll(left[:, :, ys], shifts=-((i * 7) % (2 * max_shift + 1) - max_shift), dims=-1)
    right = (right + 0.02 * torch.randn(B, 3, H, W, generator=g)).clamp(0, 1)
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    return ((left - mean) / std).contiguous(), ((right - mean) / std).contiguous()
