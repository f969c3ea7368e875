"""Factory for the reference-facing operator surface.

`semstereo_b200.submodule` (signed, mirrors models/submodule.py) and `semstereo_b200.submodule_` (unsigned,
mirrors models/submodule_.py) are both produced here: same names, argument meaning, return shapes/dtypes and
assertion behaviour as the reference functions they replace, executed by the CUDA kernels of the C-ABI.
Returned tensors are fresh and contiguous, inputs are never mutated (SURVEY.md section 8b "Ownership").
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import hotpath, ops


def _c(t):
    return t.contiguous().float()


def _grad(*ts):
    """True when the call must be recorded by autograd: the differentiable route goes through the torch.library ops of
    torch_ops.py (backward kernels in csrc/backward.cu); inference calls skip the dispatcher."""
    return torch.is_grad_enabled() and any(t.requires_grad for t in ts)


def _T():
    from . import torch_ops  # noqa: F401  (registers torch.ops.semstereo_b200.* on first use)
    return torch.ops.semstereo_b200


def make_surface(signed: bool) -> dict:
    ns: dict = {}

    def dmin(maxdisp):
        return float(-maxdisp) if signed else 0.0

    # ---- volume builders (submodule.py:173-255 / submodule_.py:166-237) -------------------------------
    def _gwc(a, b, maxdisp, groups, sg, norm):
        a, b = _c(a), _c(b)
        return _T().gwc_volume(a, b, maxdisp, groups, sg, norm) if _grad(a, b) else ops.gwc_volume(a, b, maxdisp, groups, sg, norm)

    def build_concat_volume(refimg_fea, targetimg_fea, maxdisp):
        a, b = _c(refimg_fea), _c(targetimg_fea)
        return _T().concat_volume(a, b, maxdisp, signed) if _grad(a, b) else ops.concat_volume(a, b, maxdisp, signed)

    def groupwise_correlation(fea1, fea2, num_groups):
        B, C, H, W = fea1.shape
        assert C % num_groups == 0
        return _gwc(fea1, fea2, 1, num_groups, False, False).squeeze(2)

    def groupwise_correlation_norm(fea1, fea2, num_groups):
        B, C, H, W = fea1.shape
        assert C % num_groups == 0
        return _gwc(fea1, fea2, 1, num_groups, False, True).squeeze(2)

    def norm_correlation(fea1, fea2):
        return _gwc(fea1, fea2, 1, 1, False, True).squeeze(2)

    def build_gwc_volume(refimg_fea, targetimg_fea, maxdisp, num_groups):
        return _gwc(refimg_fea, targetimg_fea, maxdisp, num_groups, signed, False)

    def build_gwc_volume_norm(refimg_fea, targetimg_fea, maxdisp, num_groups):
        return _gwc(refimg_fea, targetimg_fea, maxdisp, num_groups, signed, True)

    def build_norm_correlation_volume(refimg_fea, targetimg_fea, maxdisp):
        return _gwc(refimg_fea, targetimg_fea, maxdisp, 1, signed, True)

    # ---- regression (submodule.py:164-170, 257-263, 434-442) ------------------------------------------
    def disparity_regression(x, maxdisp):
        assert len(x.shape) == 4
        nb = 2 * maxdisp if signed else maxdisp
        if x.shape[1] != nb:
            raise RuntimeError(f"The size of tensor a ({x.shape[1]}) must match the size of tensor b ({nb}) at non-singleton dimension 1")
        x = _c(x)
        return _T().disparity_regression(x, maxdisp, signed) if _grad(x) else ops.disparity_regression(x, dmin(maxdisp))

    def disparity_variance(x, maxdisp, disparity):
        assert len(x.shape) == 4
        nb = 2 * maxdisp if signed else maxdisp
        if x.shape[1] != nb:
            raise RuntimeError(f"The size of tensor a ({x.shape[1]}) must match the size of tensor b ({nb}) at non-singleton dimension 1")
        x, disparity = _c(x), _c(disparity)
        if _grad(x, disparity):
            return _T().disparity_variance(x, maxdisp, disparity, signed)
        return ops.disparity_variance(x, disparity, dmin(maxdisp))

    def regression_topk(cost, disparity_samples, k):
        cost, disparity_samples = _c(cost), _c(disparity_samples)
        if _grad(cost, disparity_samples):
            return _T().regression_topk(cost, disparity_samples, k)
        return ops.regression_topk(cost, disparity_samples, k)

    # ---- warps / propagation (submodule.py:265-307, 361-377) ------------------------------------------
    def SpatialTransformer_grid(x, y, disp_range_samples):
        x, y, disp_range_samples = _c(x), _c(y), _c(disp_range_samples)
        if _grad(x, y, disp_range_samples):
            return _T().spatial_transformer_grid(x, y, disp_range_samples)
        return ops.spatial_transformer_grid(x, y, disp_range_samples, want_x_rep=True)

    class Propagation(nn.Module):
        def forward(self, disparity_samples):
            x = _c(disparity_samples)
            return _T().propagation(x) if _grad(x) else ops.propagation(x)

    class Propagation_prob(nn.Module):
        def forward(self, prob_volume):
            x = _c(prob_volume)
            return _T().propagation(x) if _grad(x) else ops.propagation(x)

    # ---- upsamplers (submodule.py:412-431, submodule_.py:311-323) --------------------------------------
    def context_upsample(depth_low, up_weights):
        depth_low, up_weights = _c(depth_low), _c(up_weights)
        if _grad(depth_low, up_weights):
            return _T().context_upsample(depth_low, up_weights)
        return ops.context_upsample(depth_low, up_weights)

    class SSR_upsample(hotpath._SSRParams):
        def __init__(self, num_classes):
            super().__init__(num_classes)
            self.num_classes = num_classes

        @torch.no_grad()
        def forward(self, depth_low, weights, pred_label):
            if self.training:
                raise NotImplementedError("SSR_upsample (B200 path) folds eval-mode BatchNorm; call .eval()")
            return ops.ssr_upsample(_c(depth_low), _c(weights), _c(pred_label), hotpath.pack_ssr(self))

    # ---- 3-D blocks (submodule_other.py:790-848, submodule.py:89-116) -----------------------------------
    class _ConvBN3d(nn.Sequential):
        """convbn_3d: keys '0' (Conv3d) and '1' (BatchNorm3d); forward = fused conv + folded eval-BN."""

        @torch.no_grad()
        def forward(self, x):
            conv, bn = self[0], self[1]
            if bn.training:
                raise NotImplementedError("convbn_3d (B200 path) folds eval-mode BatchNorm; call .eval()")
            k, s = conv.kernel_size[0], conv.stride[0]
            if conv.kernel_size not in ((1, 1, 1), (3, 3, 3)) or conv.padding != (k // 2,) * 3 or conv.stride not in ((1, 1, 1), (2, 2, 2)):
                raise NotImplementedError("convbn_3d (B200 path): only k in {1,3}, pad=k//2, stride in {1,2}")
            scale, shift = hotpath.bn_affine(bn)
            return ops.conv3d_f32(_c(x), ops.pack_conv3d_weight(conv.weight.detach().float()), scale, shift, k=k, stride=s)

    def convbn_3d(in_planes, out_planes, kernel_size, stride, pad):
        return _ConvBN3d(nn.Conv3d(in_planes, out_planes, kernel_size=kernel_size, padding=pad, stride=stride, bias=False),
                         nn.BatchNorm3d(out_planes))

    class attention_block(hotpath._AttentionParams):
        def __init__(self, channels_3d, num_heads=8, block=4):
            super().__init__(channels_3d)
            self.block, self.num_heads = block, num_heads

        @torch.no_grad()
        def forward(self, x):
            f = self.final1x1
            return ops.window_attention3d(_c(x), self.qkv_3d.weight.detach().float().t().contiguous(), self.qkv_3d.bias.detach().float(),
                                          f.weight.detach().float().reshape(f.out_channels, -1).t().contiguous(), f.bias.detach().float(),
                                          self.block, self.num_heads)

    class BasicConv(nn.Module):
        """BasicConv (submodule.py:89-116).  The 3-D flavours run on the B200 kernels; the 2-D flavours are outside the
        hot path and stay ordinary torch modules."""

        def __init__(self, in_channels, out_channels, deconv=False, is_3d=False, bn=True, relu=True, **kwargs):
            super().__init__()
            self.relu, self.use_bn, self.is_3d, self.deconv = relu, bn, is_3d, deconv
            if is_3d:
                cls = nn.ConvTranspose3d if deconv else nn.Conv3d
                self.conv = cls(in_channels, out_channels, bias=False, **kwargs)
                self.bn = nn.BatchNorm3d(out_channels)
            else:
                cls = nn.ConvTranspose2d if deconv else nn.Conv2d
                self.conv = cls(in_channels, out_channels, bias=False, **kwargs)
                self.bn = nn.BatchNorm2d(out_channels)

        def forward(self, x):
            if not self.is_3d:
                x = self.conv(x)
                if self.use_bn:
                    x = self.bn(x)
                return torch.relu(x) if self.relu else x
            if self.bn.training and self.use_bn:
                raise NotImplementedError("BasicConv 3-D (B200 path) folds eval-mode BatchNorm; call .eval()")
            conv = self.conv
            k, s = conv.kernel_size[0], conv.stride[0]
            scale, shift = hotpath.bn_affine(self.bn) if self.use_bn else (None, None)
            with torch.no_grad():
                if self.deconv:
                    if not (k == 3 and s == 2 and conv.padding == (1, 1, 1) and conv.output_padding == (1, 1, 1)):
                        raise NotImplementedError("BasicConv 3-D deconv (B200 path): only k3 s2 p1 op1")
                    return ops.conv3d_f32(_c(x), ops.pack_conv3d_weight(conv.weight.detach().float(), True), scale, shift,
                                          k=3, stride=2, transposed=True, relu=self.relu)
                if conv.padding != (k // 2,) * 3 or k not in (1, 3) or s not in (1, 2):
                    raise NotImplementedError("BasicConv 3-D (B200 path): only k in {1,3}, pad=k//2, stride in {1,2}")
                return ops.conv3d_f32(_c(x), ops.pack_conv3d_weight(conv.weight.detach().float()), scale, shift, k=k, stride=s,
                                      relu=self.relu)

    for k, v in list(locals().items()):
        if k not in ("ns", "signed", "dmin") and not k.startswith("_"):
            ns[k] = v
    ns["convbn_3d"] = convbn_3d
    return ns
