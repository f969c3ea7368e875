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


_PRECISION = "fp32"


def set_precision(mode: str) -> None:
    """Arithmetic of the 3-D modules of the surface (convbn_3d, BasicConv(is_3d), attention_block):
    "fp32" (default) = fp32 FFMA kernels, the parity mode; "bf16" = tcgen05 tensor cores with bf16 operands / fp32 accumulation
    for every layer geometry that has a tensor-core configuration (others stay fp32).  The free functions are always fp32."""
    global _PRECISION
    if mode not in ("fp32", "bf16"):
        raise ValueError("surface precision must be 'fp32' or 'bf16'")
    _PRECISION = mode


def _c(t):
    return t.contiguous().float()


class _Packed:
    """Per-module cache of folded / packed weights.  Keyed by the identity and in-place version of every tensor it was built
    from, so load_state_dict (in-place copy -> version bump), .to()/.cuda() (new storage) and optimizer steps all invalidate it."""

    def __init__(self):
        self.key, self.val = None, {}

    def get(self, tensors, name, build):
        key = tuple((t.data_ptr(), t._version, t.device) for t in tensors)
        if key != self.key:
            self.key, self.val = key, {}
        if name not in self.val:
            self.val[name] = build()
        return self.val[name]


def _no_grad_path(mod, *inputs):
    """The fused inference kernels record no autograd graph: refuse instead of silently returning zero gradients."""
    if torch.is_grad_enabled() and (any(t.requires_grad for t in inputs) or (mod.training and any(p.requires_grad for p in mod.parameters()))):
        raise NotImplementedError(f"{type(mod).__name__} (B200 path) is inference-only: call it under torch.no_grad() / .eval() "
                                  "(the training closure is BASELINE config #5, see DESIGN.md)")


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

        def forward(self, depth_low, weights, pred_label):
            if self.training:      # batch statistics in the four BatchNorm2d layers: differentiable composition (train_ops.py)
                from . import train_ops
                return train_ops.ssr_upsample_train(self, depth_low, weights, pred_label)
            _no_grad_path(self, depth_low, weights, pred_label)
            if not hasattr(self, "_packed"):
                object.__setattr__(self, "_packed", _Packed())
            packed = self._packed.get(list(self.parameters()) + list(self.buffers()), "ssr", lambda: hotpath.pack_ssr(self))   # one D2H per weight change
            with torch.no_grad():
                return ops.ssr_upsample(_c(depth_low), _c(weights), _c(pred_label), packed)

    # ---- 3-D blocks (submodule_other.py:790-848, submodule.py:89-116) -----------------------------------
    def _conv3d(mod, conv, bn, x, relu, transposed=False):
        """Conv3d / ConvTranspose3d (+ folded eval BatchNorm, + ReLU) of a surface module through the kernels; packed weights are
        cached on the module.  set_precision("bf16") takes the tcgen05 route where the geometry has a configuration."""
        from . import ops_tc as tc
        if not hasattr(mod, "_packed"):
            object.__setattr__(mod, "_packed", _Packed())
        src = [conv.weight] + ([bn.weight, bn.bias, bn.running_mean, bn.running_var] if bn is not None else [])
        k, st = conv.kernel_size[0], conv.stride[0]
        if transposed:
            if not (conv.kernel_size == (3, 3, 3) and conv.stride == (2, 2, 2) and conv.padding == (1, 1, 1) and conv.output_padding == (1, 1, 1)):
                raise NotImplementedError("3-D deconv (B200 path): only k3 s2 p1 op1")
            cin, cout = conv.weight.shape[:2]
        else:
            if conv.kernel_size not in ((1, 1, 1), (3, 3, 3)) or conv.padding != (k // 2,) * 3 or conv.stride not in ((1, 1, 1), (2, 2, 2)):
                raise NotImplementedError("3-D conv (B200 path): only k in {1,3}, pad=k//2, stride in {1,2}")
            cout, cin = conv.weight.shape[:2]
        scale, shift = mod._packed.get(src, "affine", lambda: hotpath.bn_affine(bn)) if bn is not None else (None, None)
        x = _c(x)
        kind = tc.T2 if transposed else (tc.K1 if k == 1 else (tc.S2 if st == 2 else tc.S1))
        if _PRECISION == "bf16" and cin % 8 == 0 and tc.ntile(kind, cin, cout) and (kind != tc.S2 or all(v % 2 == 0 for v in x.shape[2:])):
            w = mod._packed.get(src, "tc", lambda: tc.pack_weight(conv.weight.detach().float(), kind))
            return tc.conv3d_tc(kind, tc.to_blocked_bf16(x, s2d=(kind == tc.S2)), w, cout, scale, shift, relu=relu, out_mode=tc.F32)
        w = mod._packed.get(src, "f32", lambda: ops.pack_conv3d_weight(conv.weight.detach().float(), transposed))
        return ops.conv3d_f32(x, w, scale, shift, k=k, stride=st, transposed=transposed, relu=relu)

    class _ConvBN3d(nn.Sequential):
        """convbn_3d: keys '0' (Conv3d) and '1' (BatchNorm3d); forward = fused conv + folded eval-BN."""

        def forward(self, x):
            conv, bn = self[0], self[1]
            if bn.training:      # training mode (BASELINE config #5): Conv3d and BatchNorm3d with batch statistics, forward and
                from . import train_ops      # backward on the CUDA kernels (csrc/train.cu); differentiable, running stats updated
                return train_ops.batch_norm_train(train_ops.conv3d(x, conv.weight, conv.stride[0]), bn)
            _no_grad_path(self, x)
            with torch.no_grad():
                return _conv3d(self, conv, bn, x, relu=False)

    def convbn_3d(in_planes, out_planes, kernel_size, stride, pad):
        return _ConvBN3d(nn.Conv3d(in_planes, out_planes, kernel_size=kernel_size, padding=pad, stride=stride, bias=False),
                         nn.BatchNorm3d(out_planes))

    class attention_block(hotpath._AttentionParams):
        def __init__(self, channels_3d, num_heads=8, block=4):
            super().__init__(channels_3d)
            self.block, self.num_heads = block, num_heads

        def forward(self, x):
            if self.training:      # differentiable route (train_ops.py): k = 1 convs + the fp32 softmax core with its backward kernel
                from . import train_ops
                return train_ops.attention_block_train(self, x)
            _no_grad_path(self, x)
            if not hasattr(self, "_packed"):
                object.__setattr__(self, "_packed", _Packed())
            f = self.final1x1
            src = [self.qkv_3d.weight, self.qkv_3d.bias, f.weight, f.bias]
            wq, bq, wo, bo = self._packed.get(src, "att", lambda: (
                self.qkv_3d.weight.detach().float().t().contiguous(), self.qkv_3d.bias.detach().float().contiguous(),
                f.weight.detach().float().reshape(f.out_channels, -1).t().contiguous(), f.bias.detach().float().contiguous()))
            block = self.block if isinstance(self.block, (tuple, list)) else (self.block,) * 3
            with torch.no_grad():
                return ops.window_attention3d(_c(x), wq, bq, wo, bo, block, self.num_heads)

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
            if self.training and not self.deconv:      # training mode: native Conv3d / BatchNorm3d forward + backward (csrc/train.cu)
                from . import train_ops
                y = train_ops.conv3d(x, self.conv.weight, self.conv.stride[0])
                if self.use_bn:
                    y = train_ops.batch_norm_train(y, self.bn)
                return torch.relu(y) if self.relu else y
            if self.training:
                raise NotImplementedError("BasicConv 3-D deconv (B200 path) has no training mode; the models use nn.ConvTranspose3d directly")
            _no_grad_path(self, x)
            with torch.no_grad():
                return _conv3d(self, self.conv, self.bn if self.use_bn else None, x, relu=self.relu, transposed=self.deconv)

    # ---- 2-D modules the reference models take from the same file (SemStereo.py:59-86, 200-211 use Conv2x and segmenthead;
    # Fusion, DWConv2d/3d, Propagation2, Propagation_prob2, ConvSelfAttention are defined but never instantiated).  They are
    # outside the hot path (the accelerated decoder is semstereo_b200.decoder / patch_model) and are ordinary
    # torch modules here, with the reference's constructor signatures, state_dict keys and forward semantics, so that
    # `from models.submodule import *` gives SemStereo.py / SemStereo_WHU.py every name they use. ----
    class segmenthead(nn.Module):
        """models/submodule.py:31-52: BasicConv 3x3 -> Conv2d 1x1 (+bias) -> optional bilinear upsampling by scale_factor."""

        def __init__(self, inplanes, interplanes, outplanes, scale_factor=None):
            super().__init__()
            self.conv1 = BasicConv(inplanes, interplanes, kernel_size=3, padding=1)
            self.conv2 = nn.Conv2d(interplanes, outplanes, kernel_size=1, padding=0, bias=True)
            self.scale_factor = scale_factor

        def forward(self, x):
            x = self.conv1(x)
            out = self.conv2(x)
            if self.scale_factor is not None:
                size = [x.shape[-2] * self.scale_factor, x.shape[-1] * self.scale_factor]
                out = nn.functional.interpolate(out, size=size, mode="bilinear", align_corners=False)
            return out

    class Conv2x(nn.Module):
        """models/submodule.py:119-161: stride-2 (de)conv, then concat with (or add to) the skip tensor, then a 3x3 conv."""

        def __init__(self, in_channels, out_channels, deconv=False, is_3d=False, concat=True, keep_concat=True, bn=True, relu=True,
                     keep_dispc=False):
            super().__init__()
            self.concat, self.is_3d = concat, is_3d
            if deconv and is_3d and keep_dispc:
                geo = dict(kernel_size=(1, 4, 4), stride=(1, 2, 2), padding=(0, 1, 1))
            else:
                geo = dict(kernel_size=((4, 4, 4) if is_3d else 4) if deconv else 3, stride=2, padding=1)
            self.conv1 = BasicConv(in_channels, out_channels, deconv, is_3d, bn=True, relu=True, **geo)
            c2_in = out_channels * 2 if concat else out_channels
            c2_out = out_channels * (2 if keep_concat else 1) if concat else out_channels
            self.conv2 = BasicConv(c2_in, c2_out, False, is_3d, bn, relu, kernel_size=3, stride=1, padding=1)

        def forward(self, x, rem):
            x = self.conv1(x)
            if x.shape != rem.shape:
                x = nn.functional.interpolate(x, size=(rem.shape[-2], rem.shape[-1]), mode="bilinear")
            x = torch.cat((x, rem), 1) if self.concat else x + rem
            return self.conv2(x)

    class Fusion(nn.Module):
        """models/submodule.py:10-29 (never instantiated by the models)."""

        def __init__(self, inplanes, interplanes_x, interplanes_y, mid_channels):
            super().__init__()
            half = interplanes_x // 2
            self.conv_y = nn.Sequential(nn.Conv2d(interplanes_y, half, kernel_size=1, padding=0, bias=True), nn.BatchNorm2d(half))
            self.conv_x = nn.Sequential(nn.Conv2d(interplanes_x, half, kernel_size=1, padding=0, bias=True), nn.BatchNorm2d(half))
            self.conv = BasicConv(interplanes_x, interplanes_x, kernel_size=3, stride=1, padding=1)

        def forward(self, x, y):
            size = x.shape[-2:]
            y, x = self.conv_y(y), self.conv_x(x)
            if y.shape != x.shape:
                y = nn.functional.interpolate(y, tuple(size), mode="bilinear", align_corners=False)
            w = torch.softmax(self.conv(torch.cat((x, y), 1)), dim=1) + 1
            return x * w, y * w

    def _dwconv(conv_cls, bn_cls):
        class DW(nn.Module):
            def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1):
                super().__init__()
                self.depthwise = conv_cls(in_channels, in_channels, kernel_size=kernel_size, stride=stride, padding=padding, groups=in_channels)
                self.bn1 = bn_cls(in_channels)
                self.relu = nn.ReLU(inplace=True)
                self.pointwise = conv_cls(in_channels, out_channels, kernel_size=1, stride=1, padding=0)
                self.bn2 = bn_cls(out_channels)

            def forward(self, x):
                return self.bn2(self.pointwise(self.relu(self.bn1(self.depthwise(x)))))
        return DW

    DWConv2d = _dwconv(nn.Conv2d, nn.BatchNorm2d)     # models/submodule.py:54-70
    DWConv3d = _dwconv(nn.Conv3d, nn.BatchNorm3d)     # models/submodule.py:72-87
    DWConv2d.__name__, DWConv3d.__name__ = "DWConv2d", "DWConv3d"

    def _one_hot_taps(taps, radius, ndim):
        class P(nn.Module):
            """One-hot 5x5 tap gather with replicate padding (models/submodule.py:309-359; never instantiated by the models)."""

            def forward(self, x):
                k = 2 * radius + 1
                f = torch.zeros((len(taps), 1) + ((1, k, k) if ndim == 3 else (k, k)), device=x.device)
                for i, (r, c) in enumerate(taps):
                    f[(i, 0, 0, r, c) if ndim == 3 else (i, 0, r, c)] = 1.0
                if ndim == 3:
                    return nn.functional.conv3d(nn.functional.pad(x, (radius,) * 4 + (0, 0), mode="replicate"), f)
                return nn.functional.conv2d(nn.functional.pad(x, (radius,) * 4, mode="replicate"), f)
        return P

    _TAPS9 = ((0, 2), (2, 0), (2, 2), (2, 4), (4, 2), (1, 1), (3, 3), (1, 3), (3, 1))
    Propagation2 = _one_hot_taps(_TAPS9, 2, 2)                                   # models/submodule.py:309-332
    # Propagation_prob2 (:334-359) sets taps 0-4 twice; the later assignments ADD ones (the earlier stay set): reproduce both
    class Propagation_prob2(nn.Module):
        def forward(self, prob_volume):
            f = torch.zeros(9, 1, 1, 5, 5, device=prob_volume.device)
            for i, (r, c) in enumerate(((0, 0), (1, 1), (2, 2), (2, 0), (0, 2))):
                f[i, 0, 0, r, c] = 1.0
            for i, (r, c) in enumerate(_TAPS9):
                f[i, 0, 0, r, c] = 1.0
            return nn.functional.conv3d(nn.functional.pad(prob_volume, (2, 2, 2, 2, 0, 0), mode="replicate"), f)
    Propagation2.__name__ = "Propagation2"

    class ConvSelfAttention(nn.Module):
        """models/submodule.py:379-410 (never instantiated by the models): global dot-product self-attention over pixels."""

        def __init__(self, in_channels, out_channels):
            super().__init__()
            self.query_conv = nn.Conv2d(in_channels, out_channels, kernel_size=1, padding=0)
            self.key_conv = nn.Conv2d(in_channels, out_channels, kernel_size=1, padding=0)
            self.value_conv = nn.Conv2d(in_channels, out_channels, kernel_size=1, padding=0)
            self.gamma = nn.Parameter(torch.zeros(1))

        def forward(self, x):
            q, k, v = self.query_conv(x), self.key_conv(x), self.value_conv(x)
            b, c, h, w = q.shape
            att = torch.softmax(torch.bmm(q.view(b, c, -1).permute(0, 2, 1), k.view(b, c, -1)), dim=2)
            return self.gamma * torch.bmm(att, v.view(b, c, -1)).view(b, c, h, w) + x

    for k, v in list(locals().items()):
        if k not in ("ns", "signed", "dmin") and not k.startswith("_"):
            ns[k] = v
    ns["convbn_3d"] = convbn_3d
    return ns
