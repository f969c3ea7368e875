"""MobileViTv2-1.0 backbone on the B200 kernels -- the reference's `Feature` (models/SemStereo.py:33-56), which is timm's
`mobilevitv2_100` with `features_only=True`: five maps [x2 (64 ch), x4 (128), x8 (256), x16 (384), x32 (512)] of an image.
SURVEY.md section 8(f) rank 2.  timm is not in this image; the architecture is the published MobileViTv2 (Mehta & Rastegari,
"Separable Self-attention for Mobile Vision Transformers", 2022) as implemented in HuggingFace `transformers`
(`models/mobilevitv2/modeling_mobilevitv2.py`, v5.5), which IS installed and serves as the oracle (oracle/backbone.py):

  stem  Conv 3x3 s2 (3->32) BN SiLU
  L1    InvertedResidual(32->64, s1)                           -> x2      IR = 1x1 expand(x2) BN SiLU, dw 3x3 BN SiLU, 1x1 BN (+skip)
  L2    IR(64->128, s2), IR(128->128, s1)                      -> x4
  L3-5  IR(s2) to C = 256 / 384 / 512, then a MobileViTv2 block -> x8, x16, x32
        block: dw 3x3 BN SiLU, 1x1 (C->d), N x [GroupNorm(1), separable attention, +skip, GroupNorm(1), FFN d->2d->d, +skip],
               GroupNorm(1), 1x1 (d->C) BN;  d = 128 / 192 / 256, N = 2 / 4 / 3

The parameter containers carry the module names of the HuggingFace implementation, so `MobileViTV2Model(...).state_dict()`
loads unchanged (a timm checkpoint needs a key rename: its stem / stages_0..4 hold the same tensors under timm's names --
not verifiable here, timm is absent).  Containers are storage only; `forward` issues the kernels: every 1x1 convolution on
the tcgen05 tensor cores (ss_conv2d_tc_ex: bf16 operands, fp32 accumulation, folded BN / bias, SiLU and the skip connection in
the epilogue), the rest in csrc/backbone.cu.  Activations are bf16 blocked (B, C/8, H, W, 8) throughout; the `unfold`/`fold`
of the reference block never happens (see csrc/backbone.cu).  Inference only, CUDA only, no fallback.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from . import ops_tc as tc
from .hotpath import bn_affine

STAGE_CHANS = (64, 128, 256, 384, 512)
ATTN_DIMS = (128, 192, 256)
ATTN_BLOCKS = (2, 4, 3)


class _ConvLayer(nn.Module):
    """MobileViTV2ConvLayer: `convolution` (+ `normalization`)."""

    def __init__(self, cin, cout, k, stride=1, groups=1, bias=False, norm=True):
        super().__init__()
        self.convolution = nn.Conv2d(cin, cout, k, stride, (k - 1) // 2, groups=groups, bias=bias)
        if norm:
            self.normalization = nn.BatchNorm2d(cout)


class _InvertedResidual(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        mid = max(8, int(cin * 2 + 4) // 8 * 8)                  # make_divisible(round(cin * expand_ratio 2.0), 8)
        self.stride, self.use_residual = stride, (stride == 1 and cin == cout)
        self.expand_1x1 = _ConvLayer(cin, mid, 1)
        self.conv_3x3 = _ConvLayer(mid, mid, 3, stride, groups=mid)
        self.reduce_1x1 = _ConvLayer(mid, cout, 1)


class _MobileNetLayer(nn.Module):
    def __init__(self, cin, cout, stride, n):
        super().__init__()
        self.layer = nn.ModuleList(_InvertedResidual(cin if i == 0 else cout, cout, stride if i == 0 else 1) for i in range(n))


class _Attention(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.qkv_proj = _ConvLayer(d, 1 + 2 * d, 1, bias=True, norm=False)
        self.out_proj = _ConvLayer(d, d, 1, bias=True, norm=False)


class _FFN(nn.Module):
    def __init__(self, d, latent):
        super().__init__()
        self.conv1 = _ConvLayer(d, latent, 1, bias=True, norm=False)
        self.conv2 = _ConvLayer(latent, d, 1, bias=True, norm=False)


class _TransformerLayer(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.layernorm_before = nn.GroupNorm(1, d, eps=1e-5)
        self.attention = _Attention(d)
        self.layernorm_after = nn.GroupNorm(1, d, eps=1e-5)
        self.ffn = _FFN(d, (2 * d // 16) * 16)


class _Transformer(nn.Module):
    def __init__(self, d, n):
        super().__init__()
        self.layer = nn.ModuleList(_TransformerLayer(d) for _ in range(n))


class _ViTLayer(nn.Module):
    def __init__(self, cin, cout, d, n):
        super().__init__()
        self.d = d
        self.downsampling_layer = _InvertedResidual(cin, cout, 2)
        self.conv_kxk = _ConvLayer(cout, cout, 3, groups=cout)
        self.conv_1x1 = _ConvLayer(cout, d, 1, norm=False)
        self.transformer = _Transformer(d, n)
        self.layernorm = nn.GroupNorm(1, d, eps=1e-5)
        self.conv_projection = _ConvLayer(d, cout, 1)


class _Encoder(nn.Module):
    def __init__(self):
        super().__init__()
        c = STAGE_CHANS
        self.layer = nn.ModuleList([_MobileNetLayer(32, c[0], 1, 1), _MobileNetLayer(c[0], c[1], 2, 2)] +
                                   [_ViTLayer(c[i + 1], c[i + 2], ATTN_DIMS[i], ATTN_BLOCKS[i]) for i in range(3)])


class MobileViTv2Backbone(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv_stem = _ConvLayer(3, 32, 3, 2)
        self.encoder = _Encoder()
        self._cache = None
        self.eval()
        for p in self.parameters():
            p.requires_grad_(False)

    # ------------------------------------------------------------------------------------------
    def load_state_dict(self, state_dict, strict=False, **kw):
        """Accepts a HuggingFace MobileViTV2Model state_dict (optionally prefixed 'mobilevitv2.' / 'module.' / 'feature.');
        foreign keys are ignored."""
        own = self.state_dict()
        sd = {}
        for k, v in state_dict.items():
            for pre in ("module.", "feature.", "mobilevitv2."):
                if k.startswith(pre):
                    k = k[len(pre):]
            if k in own:
                sd[k] = v
        missing = [k for k in own if k not in sd and not k.endswith("num_batches_tracked")]
        if strict and missing:
            raise KeyError(f"missing backbone keys: {missing[:5]} ...")
        out = super().load_state_dict(sd, strict=False, **kw)
        self._cache = None
        return out

    def _apply(self, fn, *a, **k):
        self._cache = None
        return super()._apply(fn, *a, **k)

    def refresh(self):
        self._cache = None

    # ------------------------------------------------------------------------------------------
    def _packed(self):
        if self._cache is not None:
            return self._cache
        if self.training:
            raise NotImplementedError("MobileViTv2Backbone is inference-only (eval-mode BatchNorm is folded)")
        c = {}

        def pw(name, m, cin_pad=None, rows=None):
            """1x1 conv (+BN or bias) -> CONV1 packing; cin_pad zero-pads the input channels; rows reorders / pads the outputs."""
            w = m.convolution.weight.detach().float().reshape(m.convolution.out_channels, -1)
            b = m.convolution.bias.detach().float() if m.convolution.bias is not None else None
            if rows is not None:
                w2 = w.new_zeros((len(rows), w.shape[1]))
                b2 = w.new_zeros(len(rows))
                for i, r in enumerate(rows):
                    if r >= 0:
                        w2[i] = w[r]
                        if b is not None:
                            b2[i] = b[r]
                w, b = w2, (b2 if b is not None else None)
            if cin_pad is not None and cin_pad > w.shape[1]:
                w = torch.cat((w, w.new_zeros((w.shape[0], cin_pad - w.shape[1]))), 1)
            c[name + ".w"] = tc.pack_weight2d(w.reshape(w.shape[0], w.shape[1], 1, 1), tc.CONV1)
            if hasattr(m, "normalization"):
                c[name + ".s"], c[name + ".t"] = bn_affine(m.normalization)
            else:
                c[name + ".s"], c[name + ".t"] = None, (b.contiguous() if b is not None else None)
            c[name + ".cout"] = w.shape[0]

        def dw(name, m):
            c[name + ".w"] = m.convolution.weight.detach().float().reshape(m.convolution.out_channels, 9).contiguous()
            c[name + ".s"], c[name + ".t"] = bn_affine(m.normalization)

        def ir(name, m, cin_pad=None):
            pw(name + ".expand_1x1", m.expand_1x1, cin_pad)
            dw(name + ".conv_3x3", m.conv_3x3)
            pw(name + ".reduce_1x1", m.reduce_1x1)

        st = self.conv_stem
        c["stem.w"] = st.convolution.weight.detach().float().contiguous()
        c["stem.s"], c["stem.t"] = bn_affine(st.normalization)
        enc = self.encoder.layer
        ir("L0.0", enc[0].layer[0], cin_pad=64)                  # the stem output is zero-padded from 32 to 64 channels
        ir("L1.0", enc[1].layer[0])
        ir("L1.1", enc[1].layer[1])
        for li in (2, 3, 4):
            m, n, d = enc[li], f"L{li}", enc[li].d
            ir(n + ".down", m.downsampling_layer)
            dw(n + ".conv_kxk", m.conv_kxk)
            pw(n + ".conv_1x1", m.conv_1x1)
            for i, t in enumerate(m.transformer.layer):
                tn = f"{n}.t{i}"
                for g, mod in (("ln1", t.layernorm_before), ("ln2", t.layernorm_after)):
                    c[f"{tn}.{g}.g"], c[f"{tn}.{g}.b"] = mod.weight.detach().float().contiguous(), mod.bias.detach().float().contiguous()
                # qkv_proj rows are [query | key (d) | value (d)]; the attention kernel wants [key | value | query, 7 x zero]
                pw(tn + ".qkv", t.attention.qkv_proj, rows=list(range(1, 1 + 2 * d)) + [0] + [-1] * 7)
                pw(tn + ".out", t.attention.out_proj)
                pw(tn + ".ffn1", t.ffn.conv1)
                pw(tn + ".ffn2", t.ffn.conv2)
            c[n + ".ln.g"], c[n + ".ln.b"] = m.layernorm.weight.detach().float().contiguous(), m.layernorm.bias.detach().float().contiguous()
            pw(n + ".proj", m.conv_projection)
        self._cache = c
        return c

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _pw(c, name, x, act=tc.NONE, residual=None):
        return tc.conv2d_tc(tc.CONV1, x, c[name + ".w"], c[name + ".cout"], c[name + ".s"], c[name + ".t"], act=act, residual=residual)

    def _ir(self, c, name, m, x):
        with ops.label(name):
            h = self._pw(c, name + ".expand_1x1", x, tc.SILU)
            h = tc.dwconv3x3(h, c[name + ".conv_3x3.w"], c[name + ".conv_3x3.s"], c[name + ".conv_3x3.t"], m.stride, tc.SILU)
            return self._pw(c, name + ".reduce_1x1", h, tc.NONE, x if m.use_residual else None)

    def _vit(self, c, name, m, x):
        x = self._ir(c, name + ".down", m.downsampling_layer, x)
        d = m.d
        with ops.label(name + ".block"):
            h = tc.dwconv3x3(x, c[name + ".conv_kxk.w"], c[name + ".conv_kxk.s"], c[name + ".conv_kxk.t"], 1, tc.SILU)
            h = self._pw(c, name + ".conv_1x1", h)
            for i in range(len(m.transformer.layer)):
                tn = f"{name}.t{i}"
                n = tc.groupnorm1(h, c[tn + ".ln1.g"], c[tn + ".ln1.b"])
                a = tc.linear_attention(self._pw(c, tn + ".qkv", n), d)
                h = self._pw(c, tn + ".out", a, tc.NONE, h)
                n = tc.groupnorm1(h, c[tn + ".ln2.g"], c[tn + ".ln2.b"])
                h = self._pw(c, tn + ".ffn2", self._pw(c, tn + ".ffn1", n, tc.SILU), tc.NONE, h)
            n = tc.groupnorm1(h, c[name + ".ln.g"], c[name + ".ln.b"])
            return self._pw(c, name + ".proj", n)

    @torch.no_grad()
    def forward(self, image, as_f32: bool = False):
        """image: fp32 (B,3,H,W), H and W multiples of 64 (so that every 2x2-patch grid is whole).  Returns the five feature maps
        [x2, x4, x8, x16, x32] as bf16 blocked (B,C/8,h,w,8) (what Decoder2D takes directly) or, as_f32=True, fp32 NCHW."""
        if image.dim() != 4 or image.shape[1] != 3 or image.shape[2] % 64 or image.shape[3] % 64:
            raise ValueError("MobileViTv2Backbone: (B,3,H,W) image with H, W multiples of 64 expected")
        c = self._packed()
        enc = self.encoder.layer
        with ops.label("stem"):
            x = tc.stem_conv(image.contiguous().float(), c["stem.w"], c["stem.s"], c["stem.t"], 64)
        feats = []
        x = self._ir(c, "L0.0", enc[0].layer[0], x)
        feats.append(x)
        x = self._ir(c, "L1.1", enc[1].layer[1], self._ir(c, "L1.0", enc[1].layer[0], x))
        feats.append(x)
        for li in (2, 3, 4):
            x = self._vit(c, f"L{li}", enc[li], x)
            feats.append(x)
        return [tc.from_blocked2d(f) for f in feats] if as_f32 else feats


class SemStereoB200(nn.Module):
    """The whole model on the device: backbone -> 2-D decoder -> disparity path = `SemStereo.forward` (models/SemStereo.py:246-346)
    in eval mode, from the two normalised images.  state_dict: `feature.*` (HuggingFace MobileViTv2 names) + the reference's own
    keys for everything else."""

    def __init__(self, maxdisp: int, att_weights_only: bool = False, signed: bool = True, num_classes: int = 6):
        super().__init__()
        from .decoder import StereoHead
        self.feature = MobileViTv2Backbone()
        self.head = StereoHead(maxdisp, att_weights_only, signed, num_classes)

    def load_state_dict(self, state_dict, strict=False, **kw):
        self.feature.load_state_dict({k: v for k, v in state_dict.items() if k.split(".")[0] in ("feature", "module") or k.startswith(("conv_stem", "encoder"))},
                                     strict=strict, **kw)
        return self.head.load_state_dict(state_dict, strict=strict, **kw)

    def state_dict(self, *a, **k):
        sd = {"feature." + n: v for n, v in self.feature.state_dict(*a, **k).items()}
        sd.update(self.head.state_dict(*a, **k))
        return sd

    @torch.no_grad()
    def forward(self, left, right, keep: bool = False):
        B = left.shape[0]
        f = self.feature(torch.cat((left, right), 0))            # both images in one pass of the backbone (batch 2B)
        out = self.head([t[:B] for t in f], [t[B:] for t in f], keep=keep)
        return out

    def as_model_outputs(self, out):
        return self.head.as_model_outputs(out)
