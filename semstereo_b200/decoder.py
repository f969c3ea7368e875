"""Decoder2D — the 2-D decoder the reference wraps around the disparity path, on the tcgen05
tensor cores: FeatUp (models/SemStereo.py:59-86), the two segmentheads (models/submodule.py:31-52), chal_0..4 and the spx
chain (models/SemStereo.py:196-216, 246-271).  It turns the backbone pyramids [x2, x4, x8, x16, x32] of the left and right image
into exactly what `DisparityHotPath` consumes.  `StereoHead` chains both: everything of `SemStereo.forward` after `self.feature`.

Parameter containers mirror the reference's sub-modules, so `state_dict()` keys / shapes equal the reference's and a reference
checkpoint loads unchanged; they are storage only.  Activations are bf16 blocked (B, C/8, H, W, 8) between layers, fp32
accumulation, eval-mode BatchNorm and conv biases folded into the epilogue.  `torch.cat((x, rem), 1)` of Conv2x never
materialises (the GEMM's K loop walks both tensors).  There is no fp32 / torch fallback: CUDA only, bf16 only.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from . import ops_tc as tc
from .hotpath import DisparityHotPath, _ConvBN, bn_affine
from .params import BACKBONE_CHANS, CHANS, CHANS2


class _Conv2xParams(nn.Module):
    """Keys of Conv2x(deconv=True, concat=True) (models/submodule.py:119-146)."""

    def __init__(self, cin, cout):
        super().__init__()
        self.conv1 = _ConvBN(nn.ConvTranspose2d(cin, cout, 4, 2, 1, bias=False), nn.BatchNorm2d(cout))
        self.conv2 = _ConvBN(nn.Conv2d(2 * cout, 2 * cout, 3, 1, 1, bias=False), nn.BatchNorm2d(2 * cout))


class _FeatUpParams(nn.Module):
    """Keys of FeatUp (models/SemStereo.py:59-68)."""

    def __init__(self):
        super().__init__()
        c = BACKBONE_CHANS
        self.deconv32_16 = _Conv2xParams(c[4], c[3])
        self.deconv16_8 = _Conv2xParams(c[3] * 2, c[2])
        self.deconv8_4 = _Conv2xParams(c[2] * 2, c[1])
        self.deconv4_2 = _Conv2xParams(c[1] * 2, c[0])


class _SegHeadParams(nn.Module):
    """Keys of segmenthead (models/submodule.py:33-38)."""

    def __init__(self, inplanes, interplanes, outplanes):
        super().__init__()
        self.conv1 = _ConvBN(nn.Conv2d(inplanes, interplanes, 3, padding=1, bias=False), nn.BatchNorm2d(interplanes))
        self.conv2 = nn.Conv2d(interplanes, outplanes, 1)


class Decoder2D(nn.Module):
    def __init__(self, num_classes: int = 6, precision: str = "bf16"):
        """precision "bf16": every conv with bf16 operands.  "split": the convs that feed the disparity path's sample selection --
        FeatUp and chal_1 / chal_2 (f4_*, f8_*) -- with fp32-accurate bf16x3 products (each conv is one GEMM over the
        [hi | lo | hi] K-concat form, ops_tc.pack_weight2d_split; 3x the MMAs, fp32 activations between layers), the segmentation
        heads and the spx chain (they only feed SSR_upsample) stay bf16.  With DisparityHotPath(precision="split") behind it the
        top-k sample sets equal the fp32 oracle's end to end from fp32 backbone features."""
        super().__init__()
        if precision not in ("bf16", "split"):
            raise ValueError("Decoder2D precision must be 'bf16' or 'split'")
        self.precision = precision
        self.num_classes = num_classes
        self.feature_up = _FeatUpParams()
        self.head_l = _SegHeadParams(CHANS[0], CHANS[0] // 4, num_classes)
        self.head_r = _SegHeadParams(CHANS[0], CHANS[0] // 4, num_classes)
        self.spx2 = nn.Sequential(nn.ConvTranspose2d(CHANS2[0] * 2, 6, 4, 2, 1))
        self.spx4_2 = _Conv2xParams(CHANS2[1] * 2, CHANS2[0])
        self.spx8_4 = _Conv2xParams(CHANS2[2] * 2, CHANS2[1])
        self.spx16_8 = _Conv2xParams(CHANS2[3] * 2, CHANS2[2])
        self.spx32_16 = _Conv2xParams(CHANS2[4], CHANS2[3])
        for i in range(5):
            setattr(self, f"chal_{i}", nn.Sequential(nn.Conv2d(CHANS[i], CHANS2[i], 1), nn.BatchNorm2d(CHANS2[i])))
        self._cache = None
        self.eval()
        for p in self.parameters():
            p.requires_grad_(False)

    # ------------------------------------------------------------------------------------------
    def load_state_dict(self, state_dict, strict=False, **kw):
        """Accepts a full reference checkpoint (foreign keys ignored, leading 'module.' stripped)."""
        own = self.state_dict()
        sd = {}
        for k, v in state_dict.items():
            k = k[7:] if k.startswith("module.") else k
            if k in own:
                sd[k] = v
        missing = [k for k in own if k not in sd and not k.endswith("num_batches_tracked")]
        if strict and missing:
            raise KeyError(f"missing decoder keys: {missing[:5]} ...")
        out = super().load_state_dict(sd, strict=False, **kw)
        self._cache = None
        return out

    def refresh(self):
        self._cache = None

    def _apply(self, fn, *a, **k):
        self._cache = None
        return super()._apply(fn, *a, **k)

    def _packed(self):
        if self._cache is not None:
            return self._cache
        if self.training:
            raise NotImplementedError("Decoder2D is inference-only (eval-mode BatchNorm is folded)")
        c = {}

        def convbn(name, m, mode):
            c[name + ".w"] = tc.pack_weight2d(m.conv.weight, mode)
            c[name + ".s"], c[name + ".t"] = bn_affine(m.bn)

        def conv2x(name, m):
            convbn(name + ".conv1", m.conv1, tc.DECONV4)
            convbn(name + ".conv2", m.conv2, tc.CONV3)
            if self.precision == "split" and name.startswith("feature_up."):
                half = m.conv2.conv.weight.shape[1] // 2
                c[name + ".conv1.ws"] = tc.pack_weight2d_split(m.conv1.conv.weight, tc.DECONV4)
                c[name + ".conv2.ws"] = tc.pack_weight2d_split(m.conv2.conv.weight, tc.CONV3, (half, half))

        for n in ("deconv32_16", "deconv16_8", "deconv8_4", "deconv4_2"):
            conv2x("feature_up." + n, getattr(self.feature_up, n))
        for n in ("spx4_2", "spx8_4", "spx16_8", "spx32_16"):
            conv2x(n, getattr(self, n))
        for h in ("head_l", "head_r"):
            m = getattr(self, h)
            convbn(h + ".conv1", m.conv1, tc.CONV3)
            c[h + ".w2"] = m.conv2.weight.detach().float().reshape(self.num_classes, -1).contiguous()
            c[h + ".b2"] = m.conv2.bias.detach().float().contiguous()
        c["spx2.w"] = tc.pack_weight2d(self.spx2[0].weight, tc.DECONV4)
        c["spx2.b"] = self.spx2[0].bias.detach().float().contiguous()
        for i in range(5):
            m = getattr(self, f"chal_{i}")
            s, t = bn_affine(m[1])
            c[f"chal_{i}.w"] = tc.pack_weight2d(m[0].weight, tc.CONV1)
            if self.precision == "split" and i in (1, 2):
                c[f"chal_{i}.ws"] = tc.pack_weight2d_split(m[0].weight, tc.CONV1)
            c[f"chal_{i}.s"] = s
            c[f"chal_{i}.t"] = (t + s * m[0].bias.detach().float()).contiguous()       # BN(conv + b) = s*conv + (s*b + t)
        self._cache = c
        return c

    # ------------------------------------------------------------------------------------------
    def _conv2x(self, c, name, x, rem):
        """Conv2x.forward (models/submodule.py:148-161) for sizes that are multiples of 32 (no bilinear resize branch)."""
        cout = rem.shape[1] * 8
        with ops.label(name + ".conv1"):
            y = tc.conv2d_tc(tc.DECONV4, x, c[name + ".conv1.w"], cout, c[name + ".conv1.s"], c[name + ".conv1.t"], relu=True)
        if y.shape != rem.shape:
            raise NotImplementedError("Decoder2D: feature sizes must halve exactly between levels (H, W multiples of 32)")
        with ops.label(name + ".conv2"):
            return tc.conv2d_tc(tc.CONV3, y, c[name + ".conv2.w"], 2 * cout, c[name + ".conv2.s"], c[name + ".conv2.t"], relu=True, x1=rem)

    def _conv2x_split(self, c, name, x, rem):
        """Conv2x with fp32-accurate products: x, rem fp32 NCHW -> fp32 NCHW."""
        cout = rem.shape[1]
        with ops.label(name + ".conv1"):
            y = tc.conv2d_tc(tc.DECONV4, tc.to_blocked_tri(x), c[name + ".conv1.ws"], cout, c[name + ".conv1.s"], c[name + ".conv1.t"], relu=True,
                             out_f32=True)
        if y.shape != rem.shape:
            raise NotImplementedError("Decoder2D: feature sizes must halve exactly between levels (H, W multiples of 32)")
        with ops.label(name + ".conv2"):
            return tc.conv2d_tc(tc.CONV3, tc.to_blocked_tri(y), c[name + ".conv2.ws"], 2 * cout, c[name + ".conv2.s"], c[name + ".conv2.t"],
                                relu=True, out_f32=True, x1=tc.to_blocked_tri(rem))

    def _feat_up_split(self, c, f):
        x2, x4, x8, x16, x32 = f
        x16 = self._conv2x_split(c, "feature_up.deconv32_16", x32, x16)
        x8 = self._conv2x_split(c, "feature_up.deconv16_8", x16, x8)
        x4 = self._conv2x_split(c, "feature_up.deconv8_4", x8, x4)
        x2 = self._conv2x_split(c, "feature_up.deconv4_2", x4, x2)
        return [x2, x4, x8, x16, x32]

    def _feat_up(self, c, f):
        x2, x4, x8, x16, x32 = f
        x16 = self._conv2x(c, "feature_up.deconv32_16", x32, x16)
        x8 = self._conv2x(c, "feature_up.deconv16_8", x16, x8)
        x4 = self._conv2x(c, "feature_up.deconv8_4", x8, x4)
        x2 = self._conv2x(c, "feature_up.deconv4_2", x4, x2)
        return [x2, x4, x8, x16, x32]

    def _head(self, c, name, x):
        with ops.label(name):
            y = tc.conv2d_tc(tc.CONV3, x, c[name + ".conv1.w"], CHANS[0] // 4, c[name + ".conv1.s"], c[name + ".conv1.t"], relu=True)
            return tc.bilinear_up2(tc.pointwise_blocked_small(y, c[name + ".w2"], c[name + ".b2"]))

    def _chal(self, c, i, x, out_f32=False):
        with ops.label(f"chal_{i}"):
            return tc.conv2d_tc(tc.CONV1, x, c[f"chal_{i}.w"], CHANS2[i], c[f"chal_{i}.s"], c[f"chal_{i}.t"], out_f32=out_f32)

    @torch.no_grad()
    def forward(self, feat_l, feat_r, right_label: bool = False):
        """feat_l / feat_r: the five maps of `Feature` (SemStereo.py:47-56) for the left / right image, fp32 NCHW or (from
        backbone.MobileViTv2Backbone) bf16 blocked.
        Returns the inputs of DisparityHotPath (fp32 NCHW) plus `f4_l_blocked` (bf16) so the path does not convert f4_l again."""
        c = self._packed()
        blocked = feat_l[0].dtype == torch.bfloat16      # straight from MobileViTv2Backbone: already (B,C/8,H,W,8) bf16
        for f in (feat_l, feat_r):
            if len(f) != 5 or any(t.shape[1] * (8 if blocked else 1) != ch for t, ch in zip(f, BACKBONE_CHANS)):
                raise ValueError(f"Decoder2D: five backbone maps with {BACKBONE_CHANS} channels expected")
        split = self.precision == "split"
        if split and blocked:
            raise ValueError("Decoder2D(precision='split') takes fp32 backbone features (bf16 inputs have already lost the low bits)")
        if split:
            # fp32-accurate FeatUp + chal_1 / chal_2; bf16 blocked copies of its outputs feed the heads and the spx chain
            sl, sr = self._feat_up_split(c, [t.contiguous().float() for t in feat_l]), self._feat_up_split(c, [t.contiguous().float() for t in feat_r])
            with ops.label("to_blocked"):
                fl = [tc.to_blocked2d(t) for t in sl]
                fr = [None, None, None, None, None]
                if right_label:
                    fr[0] = tc.to_blocked2d(sr[0])
        else:
            if blocked:
                bl, br = [t.contiguous() for t in feat_l], [t.contiguous() for t in feat_r]
            else:
                with ops.label("to_blocked"):
                    bl = [tc.to_blocked2d(t) for t in feat_l]
                    br = [tc.to_blocked2d(t) for t in feat_r]
            fl, fr = self._feat_up(c, bl), self._feat_up(c, br)
        out = {"pred_label": self._head(c, "head_l", fl[0])}
        if right_label:
            out["pred_label_r"] = self._head(c, "head_r", fr[0])
        cl = [self._chal(c, i, fl[i]) for i in range(5)]
        if split:
            for i, k in ((1, "f4"), (2, "f8")):
                with ops.label(f"chal_{i}"):
                    for side, src in (("_l", sl), ("_r", sr)):
                        out[k + side] = tc.conv2d_tc(tc.CONV1, tc.to_blocked_tri(src[i]), c[f"chal_{i}.ws"], CHANS2[i], c[f"chal_{i}.s"], c[f"chal_{i}.t"],
                                                     out_f32=True)
            out["f4_l_blocked"] = None      # the path converts its own copy from the fp32-accurate f4_l
        else:
            with ops.label("from_blocked"):
                out["f4_l"], out["f8_l"] = tc.from_blocked2d(cl[1]), tc.from_blocked2d(cl[2])
            out["f4_r"], out["f8_r"] = self._chal(c, 1, fr[1], out_f32=True), self._chal(c, 2, fr[2], out_f32=True)
            out["f4_l_blocked"] = cl[1]
        x = self._conv2x(c, "spx32_16", cl[4], cl[3])
        x = self._conv2x(c, "spx16_8", x, cl[2])
        x = self._conv2x(c, "spx8_4", x, cl[1])
        x = self._conv2x(c, "spx4_2", x, cl[0])
        with ops.label("spx2"):
            out["spx_pred"] = tc.conv2d_tc(tc.DECONV4, x, c["spx2.w"], 6, None, c["spx2.b"], out_f32=True)
        return out


class StereoHead(nn.Module):
    """Everything of SemStereo.forward after `self.feature` (models/SemStereo.py:249-346): Decoder2D + DisparityHotPath.
    state_dict keys are the reference's (both sub-modules register their containers at the top level of this module)."""

    def __init__(self, maxdisp: int, att_weights_only: bool = False, signed: bool = True, num_classes: int = 6, precision: str = "bf16"):
        """precision "bf16" (default, fastest): bf16 operands everywhere -- statistical parity only (top-24 sample sets agree with
        the fp32 oracle on ~96 % of the pixels, median |disparity error| ~0.03 px in 1/4-res units; tests/test_gpu_decoder.py).
        "split": fp32-accurate bf16x3 products in FeatUp, chal_1/2 and the attention branch (Decoder2D / DisparityHotPath
        precision="split"): the sample sets equal the oracle's (>= 99.9 %), the aggregation stays bf16; ~2x slower."""
        super().__init__()
        if precision not in ("bf16", "split"):
            raise ValueError("StereoHead precision must be 'bf16' or 'split'")
        self.precision = precision
        self.decoder = Decoder2D(num_classes, precision)
        self.path = DisparityHotPath(maxdisp, att_weights_only, signed, num_classes, precision=precision)

    def load_state_dict(self, state_dict, strict=False, **kw):
        a = self.decoder.load_state_dict(state_dict, strict=strict, **kw)
        self.path.load_state_dict(state_dict, strict=strict, **kw)
        return a

    def state_dict(self, *a, **k):
        sd = self.decoder.state_dict(*a, **k)
        sd.update(self.path.state_dict(*a, **k))
        return sd

    @torch.no_grad()
    def forward(self, feat_l, feat_r, keep: bool = False, right_label: bool = False):
        d = self.decoder(feat_l, feat_r, right_label=right_label)
        out = self.path(d["f8_l"], d["f8_r"], d["f4_l"], d["f4_r"], None, None, d["spx_pred"], d["pred_label"], keep=keep,
                        f4_l_blocked=d["f4_l_blocked"])
        out["pred_label"] = d["pred_label"]
        if right_label:
            out["pred_label_r"] = d["pred_label_r"]
        if keep:
            out.update({k: d[k] for k in ("f8_l", "f8_r", "f4_l", "f4_r", "spx_pred")})
        return out

    def as_model_outputs(self, out):
        """What SemStereo.forward returns in eval mode (SemStereo.py:340-346)."""
        return self.path.as_model_outputs(out, out["pred_label"])

    def as_loss_inputs(self, out):
        """The tuple SemStereo.forward returns in TRAINING mode (SemStereo.py:329-337) — ([pred_up*4, pred*4, pred_att_up*4,
        pred_att*4], pred_label, pred_label_r), or the two-element list in attention_weights_only mode — from an inference pass
        run with right_label=True, so the reference's model_loss_* / LRSC_loss (models/loss.py:19-135, BASELINE config #4) can be
        evaluated on top.  No gradients: the path is inference-only."""
        if "pred_label_r" not in out:
            raise ValueError("as_loss_inputs: run the forward with right_label=True")
        if self.path.att_weights_only:
            disp = [out["pred_att_up"] * 4, out["pred_att"] * 4]
        else:
            disp = [out["pred_up"] * 4, out["pred"].squeeze(1) * 4, out["pred_att_up"] * 4, out["pred_att"] * 4]
        return disp, out["pred_label"], out["pred_label_r"]
