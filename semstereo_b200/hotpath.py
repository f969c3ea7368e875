"""DisparityHotPath — the reference's disparity path (models/SemStereo.py:273-324 and
models/SemStereo_WHU.py:273-324) executed by the sm_100a kernels behind the C-ABI.

The module owns ordinary torch parameter containers (nn.Conv3d, nn.BatchNorm3d, nn.Linear ...) laid out
exactly like the reference's sub-modules, so `state_dict()` keys/shapes equal the reference's
 and a reference checkpoint loads with `load_state_dict(..., strict=False)`.
Those containers are storage only: their `forward` is never called.  `forward` here issues the fused
CUDA kernels; folded/packed weights are cached and rebuilt after `load_state_dict` / parameter edits
(`refresh()`).

Inputs (what the out-of-scope 2-D part of the model produces):
  f8_l,f8_r (B,256,H/8,W/8); f4_l,f4_r (B,128,H/4,W/4); cf_l,cf_r (B,32,H/4,W/4) = concat_feature(f4_*);
  spx_pred (B,6,H,W); pred_label (B,6,H,W).
Outputs are in 1/4-res disparity units exactly as `ssr_upsample` returns them; `SemStereo.forward`
multiplies by 4 (SemStereo.py:329-346) — `as_model_outputs` does the same.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from . import ops_tc as tc

TOPK = 24  # SemStereo.py:301


def _convbn3d(cin, cout, k, stride, pad):
    """Parameter container with the key layout of convbn_3d (submodule_other.py:845-848): '0' conv, '1' bn."""
    return nn.Sequential(nn.Conv3d(cin, cout, k, stride, pad, bias=False), nn.BatchNorm3d(cout))


class _AttentionParams(nn.Module):
    """Keys of attention_block (submodule_other.py:790-803)."""

    def __init__(self, c):
        super().__init__()
        self.qkv_3d = nn.Linear(c, 3 * c, bias=True)
        self.final1x1 = nn.Conv3d(c, c, 1)


class _HourglassParams(nn.Module):
    """Keys of hourglass / hourglass2 (SemStereo.py:106-132, 145-171)."""

    def __init__(self, c, block):
        super().__init__()
        self.block = tuple(block)
        self.conv1 = nn.Sequential(_convbn3d(c, 2 * c, 3, 2, 1), nn.ReLU())
        self.conv2 = nn.Sequential(_convbn3d(2 * c, 2 * c, 3, 1, 1), nn.ReLU())
        self.conv3 = nn.Sequential(_convbn3d(2 * c, 4 * c, 3, 2, 1), nn.ReLU())
        self.conv4 = nn.Sequential(_convbn3d(4 * c, 4 * c, 3, 1, 1), nn.ReLU())
        self.attention_block = _AttentionParams(4 * c)
        self.conv5 = nn.Sequential(nn.ConvTranspose3d(4 * c, 2 * c, 3, 2, 1, 1, bias=False), nn.BatchNorm3d(2 * c))
        self.conv6 = nn.Sequential(nn.ConvTranspose3d(2 * c, c, 3, 2, 1, 1, bias=False), nn.BatchNorm3d(c))
        self.redir1 = _convbn3d(c, c, 1, 1, 0)
        self.redir2 = _convbn3d(2 * c, 2 * c, 1, 1, 0)


class _ConvBN(nn.Module):
    """Keys of BasicConv (submodule.py:89-107): 'conv', 'bn'."""

    def __init__(self, conv, bn):
        super().__init__()
        self.conv, self.bn = conv, bn


class _ChannelAttParams(nn.Module):
    """Keys of channelAtt (SemStereo.py:89-96)."""

    def __init__(self, cv_chan, im_chan):
        super().__init__()
        self.im_att = nn.Sequential(_ConvBN(nn.Conv2d(im_chan, im_chan // 2, 1, bias=False), nn.BatchNorm2d(im_chan // 2)),
                                    nn.Conv2d(im_chan // 2, cv_chan, 1))


class _SSRParams(nn.Module):
    """Keys of SSR_upsample (submodule.py:413-419)."""

    def __init__(self, nc):
        super().__init__()
        self.conv = nn.Sequential(nn.BatchNorm2d(1), nn.Conv2d(1, nc, 3, padding=1), nn.BatchNorm2d(nc))
        self.conv1 = nn.Sequential(nn.Conv2d(nc, nc, 1), nn.BatchNorm2d(nc))
        self.conv2 = nn.Sequential(nn.Conv2d(nc, nc, 1), nn.BatchNorm2d(nc))
        self.conv3 = nn.Conv2d(nc, 1, 1)


def bn_affine(bn):
    """Eval-mode BatchNorm as per-channel (scale, shift) from the running statistics."""
    scale = bn.weight.detach() / torch.sqrt(bn.running_var + bn.eps)
    shift = bn.bias.detach() - bn.running_mean * scale
    return scale.float().contiguous(), shift.float().contiguous()


def pack_ssr(ssr: nn.Module) -> torch.Tensor:
    """Folded SSR_upsample parameters in the order ss_ssr_upsample expects (include/semstereo_b200.h)."""
    a0, b0 = bn_affine(ssr.conv[0])
    s2, t2 = bn_affine(ssr.conv[2])
    s1, t1 = bn_affine(ssr.conv1[1])
    sb, tb = bn_affine(ssr.conv2[1])
    parts = [a0, b0, ssr.conv[1].weight.detach().reshape(-1), ssr.conv[1].bias.detach(), s2, t2,
             ssr.conv1[0].weight.detach().reshape(-1), ssr.conv1[0].bias.detach(), s1, t1,
             ssr.conv2[0].weight.detach().reshape(-1), ssr.conv2[0].bias.detach(), sb, tb,
             ssr.conv3.weight.detach().reshape(-1), ssr.conv3.bias.detach()]
    return torch.cat([p.float().reshape(-1).cpu() for p in parts]).contiguous()


class DisparityHotPath(nn.Module):
    def __init__(self, maxdisp: int, att_weights_only: bool = False, signed: bool = True, num_classes: int = 6,
                 precision: str = "split"):
        """precision:
        "fp32"  = everything on the fp32 pipe (FFMA; the slow parity mode);
        "bf16"  = every conv and the window attention on the tensor cores with bf16 operands and fp32 accumulation;
        "mixed" = the 1/8-resolution attention branch (9 % of the FLOPs, but it alone decides the top-k sample selection,
                  SURVEY 0.7) on the fp32 pipe and the aggregation branch in bf16;
        "split" = the attention branch on the tensor cores with fp32-ACCURATE products (bf16x3: every operand is carried as a
                  hi/lo bf16 pair, x*w = x_hi*w_hi + x_lo*w_hi + x_hi*w_lo, fp32 accumulation; the window attention core and the
                  gate convs stay fp32) and the aggregation branch in bf16.  Sample selection agrees with the fp32 oracle on
                  > 99.9 % of the pixels (differences only at near-ties), at tensor-core speed: the benchmarked default."""
        super().__init__()
        if precision not in ("fp32", "bf16", "mixed", "split"):
            raise ValueError("precision must be 'fp32', 'bf16', 'mixed' or 'split'")
        self.precision = precision
        if maxdisp % 8:
            raise ValueError("maxdisp must be a multiple of 8 (the 1/8-res volume is upsampled exactly x2)")
        nb = (2 if signed else 1) * (maxdisp // 4)
        if nb < TOPK or nb > 64:
            raise NotImplementedError(f"{nb} attention bins unsupported (need {TOPK} <= bins <= 64)")
        self.maxdisp, self.att_weights_only, self.signed, self.num_classes = maxdisp, att_weights_only, signed, num_classes
        self.gamma = nn.Parameter(torch.zeros(1))
        self.beta = nn.Parameter(2 * torch.ones(1))
        self.patch = nn.Conv3d(32, 32, (1, 3, 3), 1, (0, 1, 1), groups=32, bias=False)
        self.corr_feature_att_8 = _ChannelAttParams(32, 256)
        self.concat_feature_att_4 = _ChannelAttParams(32, 128)
        self.hourglass_att = _HourglassParams(32, (4, 4, 4))
        self.classif_att_ = nn.Sequential(_convbn3d(32, 32, 3, 1, 1), nn.ReLU(), nn.Conv3d(32, 1, 3, 1, 1, bias=False))
        self.hourglass = _HourglassParams(32, (6, 4, 4))
        self.classif = nn.Sequential(_convbn3d(32, 32, 3, 1, 1), nn.ReLU(), nn.Conv3d(32, 1, 3, 1, 1, bias=False))
        self.concat_stem = _ConvBN(nn.Conv3d(64, 32, 3, 1, 1, bias=False), nn.BatchNorm3d(32))
        self.ssr_upsample = _SSRParams(num_classes)
        # concat_feature (SemStereo.py:221-223): used when the caller does not supply cf_l / cf_r
        self.concat_feature = nn.Sequential(_ConvBN(nn.Conv2d(128, 64, 3, 1, 1, bias=False), nn.BatchNorm2d(64)),
                                            nn.Conv2d(64, 32, 3, 1, 1, bias=False))
        self._cache = None
        self.eval()
        for p in self.parameters():
            p.requires_grad_(False)

    # ------------------------------------------------------------------------------------------
    def load_state_dict(self, state_dict, strict=False, **kw):
        """Accepts a full reference checkpoint: keys outside the path are ignored, a leading 'module.' (DataParallel,
        main_us3d.py:100) is stripped."""
        own = self.state_dict()
        sd = {}
        for k, v in state_dict.items():
            k = k[7:] if k.startswith("module.") else k
            if k in own:
                sd[k] = v
        missing = [k for k in own if k not in sd and not k.endswith("num_batches_tracked")]
        if strict and missing:
            raise KeyError(f"missing hot-path keys: {missing[:5]} ...")
        out = super().load_state_dict(sd, strict=False, **kw)
        self._cache = None
        return out

    def refresh(self):
        self._cache = None

    _ATT_BRANCH = ("hourglass_att", "classif_att_", "corr_feature_att_8")

    def _is_bf16(self, name: str) -> bool:
        """Does the module `name` belong to a branch that runs on the tensor cores with plain bf16 operands in this mode?"""
        if self.precision in ("mixed", "split"):
            return not name.startswith(self._ATT_BRANCH)
        return self.precision == "bf16"

    def _is_split(self, name: str) -> bool:
        """bf16x3 split route (fp32-accurate tensor-core products): the 3-D stack of the attention branch in "split" mode."""
        return self.precision == "split" and name.startswith(self._ATT_BRANCH)

    def _apply(self, fn, *a, **k):
        self._cache = None
        return super()._apply(fn, *a, **k)

    # ------------------------------------------------------------------------------------------
    def _packed(self):
        if self._cache is not None:
            return self._cache
        if self.training:
            raise NotImplementedError("DisparityHotPath is inference-only (eval-mode BatchNorm is folded)")
        c = {}

        def conv(name, convmod, bn, transposed=False):
            w = convmod.weight.detach().float()
            if self._is_bf16(name):      # tensor-core packing: kind from the layer geometry
                k, st = convmod.kernel_size[0], convmod.stride[0]
                kind = tc.T2 if transposed else (tc.K1 if k == 1 else (tc.S2 if st == 2 else tc.S1))
                if kind == tc.S1 and tc.ntile(tc.S1F, w.shape[1], w.shape[0]) == w.shape[0]:
                    kind = tc.S1F          # narrow layers: depth taps folded into the GEMM N (csrc/conv3d_tc.cu, s1f)
                c[name + ".kind"] = kind
                c[name + ".tc"] = tc.pack_weight(w, kind)
            elif self._is_split(name):
                k, st = convmod.kernel_size[0], convmod.stride[0]
                kind = tc.T2 if transposed else (tc.K1 if k == 1 else (tc.S2 if st == 2 else tc.S1))
                if kind == tc.S1 and tc.ntile(tc.S1F, w.shape[1], w.shape[0]) == w.shape[0]:
                    kind = tc.S1F
                c[name + ".kind"] = kind
                c[name + ".tcs"] = tc.pack_weight_split(w, kind)
            else:
                c[name + ".w"] = ops.pack_conv3d_weight(w, transposed)
            if bn is not None:
                c[name + ".scale"], c[name + ".shift"] = bn_affine(bn)

        for hg in ("hourglass_att", "hourglass"):
            m = getattr(self, hg)
            for n in ("conv1", "conv2", "conv3", "conv4"):
                conv(f"{hg}.{n}", getattr(m, n)[0][0], getattr(m, n)[0][1])
            conv(f"{hg}.conv5", m.conv5[0], m.conv5[1], True)
            conv(f"{hg}.conv6", m.conv6[0], m.conv6[1], True)
            conv(f"{hg}.redir1", m.redir1[0], m.redir1[1])
            conv(f"{hg}.redir2", m.redir2[0], m.redir2[1])
            bf16 = self._is_bf16(hg)
            if bf16 or self._is_split(hg):
                for dc, rd in (("conv5", "redir2"), ("conv6", "redir1")):
                    deconv, dbn = getattr(m, dc)[0], getattr(m, dc)[1]
                    rconv, rbn = getattr(m, rd)[0], getattr(m, rd)[1]
                    ds, dt = bn_affine(dbn)
                    rs, rt = bn_affine(rbn)
                    wf = deconv.weight.detach().float() * ds.view(1, -1, 1, 1, 1)           # BN scales folded into both weights
                    cc = rconv.weight.shape[0]
                    sf = rconv.weight.detach().float().reshape(cc, cc) * rs.reshape(cc, 1)
                    c[f"{hg}.{dc}.fshift"] = (dt + rt).contiguous()
                    if bf16:
                        c[f"{hg}.{dc}.ftc"] = tc.pack_weight(wf, tc.T2)
                        c[f"{hg}.{dc}.skipw"] = tc.pack_skip_weight(sf)
                    else:
                        c[f"{hg}.{dc}.ftcs"] = tc.pack_weight_split(wf, tc.T2)
                        c[f"{hg}.{dc}.skipws"] = tc.pack_skip_weight_split(sf)
            a = m.attention_block
            c[hg + ".wqkv_t"] = a.qkv_3d.weight.detach().float().t().contiguous()
            c[hg + ".bqkv"] = a.qkv_3d.bias.detach().float().contiguous()
            c[hg + ".wo_t"] = a.final1x1.weight.detach().float().reshape(a.final1x1.out_channels, -1).t().contiguous()
            c[hg + ".bo"] = a.final1x1.bias.detach().float().contiguous()
            if self._is_split(hg):          # fp32-accurate projections as K-concat GEMMs (ops_tc.pointwise_split)
                c[hg + ".attn_qkv.tri"] = tc.pack_pointwise_split(a.qkv_3d.weight)
                c[hg + ".attn_out.tri"] = tc.pack_pointwise_split(a.final1x1.weight)
            if bf16:
                c[hg + ".attn_qkv.tc"] = tc.pack_weight(a.qkv_3d.weight.detach().float().reshape(384, 128, 1, 1, 1), tc.K1)
                c[hg + ".attn_qkv.shift"] = c[hg + ".bqkv"]
                c[hg + ".attn_out.tc"] = tc.pack_weight(a.final1x1.weight.detach().float(), tc.K1)
                c[hg + ".attn_out.shift"] = c[hg + ".bo"]
        for cl in ("classif_att_", "classif"):
            m = getattr(self, cl)
            conv(cl + ".0", m[0][0], m[0][1])
            if self._is_bf16(cl):
                c[cl + ".2.tc"] = tc.pack_head_weight(m[2].weight.detach().float())          # taps-as-N head kernel
            elif self._is_split(cl):
                c[cl + ".2.tcs"] = tc.pack_head_weight_split(m[2].weight.detach().float())
            else:
                c[cl + ".2.w"] = m[2].weight.detach().float().contiguous()
        conv("concat_stem", self.concat_stem.conv, self.concat_stem.bn)
        for ca in ("corr_feature_att_8", "concat_feature_att_4"):
            m = getattr(self, ca).im_att
            c[ca + ".w0"] = m[0].conv.weight.detach().float().reshape(m[0].conv.out_channels, -1).contiguous()
            c[ca + ".s0"], c[ca + ".t0"] = bn_affine(m[0].bn)
            c[ca + ".w1"] = m[1].weight.detach().float().reshape(m[1].out_channels, -1).contiguous()
            c[ca + ".b1"] = m[1].bias.detach().float().contiguous()
            w0, w1 = m[0].conv.weight.detach().float(), m[1].weight.detach().float()
            if self._is_bf16(ca):
                c[ca + ".tc0"] = tc.pack_weight2d(w0, tc.CONV1)
                c[ca + ".tc1"] = tc.pack_weight2d(w1, tc.CONV1)
            elif self._is_split(ca):
                c[ca + ".tri0"], c[ca + ".tri1"] = tc.pack_pointwise_split(w0), tc.pack_pointwise_split(w1)
        cf0, cf1 = self.concat_feature[0], self.concat_feature[1]
        c["cf0.scale"], c["cf0.shift"] = bn_affine(cf0.bn)
        if self._is_bf16("concat_feature"):
            c["cf0.tc"] = tc.pack_weight(cf0.conv.weight.detach().float(), tc.C2D)
            c["cf1.tc"] = tc.pack_weight(cf1.weight.detach().float(), tc.C2D)
        else:      # fp32 mode: the 2-D 3x3 conv as the centre depth plane of a 3x3x3 kernel on a depth-1 volume
            for name, w in (("cf0", cf0.conv.weight), ("cf1", cf1.weight)):
                w3 = w.detach().float().new_zeros((w.shape[0], w.shape[1], 3, 3, 3))
                w3[:, :, 1] = w.detach().float()
                c[name + ".w"] = ops.pack_conv3d_weight(w3)
        c["patch.w"] = self.patch.weight.detach().float().reshape(32, 9).contiguous()
        c["ssr"] = pack_ssr(self.ssr_upsample)
        self._cache = c
        return c

    # ------------------------------------------------------------------------------------------
    def _concat_feature(self, c, f4, f4_blocked=None, blocked=False):
        """concat_feature(features[1]) (SemStereo.py:314-315): (B,128,H/4,W/4) -> (B,32,H/4,W/4) fp32, or (bf16 mode, blocked=True)
        the bf16 blocked (B,4,1,H/4,W/4,8) tensor the fused concat_stem kernel stages."""
        with ops.label("concat_feature"):
            x = f4.unsqueeze(2)
            if self._is_bf16("concat_feature"):
                xb = f4_blocked if f4_blocked is not None else tc.to_blocked_bf16(x)
                y = tc.conv3d_tc(tc.C2D, xb, c["cf0.tc"], 64, c["cf0.scale"], c["cf0.shift"], relu=True)
                if blocked:
                    return tc.conv3d_tc(tc.C2D, y, c["cf1.tc"], 32)
                return tc.conv3d_tc(tc.C2D, y, c["cf1.tc"], 32, out_mode=tc.F32).squeeze(2)
            y = ops.conv3d_f32(x, c["cf0.w"], c["cf0.scale"], c["cf0.shift"], k=3, relu=True)
            return ops.conv3d_f32(y, c["cf1.w"], k=3).squeeze(2)

    def _gate_logits(self, c, name, im, im_blocked=None):
        """channelAtt.im_att (SemStereo.py:93-95): 1x1 conv + BN + ReLU -> 1x1 conv + bias.  bf16 mode: both 1x1 convs run on
        the tensor cores (csrc/conv2d_tc.cu, mode CONV1) from a blocked bf16 copy of the image features."""
        if self._is_bf16(name):
            with ops.label(name):
                B, C, H, W = im.shape
                xb = im_blocked.view(B, C // 8, H, W, 8) if im_blocked is not None else tc.to_blocked2d(im)
                y = tc.conv2d_tc(tc.CONV1, xb, c[name + ".tc0"], C // 2, c[name + ".s0"], c[name + ".t0"], relu=True)
                return tc.conv2d_tc(tc.CONV1, y, c[name + ".tc1"], 32, None, c[name + ".b1"], out_f32=True)
        if self._is_split(name):       # fp32-accurate on the tensor cores: each 1x1 conv is one GEMM over the [hi | lo | hi] K-concat form
            with ops.label(name):
                y = tc.pointwise_split(tc.to_blocked_tri(im), c[name + ".tri0"], im.shape[1] // 2, c[name + ".s0"], c[name + ".t0"], relu=True)
                return tc.pointwise_split(tc.to_blocked_tri(y), c[name + ".tri1"], 32, None, c[name + ".b1"])
        y = ops.pointwise_conv2d(im, c[name + ".w0"], c[name + ".s0"], c[name + ".t0"], relu=True)
        return ops.pointwise_conv2d(y, c[name + ".w1"], None, c[name + ".b1"], relu=False)

    def _conv(self, c, name, x, k=3, stride=1, relu=True, transposed=False, residual=None, gate=None):
        with ops.label(name):
            return self._conv_impl(c, name, x, k, stride, relu, transposed, residual, gate)

    def _conv_impl(self, c, name, x, k, stride, relu, transposed, residual, gate):
        return ops.conv3d_f32(x, c[name + ".w"], c.get(name + ".scale"), c.get(name + ".shift"), residual, gate,
                              k=k, stride=stride, transposed=transposed, relu=relu)

    # ---- bf16 tensor-core flavour: activations stay in the blocked / phase-split bf16 layouts between the layers ----
    def _tc(self, c, name, kind, x, cout, relu=True, gate=None, residual=None, out_mode=tc.BLOCKED):
        with ops.label(name):
            return tc.conv3d_tc(c.get(name + ".kind", kind), x, c[name + ".tc"], cout, c.get(name + ".scale"), c.get(name + ".shift"), gate, residual,
                                relu=relu, out_mode=out_mode)

    def _hourglass_tc(self, c, hg, x_s2d):
        """hourglass.forward (SemStereo.py:134-143) on tensor cores.  x_s2d: phase-split bf16 (B,8,4,D/2,H/2,W/2,8); returns
        blocked bf16 (B,4,D,H,W,8).  The redir 1x1 convs run position-wise on the phase-split tensors and come back as the
        residual of the transposed layers, which add it before their ReLU."""
        block = getattr(self, hg).block
        c1 = self._tc(c, hg + ".conv1", tc.S2, x_s2d, 64)
        c2s = self._tc(c, hg + ".conv2", tc.S1, c1, 64, out_mode=tc.S2D)         # written phase-split for conv3 / redir2
        c3 = self._tc(c, hg + ".conv3", tc.S2, c2s, 128)
        c4 = self._tc(c, hg + ".conv4", tc.S1, c3, 128)
        # attention_block (submodule_other.py:805-837): qkv Linear and final 1x1x1 conv as tensor-core 1x1 layers (bias = shift),
        # the per-(window, head) softmax core in between; a head is one channel chunk of the blocked layout
        if ops.window_needs_mask(c4.shape, block):       # H and W both padded: the reference's masked branch, fp32 (ops.window_attention3d)
            with ops.label(hg + ".attention_masked"):
                c4b = tc.to_blocked_bf16(ops.window_attention3d(tc.from_blocked_bf16(c4), c[hg + ".wqkv_t"], c[hg + ".bqkv"], c[hg + ".wo_t"],
                                                                c[hg + ".bo"], block, 16))
        else:
            c4, H0, W0 = ops.window_pad(c4, block)       # H / W not multiples of the window: zero-pad (one axis), crop after the block
            qkv = self._tc(c, hg + ".attn_qkv", tc.K1, c4, 384, relu=False)
            with ops.label(hg + ".attn_core"):
                att = tc.window_attention_core(qkv, block, 16)
            c4b = ops.window_crop(self._tc(c, hg + ".attn_out", tc.K1, att, 128, relu=False), H0, W0)
        # conv5/conv6 (transposed) with the redir2/redir1 1x1 skip convs fused in as one more GEMM tap on the TMA-staged skip
        # tile (BN scales folded into both weights, shifts summed): relu(bn(deconv(x)) + bn(redir(skip)))  (SemStereo.py:141-142)
        with ops.label(hg + ".conv5"):
            c5 = tc.conv3d_tc(tc.T2, c4b, c[hg + ".conv5.ftc"], 64, None, c[hg + ".conv5.fshift"], residual_s2d=c2s,
                              skip_weight=c[hg + ".conv5.skipw"], relu=True)
        with ops.label(hg + ".conv6"):
            return tc.conv3d_tc(tc.T2, c5, c[hg + ".conv6.ftc"], 32, None, c[hg + ".conv6.fshift"], residual_s2d=x_s2d,
                                skip_weight=c[hg + ".conv6.skipw"], relu=True)

    # ---- bf16x3 split flavour (attention branch, precision="split"): same layer graph, every activation a hi/lo pair stacked
    # on the batch axis, two launches per layer (ops_tc.conv3d_tc_split); the window attention block runs in fp32 ----
    def _tcs(self, c, name, x, cout, relu=True, out_mode=tc.BLOCKED):
        with ops.label(name):
            return tc.conv3d_tc_split(c[name + ".kind"], x, c[name + ".tcs"], cout, c.get(name + ".scale"), c.get(name + ".shift"),
                                      relu=relu, out_mode=out_mode)

    def _hourglass_split(self, c, hg, x_s2d):
        block = getattr(self, hg).block
        c1 = self._tcs(c, hg + ".conv1", x_s2d, 64)
        c2s = self._tcs(c, hg + ".conv2", c1, 64, out_mode=tc.S2D)
        c3 = self._tcs(c, hg + ".conv3", c2s, 128)
        c4 = self._tcs(c, hg + ".conv4", c3, 128, out_mode=tc.F32)                 # fp32 NCDHW for the fp32 attention block
        # attention_block (submodule_other.py:805-837): qkv Linear and final 1x1x1 conv as fp32-accurate K-concat GEMMs (a volume is
        # a 2-D image of D*H rows for a 1x1 conv), the fp32 softmax core in between writes the K-concat form directly
        with ops.label(hg + ".attention"):
            if ops.window_needs_mask(c4.shape, block):   # H and W both padded: the reference's masked branch (ops.window_attention3d)
                c4b = tc.to_blocked_bf16(ops.window_attention3d(c4, c[hg + ".wqkv_t"], c[hg + ".bqkv"], c[hg + ".wo_t"], c[hg + ".bo"], block, 16),
                                         split=True)
            else:
                c4, H0, W0 = ops.window_pad(c4, block)   # H / W not multiples of the window: zero-pad (one axis), crop after the block
                B, C, D, H, W = c4.shape
                qkv = tc.pointwise_split(tc.to_blocked_tri(c4.view(B, C, D * H, W)), c[hg + ".attn_qkv.tri"], 3 * C, None, c[hg + ".bqkv"])
                att = tc.window_attention_core_f32(qkv.view(B, 3 * C, D, H, W), block, 16)
                c4 = tc.pointwise_split(att.view(B, 3 * C // 8, D * H, W, 8), c[hg + ".attn_out.tri"], C, None, c[hg + ".bo"])
                c4b = tc.to_blocked_bf16(ops.window_crop(c4.view(B, C, D, H, W), H0, W0), split=True)
        with ops.label(hg + ".conv5"):
            c5 = tc.conv3d_tc_split(tc.T2, c4b, c[hg + ".conv5.ftcs"], 64, None, c[hg + ".conv5.fshift"], residual_s2d=c2s,
                                    skip_split=c[hg + ".conv5.skipws"], relu=True)
        with ops.label(hg + ".conv6"):
            return tc.conv3d_tc_split(tc.T2, c5, c[hg + ".conv6.ftcs"], 32, None, c[hg + ".conv6.fshift"], residual_s2d=x_s2d,
                                      skip_split=c[hg + ".conv6.skipws"], relu=True)

    def _classifier_split(self, c, cl, xs):
        y = self._tcs(c, cl + ".0", xs, 32)
        with ops.label(cl + ".2"):
            return tc.conv3d_tc_head(y, c[cl + ".2.tcs"], in_split=True)

    def _classifier_tc(self, c, cl, xb):
        y = self._tc(c, cl + ".0", tc.S1, xb, 32)
        with ops.label(cl + ".2"):
            return tc.conv3d_tc_head(y, c[cl + ".2.tc"])

    def _hourglass(self, c, hg, x):
        """hourglass.forward (SemStereo.py:134-143): residual adds and ReLUs ride in the deconv epilogues."""
        block = getattr(self, hg).block
        c1 = self._conv(c, hg + ".conv1", x, stride=2)
        c2 = self._conv(c, hg + ".conv2", c1)
        c3 = self._conv(c, hg + ".conv3", c2, stride=2)
        c4 = self._conv(c, hg + ".conv4", c3)
        c4 = ops.window_attention3d(c4, c[hg + ".wqkv_t"], c[hg + ".bqkv"], c[hg + ".wo_t"], c[hg + ".bo"], block, 16)
        r2 = self._conv(c, hg + ".redir2", c2, k=1, relu=False)
        c5 = self._conv(c, hg + ".conv5", c4, stride=2, transposed=True, residual=r2, relu=True)
        r1 = self._conv(c, hg + ".redir1", x, k=1, relu=False)
        return self._conv(c, hg + ".conv6", c5, stride=2, transposed=True, residual=r1, relu=True)

    def _classifier(self, c, cl, x):
        y = self._conv(c, cl + ".0", x)
        with ops.label(cl + ".2"):
            return ops.conv3d_cout1_f32(y, c[cl + ".2.w"])

    @torch.no_grad()
    def forward(self, f8_l, f8_r, f4_l, f4_r, cf_l, cf_r, spx_pred, pred_label, keep: bool = False, f4_l_blocked=None):
        """cf_l / cf_r may be None: concat_feature(f4_*) is then computed here (tensor cores in bf16 mode)."""
        c = self._packed()
        m8, m4 = self.maxdisp // 8, self.maxdisp // 4
        out = {}
        # --- attention branch @1/8 (SemStereo.py:273-278) ---
        corr = ops.gwc_volume(f8_l, f8_r, m8, 32, signed=self.signed, norm=True)
        gate8 = self._gate_logits(c, "corr_feature_att_8", f8_l)
        if self._is_bf16("hourglass_att"):
            vol = tc.patch_gate_blocked(corr, c["patch.w"], gate8)
            cost_att = self._classifier_tc(c, "classif_att_", self._hourglass_tc(c, "hourglass_att", vol))
        elif self._is_split("hourglass_att"):
            vol = tc.patch_gate_blocked(corr, c["patch.w"], gate8, split=True)
            cost_att = self._classifier_split(c, "classif_att_", self._hourglass_split(c, "hourglass_att", vol))
        else:
            vol = ops.patch_gate(corr, c["patch.w"], gate8)
            vol = self._hourglass(c, "hourglass_att", vol)
            cost_att = self._classifier(c, "classif_att_", vol)
        # --- statistics, propagation, top-k @1/4 (SemStereo.py:279-310) ---
        dmin = float(-m4) if self.signed else 0.0
        att_up, mu, gate = ops.att_stats(cost_att, self.beta.data, self.gamma.data, dmin)
        strength = ops.sample_strength(f4_l, f4_r, mu, gate)
        ind_k, att_topk, disp_topk, pred_att, prob = ops.topk_select(att_up, strength, TOPK, -dmin, want_indices=keep, want_prob=keep)
        out.update(cost_att=cost_att, pred_att=pred_att, disp_topk=disp_topk, att_topk=att_topk)
        if keep:
            out.update(corr_volume=corr, att_weights=att_up, pred_att0=mu, var_gate=gate, strength=strength, ind_k=ind_k, prob=prob)
        if self.att_weights_only:
            out["pred_att_up"] = ops.ssr_upsample(pred_att.unsqueeze(1), spx_pred, pred_label, c["ssr"])
            return out
        # --- sparse concat volume + aggregation (SemStereo.py:314-324) ---
        main_bf16 = self._is_bf16("hourglass")
        f4l_b = None                            # bf16 blocked copy of f4_l: feeds concat_feature and the gate convs
        # bf16 mode without kept intermediates: the sparse concat volume is generated inside the concat_stem kernel (never in HBM)
        nbins = (2 if self.signed else 1) * m4
        fused = main_bf16 and not keep and nbins == 32
        if fused and cf_l is None and cf_r is None and f4_l_blocked is None:
            # concat_feature of BOTH images as one batch of 2B through its two tensor-core layers (2 launches instead of 4)
            Bn = f4_l.shape[0]
            both = torch.empty((2 * Bn, 16, 1, *f4_l.shape[2:], 8), device=f4_l.device, dtype=torch.bfloat16)
            with ops.label("concat_feature"):
                tc.to_blocked_bf16(f4_l.unsqueeze(2), out=both[:Bn])
                tc.to_blocked_bf16(f4_r.unsqueeze(2), out=both[Bn:])
                y = tc.conv3d_tc(tc.C2D, both, c["cf0.tc"], 64, c["cf0.scale"], c["cf0.shift"], relu=True)
                cf = tc.conv3d_tc(tc.C2D, y, c["cf1.tc"], 32)
            f4l_b, cf_l, cf_r = both[:Bn], cf[:Bn], cf[Bn:]
        elif main_bf16:
            f4l_b = (f4_l_blocked.view(f4_l.shape[0], 16, 1, *f4_l.shape[2:], 8) if f4_l_blocked is not None
                     else tc.to_blocked_bf16(f4_l.unsqueeze(2)))
        if cf_l is None:
            cf_l = self._concat_feature(c, f4_l, f4l_b, blocked=fused)
        elif fused and cf_l.dtype != torch.bfloat16:
            cf_l = tc.to_blocked2d(cf_l)
        if cf_r is None:
            cf_r = self._concat_feature(c, f4_r, blocked=fused)
        elif fused and cf_r.dtype != torch.bfloat16:
            cf_r = tc.to_blocked2d(cf_r)
        gate4 = self._gate_logits(c, "concat_feature_att_4", f4_l, f4l_b)
        if fused:
            volume = None
            gb = tc.gate_sigmoid_blocked(gate4)
            with ops.label("concat_stem"):
                v = tc.concat_stem_fused(cf_l, cf_r, disp_topk, att_topk, c["concat_stem.tc"], int(dmin), c["concat_stem.scale"],
                                         c["concat_stem.shift"], gb, relu=True, out_mode=tc.S2D)
            cost = self._classifier_tc(c, "classif", self._hourglass_tc(c, "hourglass", v))
        elif main_bf16:
            volume = tc.sparse_concat_volume_blocked(cf_l, cf_r, disp_topk, att_topk)      # (B,8,24,H/4,W/4,8) bf16
            v = self._tc(c, "concat_stem", tc.S1, volume, 32, gate=tc.gate_sigmoid_blocked(gate4), out_mode=tc.S2D)
            cost = self._classifier_tc(c, "classif", self._hourglass_tc(c, "hourglass", v))
        else:
            volume = ops.sparse_concat_volume(cf_l, cf_r, disp_topk, att_topk)
            v = self._conv(c, "concat_stem", volume, gate=gate4)
            v = self._hourglass(c, "hourglass", v)
            cost = self._classifier(c, "classif", v)
        pred = ops.regression_topk(cost.squeeze(1), disp_topk, 2)
        out.update(cost=cost, pred=pred)
        if keep:
            out["volume"] = volume
        # both SSR_upsample calls of the model (:312 on pred_att, :324 on pred) share spx / label: one pass
        out["pred_att_up"], out["pred_up"] = ops.ssr_upsample2(pred_att.unsqueeze(1), pred, spx_pred, pred_label, c["ssr"])
        return out

    def as_model_outputs(self, out, pred_label):
        """What SemStereo.forward returns in eval mode with seg_if (SemStereo.py:338-346)."""
        key = "pred_att_up" if self.att_weights_only else "pred_up"
        return [out[key] * 4], pred_label
