"""Autograd for the 3-D convolution stack in TRAINING mode (BASELINE config #5): what
`convbn_3d` / `BasicConv(is_3d)` (models/submodule_other.py:845-848, models/submodule.py:89-116) do when the reference trains
(main_us3d.py:186-222) -- Conv3d and BatchNorm3d with batch statistics, forward AND backward on the CUDA kernels (fp32):

  conv3d        forward  the bf16x3 tcgen05 kernels (ops_tc.conv3d_tc_split: fp32-accurate products, fp32 accumulation) for the k = 3
                         layers that have such a configuration (32->32, 64->64, 128->128, s2 32->64 / 64->128), else ss_conv3d_f32
                         (csrc/conv3d_f32.cu, the inference kernel without the folded BN)
                dX       the same kernels on dY with a re-packed weight: flipped taps + swapped channels (k3 s1), the
                         ConvTranspose3d(k3,s2,p1,op1) layer with the same weight (k3 s2: t2 64->32 / 128->64), W^T (k1)
                dW       ss_conv3d_wgrad_f32                 (csrc/train.cu)
  batch_norm    forward / backward ss_bn_train_forward / ss_bn_train_backward, running statistics updated like nn.BatchNorm3d
                (momentum, unbiased variance)

  attention     qkv Linear / final 1x1x1 conv = k = 1 convs above; softmax core ss_window_attention_core_f32_out / _backward
  SSR_upsample  bilinear x4 (ss_bilinear_up4 / _backward), 2-D convs through the 3-D kernels, batch_norm; pointwise glue in torch

The stateless operators of the surface (volume builders, regression, warps, propagation) already carry their backward kernels
(torch_ops.py / csrc/backward.cu).  With these the reference model TRAINS through the level-1 drop-in: every module /
function it takes from models.submodule(_other) is differentiable on the CUDA kernels; what the model does inline
(nn.ConvTranspose3d, the `patch` / classifier nn.Conv3d, interpolate, softmax, sort, gather, the 2-D decoder) is torch, as in the
reference.  Forward and dX of the k = 3 layers run on the bf16x3 tensor-core kernels (fp32-accurate), dW and the rest are fp32 FFMA
kernels; tools/train_step.py measures the step (config #5)."""
from __future__ import annotations

import ctypes

import torch

from . import ops
from . import ops_tc as tc
from .ops import _call, _ptr, _require_cuda


def _geometry(w, stride):
    k = w.shape[2]
    if tuple(w.shape[2:]) != (k, k, k) or k not in (1, 3) or stride not in (1, 2):
        raise NotImplementedError("conv3d (training): only k in {1,3} cubic kernels with pad k//2 and stride in {1,2}")
    return k


_ROUTE = {"tensor_cores": True}


def set_tensor_core_route(on: bool):
    """Forward and dX of the k = 3 layers that have a bf16x3 split configuration run on the tcgen05 kernels (fp32-accurate products,
    fp32 accumulation: ops_tc.conv3d_tc_split) by default; False forces the fp32 FFMA kernels everywhere."""
    _ROUTE["tensor_cores"] = bool(on)


def _tc_kind(cin, cout, k, stride, transposed=False):
    """The tensor-core kind of a layer that has a bf16x3 split configuration (single launch, or the two-launch route of the
    128-channel layers), else None."""
    if k != 3 or not _ROUTE["tensor_cores"]:
        return None
    if transposed:
        return tc.T2 if (cin, cout) in ((64, 32), (128, 64)) else None
    if stride == 2:
        return tc.S2 if (cin, cout) in ((32, 64), (64, 128)) else None
    if (cin, cout) in ((32, 32), (64, 64)):
        return tc.S1F
    return tc.S1 if (cin, cout) == (128, 128) else None


def _conv_tc(kind, x, w, cout):
    """x (B,Cin,D,H,W) fp32, w in the layout pack_weight(kind) expects -> (B,Cout,...) fp32, on the tensor cores (bf16x3)."""
    xs = tc.to_blocked_bf16(x, s2d=(kind == tc.S2), split=True)
    return tc.conv3d_tc_split(kind, xs, tc.pack_weight_split(w, kind), cout, out_mode=tc.F32)


class _Conv3dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, stride):
        k = _geometry(w, stride)
        x = x.contiguous().float()
        ctx.save_for_backward(x, w)
        ctx.stride, ctx.k = stride, k
        cout, cin = w.shape[:2]
        kind = _tc_kind(cin, cout, k, stride)
        if kind is not None and not (stride == 2 and any(n % 2 for n in x.shape[2:])):
            return _conv_tc(kind, x, w.detach().float(), cout)
        return ops.conv3d_f32(x, ops.pack_conv3d_weight(w.detach().float()), k=k, stride=stride)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        k, stride = ctx.k, ctx.stride
        dy = dy.contiguous().float()
        wf = w.detach().float()
        cout, cin = w.shape[:2]
        dx = dw = None
        if ctx.needs_input_grad[0]:
            if stride == 1:      # correlation with the flipped kernel, channels swapped: a plain k s1 conv of dY
                wt = wf.flip(2, 3, 4).transpose(0, 1).contiguous()
                kind = _tc_kind(cout, cin, k, 1)
                dx = _conv_tc(kind, dy, wt, cin) if kind is not None else ops.conv3d_f32(dy, ops.pack_conv3d_weight(wt), k=k, stride=1)
            else:                # k3 s2 p1 on even dims: dX = conv_transpose3d(dY, W, 2, 1, output_padding 1); W already has that layout
                if k != 3 or any(n % 2 for n in x.shape[2:]):
                    raise NotImplementedError("conv3d (training): stride 2 needs k = 3 and even input dims")
                kind = _tc_kind(cout, cin, 3, 2, transposed=True)
                dx = (_conv_tc(kind, dy, wf, cin) if kind is not None
                      else ops.conv3d_f32(dy, ops.pack_conv3d_weight(wf, transposed=True), k=3, stride=2, transposed=True))
        if ctx.needs_input_grad[1]:
            dev = x.device
            B, Cin, Di, Hi, Wi = x.shape
            Cout = w.shape[0]
            dwp = torch.zeros((k ** 3, Cin, Cout), device=dev, dtype=torch.float32)
            _call("ss_conv3d_wgrad_f32", dev, _ptr(x), _ptr(dy), _ptr(dwp), B, Cin, Cout, Di, Hi, Wi, int(k), int(stride))
            dw = dwp.permute(2, 1, 0).reshape(Cout, Cin, k, k, k).to(w.dtype)          # [tap][ci][co] -> (Cout,Cin,kd,kh,kw)
        return dx, dw, None


def conv3d(x, weight, stride=1):
    """Differentiable Conv3d(k, stride, padding=k//2, bias=False) on the CUDA kernels (forward, dX and dW)."""
    return _Conv3dFn.apply(x, weight, stride)


def _ws_bytes(C):
    from . import _lib
    return _lib.load().ss_bn_workspace_bytes(int(C))


class _BatchNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        dev = _require_cuda(x)
        x = x.contiguous()
        B, C = x.shape[:2]
        S = x.numel() // (B * C)
        out = torch.empty_like(x)
        mean = torch.empty(C, device=dev, dtype=torch.float32)
        var = torch.empty(C, device=dev, dtype=torch.float32)
        wd = None if weight is None else weight.detach().float().contiguous()
        bd = None if bias is None else bias.detach().float().contiguous()
        ws = torch.empty(_ws_bytes(C), device=dev, dtype=torch.uint8)
        _call("ss_bn_train_forward", dev, _ptr(x), _ptr(wd), _ptr(bd), _ptr(out), _ptr(mean), _ptr(var), _ptr(ws), B, C, ctypes.c_longlong(S),
              ctypes.c_float(eps), 0)
        ctx.save_for_backward(x, mean, var, wd if wd is not None else x.new_empty(0))
        ctx.eps, ctx.has_w, ctx.has_b = eps, weight is not None, bias is not None
        ctx.mark_non_differentiable(mean, var)
        return out, mean, var

    @staticmethod
    def backward(ctx, dy, _dm, _dv):
        x, mean, var, wd = ctx.saved_tensors
        dev = x.device
        B, C = x.shape[:2]
        S = x.numel() // (B * C)
        dy = dy.contiguous().float()
        dx = torch.empty_like(x)
        dw = torch.empty(C, device=dev, dtype=torch.float32)
        db = torch.empty(C, device=dev, dtype=torch.float32)
        ws = torch.empty(_ws_bytes(C), device=dev, dtype=torch.uint8)
        _call("ss_bn_train_backward", dev, _ptr(x), _ptr(dy), _ptr(mean), _ptr(var), _ptr(wd if ctx.has_w else None), _ptr(dx), _ptr(dw), _ptr(db),
              _ptr(ws), B, C, ctypes.c_longlong(S), ctypes.c_float(ctx.eps))
        return dx, (dw if ctx.has_w else None), (db if ctx.has_b else None), None


def batch_norm_train(x, bn: torch.nn.modules.batchnorm._BatchNorm):
    """nn.BatchNorm{2,3}d in training mode on the CUDA kernels: batch statistics, differentiable w.r.t. x / weight / bias, and
    the module's running statistics updated as torch does (momentum; unbiased variance; num_batches_tracked)."""
    out, mean, var = _BatchNormFn.apply(x.float(), bn.weight, bn.bias, bn.eps)
    if bn.track_running_stats and bn.running_mean is not None:
        with torch.no_grad():
            n = x.numel() // x.shape[1]
            bn.num_batches_tracked += 1
            mom = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
            bn.running_mean.mul_(1 - mom).add_(mean, alpha=mom)
            bn.running_var.mul_(1 - mom).add_(var * (n / max(n - 1, 1)), alpha=mom)
    return out


# ---------------------------------------------------------------------------------------------------------------------
# attention_block and SSR_upsample in training mode: compositions of differentiable kernels
# ---------------------------------------------------------------------------------------------------------------------
class _AttnCoreFn(torch.autograd.Function):
    """The softmax core.  valid = (H0, W0): the masked branch -- qkv is that of the volume zero-padded on BOTH axes, tokens beyond
    (H0, W0) are padding and scores between padded and real tokens get -1000 (submodule_other.py:822-829); None: nothing is masked."""

    @staticmethod
    def forward(ctx, qkv, block, heads, valid=None):
        dev = _require_cuda(qkv)
        qkv = qkv.contiguous()
        B, C3, D, H, W = qkv.shape
        out = torch.empty((B, C3 // 3, D, H, W), device=dev, dtype=torch.float32)
        if valid is None:
            _call("ss_window_attention_core_f32_out", dev, _ptr(qkv), _ptr(out), B, C3 // 3, D, H, W, int(block[0]), int(block[1]), int(block[2]),
                  int(heads))
        else:
            _call("ss_window_attention_core_f32_masked", dev, _ptr(qkv), _ptr(out), B, C3 // 3, D, H, W, int(block[0]), int(block[1]),
                  int(block[2]), int(heads), int(valid[0]), int(valid[1]))
        ctx.save_for_backward(qkv)
        ctx.block, ctx.heads = tuple(int(v) for v in block), int(heads)
        ctx.valid = (H, W) if valid is None else (int(valid[0]), int(valid[1]))
        return out

    @staticmethod
    def backward(ctx, dout):
        (qkv,) = ctx.saved_tensors
        B, C3, D, H, W = qkv.shape
        dq = torch.empty_like(qkv)
        _call("ss_window_attention_core_backward_masked", qkv.device, _ptr(qkv), _ptr(dout.contiguous().float()), _ptr(dq), B, C3 // 3, D, H, W,
              ctx.block[0], ctx.block[1], ctx.block[2], ctx.heads, ctx.valid[0], ctx.valid[1])
        return dq, None, None, None


def attention_block_train(mod, x):
    """attention_block.forward (submodule_other.py:805-837) with gradients: the qkv Linear and the final 1x1x1 conv are k = 1 layers
    of the differentiable conv3d above (+ bias), the softmax core in between has its own forward / backward kernels."""
    block = mod.block if isinstance(mod.block, (tuple, list)) else (mod.block,) * 3
    C = mod.qkv_3d.in_features
    H0, W0 = x.shape[3], x.shape[4]
    pb, pr = (-H0) % block[1], (-W0) % block[2]
    if pb or pr:        # zero-pad before the qkv Linear (padded tokens carry the bias), crop before the final conv
        x = torch.nn.functional.pad(x, (0, pr, 0, pb))
    qkv = conv3d(x, mod.qkv_3d.weight.view(3 * C, C, 1, 1, 1), 1) + mod.qkv_3d.bias.view(1, -1, 1, 1, 1)
    # one padded axis: the reference masks nothing (ops.window_pad); both: -1000 between padded and real tokens (submodule_other.py:822-829)
    o = _AttnCoreFn.apply(qkv, block, mod.num_heads, (H0, W0) if (pb and pr) else None)[:, :, :, :H0, :W0]
    return conv3d(o, mod.final1x1.weight, 1) + mod.final1x1.bias.view(1, -1, 1, 1, 1)


class _Up4Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        dev = _require_cuda(x)
        x = x.contiguous()
        h, w = x.shape[-2:]
        planes = x.numel() // (h * w)
        out = torch.empty(tuple(x.shape[:-2]) + (4 * h, 4 * w), device=dev, dtype=torch.float32)
        _call("ss_bilinear_up4", dev, _ptr(x), _ptr(out), planes, h, w)
        ctx.shape = tuple(x.shape)
        return out

    @staticmethod
    def backward(ctx, g):
        h, w = ctx.shape[-2:]
        gin = torch.empty(ctx.shape, device=g.device, dtype=torch.float32)
        _call("ss_bilinear_up4_backward", g.device, _ptr(g.contiguous().float()), _ptr(gin), gin.numel() // (h * w), h, w)
        return gin


class _SmallConv2dFn(torch.autograd.Function):
    """Conv2d with Cin, Cout <= 8 (k in {1,3}, stride 1, pad k//2) on the small-channel kernels of csrc/train.cu."""

    @staticmethod
    def _run(x, w):
        dev = _require_cuda(x, w)
        B, Cin, H, W = x.shape
        Cout, k = w.shape[0], w.shape[-1]
        out = torch.empty((B, Cout, H, W), device=dev, dtype=torch.float32)
        _call("ss_conv2d_small_f32", dev, _ptr(x), _ptr(w), _ptr(out), B, Cin, Cout, H, W, int(k))
        return out

    @staticmethod
    def forward(ctx, x, w):
        x = x.contiguous().float()
        ctx.save_for_backward(x, w)
        return _SmallConv2dFn._run(x, w.detach().float().contiguous())

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.contiguous().float()
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = _SmallConv2dFn._run(dy, w.detach().float().flip(2, 3).transpose(0, 1).contiguous())
        if ctx.needs_input_grad[1]:
            B, Cin, H, W = x.shape
            Cout, k = w.shape[0], w.shape[-1]
            dwf = torch.zeros((Cout, Cin, k, k), device=x.device, dtype=torch.float32)
            _call("ss_conv2d_small_wgrad_f32", x.device, _ptr(x), _ptr(dy), _ptr(dwf), B, Cin, Cout, H, W, int(k))
            dw = dwf.to(w.dtype)
        return dx, dw


def _small_conv_ok(cin, cout, k):
    from . import _lib
    return cin <= 8 and cout <= 8 and bool(_lib.load().ss_conv2d_small_wgrad_supported(int(cin), int(cout), int(k)))


def conv2d(x, weight, bias=None):
    """Differentiable Conv2d (k in {1,3}, stride 1, padding k//2).  The SSR_upsample shapes (1 -> 6 3x3, 6 -> 6 / 6 -> 1 1x1) run on
    the small-channel kernels; everything else goes through the 3-D kernels: the image is a depth-1 volume, a 3x3 kernel is the
    centre depth plane of a 3x3x3 one."""
    k = weight.shape[-1]
    if _small_conv_ok(weight.shape[1], weight.shape[0], k):
        y = _SmallConv2dFn.apply(x, weight)
        return y if bias is None else y + bias.view(1, -1, 1, 1)
    w3 = weight.unsqueeze(2)
    if k == 3:
        w3 = torch.nn.functional.pad(w3, (0, 0, 0, 0, 1, 1))
    y = conv3d(x.unsqueeze(2), w3, 1).squeeze(2)
    return y if bias is None else y + bias.view(1, -1, 1, 1)


def ssr_upsample_train(mod, depth_low, weights, pred_label):
    """SSR_upsample.forward (submodule.py:421-431) in TRAINING mode (its four BatchNorm2d layers use batch statistics, which rules
    out the fused inference kernel): bilinear x4, the convolutions and the BatchNorms are differentiable kernels of this file, the
    pointwise glue (softmax over the 6 classes, sigmoids, products) is torch, as in the reference."""
    lab = torch.softmax(pred_label, dim=1)
    d_up = _Up4Fn.apply(depth_low.float())
    d = batch_norm_train(d_up, mod.conv[0])
    d = batch_norm_train(conv2d(d, mod.conv[1].weight, mod.conv[1].bias), mod.conv[2])
    g = torch.sigmoid(batch_norm_train(conv2d(lab * weights, mod.conv1[0].weight, mod.conv1[0].bias), mod.conv1[1]))
    g = torch.sigmoid(batch_norm_train(conv2d(g * weights, mod.conv2[0].weight, mod.conv2[0].bias), mod.conv2[1]))
    res = conv2d(d * g, mod.conv3.weight, mod.conv3.bias)
    return (d_up + res).squeeze(1)
