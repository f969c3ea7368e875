"""Autograd for the 3-D convolution stack in TRAINING mode (BASELINE config #5): what
`convbn_3d` / `BasicConv(is_3d)` (models/submodule_other.py:845-848, models/submodule.py:89-116) do when the reference trains
(main_us3d.py:186-222) -- Conv3d and BatchNorm3d with batch statistics, forward AND backward on the CUDA kernels (fp32):

  conv3d        forward  ss_conv3d_f32                       (csrc/conv3d_f32.cu, the inference kernel without the folded BN)
                dX       ss_conv3d_f32 on dY with a re-packed weight: flipped taps + swapped channels (k3 s1), the
                         ConvTranspose3d(k3,s2,p1,op1) phase GEMMs with the same weight (k3 s2), W^T (k1)
                dW       ss_conv3d_wgrad_f32                 (csrc/train.cu)
  batch_norm    forward / backward ss_bn_train_forward / ss_bn_train_backward, running statistics updated like nn.BatchNorm3d
                (momentum, unbiased variance)

The stateless operators of the surface (volume builders, regression, warps, propagation) already carry their backward kernels
(torch_ops.py / csrc/backward.cu).  NOT native yet in training mode: `attention_block` and `SSR_upsample` (they raise), so a
full training step of the reference model still needs torch modules for those two (tools/train_step.py states which)."""
from __future__ import annotations

import ctypes

import torch

from . import ops
from .ops import _call, _ptr, _require_cuda


def _geometry(w, stride):
    k = w.shape[2]
    if tuple(w.shape[2:]) != (k, k, k) or k not in (1, 3) or stride not in (1, 2):
        raise NotImplementedError("conv3d (training): only k in {1,3} cubic kernels with pad k//2 and stride in {1,2}")
    return k


class _Conv3dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, stride):
        k = _geometry(w, stride)
        x = x.contiguous().float()
        ctx.save_for_backward(x, w)
        ctx.stride, ctx.k = stride, k
        return ops.conv3d_f32(x, ops.pack_conv3d_weight(w.detach().float()), k=k, stride=stride)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        k, stride = ctx.k, ctx.stride
        dy = dy.contiguous().float()
        wf = w.detach().float()
        dx = dw = None
        if ctx.needs_input_grad[0]:
            if stride == 1:      # correlation with the flipped kernel, channels swapped: a plain k s1 conv of dY
                dx = ops.conv3d_f32(dy, ops.pack_conv3d_weight(wf.flip(2, 3, 4).transpose(0, 1).contiguous()), k=k, stride=1)
            else:                # k3 s2 p1 on even dims: dX = conv_transpose3d(dY, W, 2, 1, output_padding 1); W already has that layout
                if k != 3 or any(n % 2 for n in x.shape[2:]):
                    raise NotImplementedError("conv3d (training): stride 2 needs k = 3 and even input dims")
                dx = ops.conv3d_f32(dy, ops.pack_conv3d_weight(wf, transposed=True), k=3, stride=2, transposed=True)
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
        _call("ss_bn_train_forward", dev, _ptr(x), _ptr(wd), _ptr(bd), _ptr(out), _ptr(mean), _ptr(var), B, C, ctypes.c_longlong(S),
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
        _call("ss_bn_train_backward", dev, _ptr(x), _ptr(dy), _ptr(mean), _ptr(var), _ptr(wd if ctx.has_w else None), _ptr(dx), _ptr(dw), _ptr(db),
              B, C, ctypes.c_longlong(S), ctypes.c_float(ctx.eps))
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
