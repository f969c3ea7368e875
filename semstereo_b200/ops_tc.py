"""Wrappers for the tensor-core (tcgen05 / TMEM / TMA) 3-D convolutions and their bf16 "blocked channels" layouts.

Layouts (bf16): blocked  = (B, C/8, D, H, W, 8);  phase-split ("s2d") = (B, 8, C/8, D/2, H/2, W/2, 8) with
phase = 4*(d&1) + 2*(h&1) + (w&1).  See include/semstereo_b200.h and csrc/conv3d_tc.cu.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from .ops import _call, _ptr, _require_cuda

S1, K1, S2, T2, C2D, S1F = 0, 1, 2, 3, 4, 5   # layer kinds of ss_conv3d_tc (C2D: Conv2d 3x3 on a depth-1 volume; S1F: S1 with depth taps folded into N)


def _require_bf16(t, ndim):
    if not t.is_cuda:
        raise RuntimeError("semstereo_b200 kernels need CUDA tensors: there is no CPU fallback on this path")
    if t.dtype != torch.bfloat16 or t.dim() != ndim or t.shape[-1] != 8 or not t.is_contiguous():
        raise ValueError(f"expected a contiguous bf16 tensor with {ndim} dims and 8 innermost channels")
    return t.device


def to_blocked_bf16(x, s2d=False, split=False, out=None):
    """split=True: the hi/lo pair of the bf16x3 route stacked on the batch axis -> batch 2B (see conv3d_tc_split).
    out: write into this (contiguous, right-shaped) tensor, e.g. one half of a batch-stacked buffer."""
    dev = _require_cuda(x)
    B, C, D, H, W = x.shape
    nb = 2 * B if split else B
    shape = (nb, 8, C // 8, D // 2, H // 2, W // 2, 8) if s2d else (nb, C // 8, D, H, W, 8)
    if out is None:
        out = torch.empty(shape, device=dev, dtype=torch.bfloat16)
    elif tuple(out.shape) != shape or out.dtype != torch.bfloat16 or not out.is_contiguous() or out.device != dev:
        raise ValueError("to_blocked_bf16: `out` must be a contiguous bf16 tensor of shape %s" % (shape,))
    _call("ss_to_blocked_bf16_ex", dev, _ptr(x), _ptr(out), B, C, D, H, W, int(s2d), int(split))
    return out


def to_blocked_tri(x):
    """fp32 (B,C,H,W) -> bf16 (B, 3C/8, H, W, 8) = [hi | lo | hi] chunks: the K-concat operand of pointwise_split."""
    dev = _require_cuda(x)
    B, C, H, W = x.shape
    out = torch.empty((B, 3 * C // 8, H, W, 8), device=dev, dtype=torch.bfloat16)
    _call("ss_to_blocked_bf16_ex", dev, _ptr(x), _ptr(out), B, C, 1, H, W, 0, 2)
    return out


def pack_pointwise_split(w):
    """1x1 conv weight (Cout,Cin[,1,1[,1]]) fp32 -> the CONV1 packing of [w_hi | w_hi | w_lo] (K = 3*Cin) for pointwise_split."""
    w = w.detach().float().reshape(w.shape[0], -1)
    hi, lo = split_f32(w)
    return pack_weight2d(torch.cat((hi, hi, lo), 1).reshape(w.shape[0], -1, 1, 1), CONV1)


def pointwise_split(x_tri, w_tri, cout, scale=None, shift=None, relu=False):
    """fp32-accurate 1x1 convolution on the tensor cores: x_tri from to_blocked_tri (or a kernel that writes that form), w_tri
    from pack_pointwise_split; one GEMM with K = 3*Cin = x_hi*w_hi + x_lo*w_hi + x_hi*w_lo.  Returns fp32 (B,Cout,H,W)."""
    return conv2d_tc(CONV1, x_tri, w_tri, cout, scale, shift, relu=relu, out_f32=True)


def widen_bf16(x, out=None):
    """bf16 tensor -> fp32 tensor of the same shape (our kernel; used where features cross PCIe as bf16)."""
    if not x.is_cuda or x.dtype != torch.bfloat16 or not x.is_contiguous() or x.numel() % 8:
        raise ValueError("widen_bf16: contiguous CUDA bf16 tensor with a multiple of 8 elements expected")
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=torch.float32)
    _call("ss_widen_bf16", x.device, _ptr(x), _ptr(out), x.numel())
    return out


def from_blocked_bf16(xb):
    dev = _require_bf16(xb, 6)
    B, C8, D, H, W, _ = xb.shape
    out = torch.empty((B, C8 * 8, D, H, W), device=dev, dtype=torch.float32)
    _call("ss_from_blocked_bf16", dev, _ptr(xb), _ptr(out), B, C8 * 8, D, H, W)
    return out


def blocked_to_s2d(xb):
    dev = _require_bf16(xb, 6)
    B, C8, D, H, W, _ = xb.shape
    out = torch.empty((B, 8, C8, D // 2, H // 2, W // 2, 8), device=dev, dtype=torch.bfloat16)
    _call("ss_blocked_to_s2d", dev, _ptr(xb), _ptr(out), B, C8 * 8, D, H, W)
    return out


def s2d_as_batch(xs):
    """View a phase-split tensor as a blocked tensor with batch B*8 (position-wise layers do not care)."""
    B, P, C8, D, H, W, _ = xs.shape
    return xs.view(B * P, C8, D, H, W, 8)


def ntile(kind, cin, cout):
    return _lib.load().ss_conv3d_tc_ntile(int(kind), int(cin), int(cout))


def pack_weight(w, kind):
    """fp32 conv weight -> bf16 [ceil(Cout/N)][taps][Cin/8][N][8], Cout zero-padded to the kernel's tile N.
    w: (Cout,Cin,k,k,k) for kinds S1/K1/S2, ConvTranspose3d (Cin,Cout,3,3,3) for T2."""
    if kind == T2:
        w = w.permute(1, 0, 2, 3, 4)                                   # -> (Cout, Cin, kd, kh, kw), taps index the weight directly
    if kind == C2D:
        w = w.reshape(w.shape[0], w.shape[1], 1, 3, 3)                 # Conv2d (Cout,Cin,3,3): 9 in-plane taps
    if kind == S1F:
        # depth taps folded into N: [9 in-plane taps][Cin/8][3*Cout (j*Cout + co, kd = 2 - j)][8]
        cout, cin = w.shape[:2]
        if ntile(kind, cin, cout) != cout:
            raise NotImplementedError(f"conv3d_tc: kind {kind} with (Cin={cin}, Cout={cout}) has no tensor-core configuration")
        t = w.flip(2).reshape(cout, cin // 8, 8, 3, 9).permute(4, 1, 3, 0, 2)        # (t9, chunk, j, co, c8)
        return t.reshape(9, cin // 8, 3 * cout, 8).contiguous().to(torch.bfloat16)
    cout, cin = w.shape[:2]
    taps = w.shape[2] * w.shape[3] * w.shape[4]
    n = ntile(kind, cin, cout)
    if n == 0:
        raise NotImplementedError(f"conv3d_tc: kind {kind} with (Cin={cin}, Cout={cout}) has no tensor-core configuration")
    nt = -(-cout // n)
    wp = w.new_zeros((nt * n, cin, taps))
    wp[:cout] = w.reshape(cout, cin, taps)
    t = wp.reshape(nt, n, cin // 8, 8, taps).permute(0, 4, 2, 1, 3)     # (ntile, tap, chunk, n, c8)
    return t.contiguous().to(torch.bfloat16)


def gate_sigmoid_blocked(gate_logits):
    """sigmoid of the channelAtt logits (B,C,H,W) as fp32 (B,C/8,H,W,8): what the conv epilogue multiplies with."""
    dev = _require_cuda(gate_logits)
    B, C, H, W = gate_logits.shape
    out = torch.empty((B, C // 8, H, W, 8), device=dev, dtype=torch.float32)
    _call("ss_gate_sigmoid_blocked", dev, _ptr(gate_logits), _ptr(out), B, C, H, W)
    return out


def patch_gate_blocked(volume, patch_w, gate_logits, split=False):
    """`patch` depthwise conv * sigmoid(gate) written straight into the phase-split bf16 layout the stride-2 layer reads
    (split=True: as the hi/lo pair of the bf16x3 route, batch 2B)."""
    dev = _require_cuda(volume, patch_w, gate_logits)
    B, G, D, H, W = volume.shape
    if patch_w.numel() != G * 9 or tuple(gate_logits.shape) != (B, G, H, W):
        raise ValueError("patch_gate_blocked: patch weight (G,9) and gate logits (B,G,H,W) expected")
    out = torch.empty(((2 * B if split else B), 8, G // 8, D // 2, H // 2, W // 2, 8), device=dev, dtype=torch.bfloat16)
    _call("ss_patch_gate_blocked_ex", dev, _ptr(volume), _ptr(patch_w), _ptr(gate_logits), _ptr(out), B, G, D, H, W, int(split))
    return out


def sparse_concat_volume_blocked(cf_l, cf_r, disp_topk, att_topk=None):
    dev = _require_cuda(cf_l, cf_r, disp_topk, att_topk)
    B, C, H, W = cf_l.shape
    K = disp_topk.shape[1]
    if cf_r.shape != cf_l.shape or tuple(disp_topk.shape) != (B, K, H, W):
        raise ValueError("sparse_concat_volume_blocked: shape mismatch")
    out = torch.empty((B, 2 * C // 8, K, H, W, 8), device=dev, dtype=torch.bfloat16)
    _call("ss_sparse_concat_volume_blocked", dev, _ptr(cf_l), _ptr(cf_r), _ptr(disp_topk), _ptr(att_topk), _ptr(out), B, C, K, H, W)
    return out


def window_attention_core(qkv_blocked, block, num_heads=16):
    """softmax(q k^T * hd^-0.5) v per (window, head) on the blocked layout: (B,48,D,H,W,8) -> (B,16,D,H,W,8)."""
    dev = _require_bf16(qkv_blocked, 6)
    B, C3, D, H, W, _ = qkv_blocked.shape
    if C3 % 3:
        raise ValueError("window_attention_core: qkv must have 3*C/8 channel chunks")
    out = torch.empty((B, C3 // 3, D, H, W, 8), device=dev, dtype=torch.bfloat16)
    _call("ss_window_attention_core_blocked", dev, _ptr(qkv_blocked), _ptr(out), B, (C3 // 3) * 8, D, H, W,
          int(block[0]), int(block[1]), int(block[2]), int(num_heads))
    return out


def window_attention_core_f32(qkv, block, num_heads=16):
    """fp32 softmax(q k^T * hd^-0.5) v per (window, head): qkv fp32 (B,3C,D,H,W) -> bf16 (B,3C/8,D,H,W,8) in the [hi | lo | hi]
    K-concat form (the operand of pointwise_split for the final 1x1x1 conv)."""
    dev = _require_cuda(qkv)
    B, C3, D, H, W = qkv.shape
    if C3 % 3:
        raise ValueError("window_attention_core_f32: qkv must have 3*C channels")
    C = C3 // 3
    out = torch.empty((B, 3 * C // 8, D, H, W, 8), device=dev, dtype=torch.bfloat16)
    _call("ss_window_attention_core_f32", dev, _ptr(qkv), _ptr(out), B, C, D, H, W, int(block[0]), int(block[1]), int(block[2]),
          int(num_heads))
    return out


def pack_head_weight(w):
    """(1,32,3,3,3) fp32 -> bf16 [4][48][8]: rows j*16 + t9 (depth tap kd = 2 - j, 9 in-plane taps + 7 zero rows), K-major chunks
    of 8 input channels (csrc/conv3d_tc_head.cu)."""
    if tuple(w.shape) != (1, 32, 3, 3, 3):
        raise NotImplementedError("conv3d_tc_head: only Conv3d(32, 1, 3) is supported")
    t = w.new_zeros((32, 3, 16))                   # [c][j][t9]
    t[:, :, :9] = w[0].flip(1).reshape(32, 3, 9)
    return t.reshape(4, 8, 48).permute(0, 2, 1).contiguous().to(torch.bfloat16)


def conv3d_tc_head(xb, w_head, acc_in=None, acc_in2=None, in_split=False):
    """Cout = 1 classifier head on the blocked layout: (B,4,D,H,W,8) bf16 -> (B,1,D,H,W) fp32 (+ fp32 partial sums acc_in*).
    in_split: xb is a split tensor (batch 2B), w_head = pack_head_weight_split(w): the fp32-accurate bf16x3 product in one launch."""
    dev = _require_bf16(xb, 6)
    B, C8, D, H, W, _ = xb.shape
    if in_split:
        B //= 2
    if w_head.dtype != torch.bfloat16 or tuple(w_head.shape) != ((2, 4, 48, 8) if in_split else (4, 48, 8)) or not w_head.is_contiguous():
        raise ValueError("conv3d_tc_head: weight must come from pack_head_weight(_split)")
    for a in (acc_in, acc_in2):
        if a is not None and (_require_cuda(a) != dev or a.dtype != torch.float32 or a.numel() != B * D * H * W or not a.is_contiguous()):
            raise ValueError("conv3d_tc_head: acc_in must be a contiguous fp32 (B,1,D,H,W) tensor")
    out = torch.empty((B, 1, D, H, W), device=dev, dtype=torch.float32)
    _call("ss_conv3d_tc_head_ex", dev, _ptr(xb), _ptr(w_head), _ptr(acc_in), _ptr(acc_in2), _ptr(out), int(in_split), B, C8 * 8, D, H, W)
    return out


BLOCKED, F32, S2D, F32B4 = 0, 1, 2, 3          # out_mode of ss_conv3d_tc (F32B4: fp32 (B,Cout/4,D,H,W,4), the partial-sum layout)


def pack_skip_weight(w, scale=None):
    """1x1x1 redir conv weight (C,C[,1,1,1]) (with its BN scale folded in when given) -> bf16 [C/8][C][8] (K-major rows = output
    channels)."""
    c = w.shape[0]
    w = w.reshape(c, c)
    if scale is not None:
        w = w * scale.reshape(c, 1)
    return w.reshape(c, c // 8, 8).permute(1, 0, 2).contiguous().to(torch.bfloat16)      # (chunk, cout, c8)


def split_f32(w):
    """fp32 tensor -> (hi, lo) fp32 tensors with hi = bf16(w), lo = w - hi (bf16(lo) is the second term of the bf16x3 split)."""
    hi = w.to(torch.bfloat16).float()
    return hi, w - hi


def split_supported(kind, cin, cout):
    """Does this layer have an in-kernel (single launch) bf16x3 split configuration?"""
    return bool(_lib.load().ss_conv3d_tc_split_supported(int(kind), int(cin), int(cout)))


def pack_weight_split(w, kind):
    """fp32 weight -> (hi, lo) packings of pack_weight for the bf16x3 split route; for layers with an in-kernel split
    configuration a third tensor: per tap [hi chunks | lo chunks] (what ss_conv3d_tc_ex(in_split=1) reads), else None."""
    hi, lo = split_f32(w)
    ph, pl = pack_weight(hi, kind), pack_weight(lo, kind)
    cin, cout = (w.shape[0], w.shape[1]) if kind == T2 else (w.shape[1], w.shape[0])
    both = torch.cat((ph, pl), dim=1 if kind == S1F else 2).contiguous() if split_supported(kind, cin, cout) else None
    return ph, pl, both


def pack_skip_weight_split(w):
    """(C,C) redir weight (BN scale folded) -> (hi, lo, [hi chunks | lo chunks]) packings of pack_skip_weight."""
    hi, lo = split_f32(w)
    ph, pl = pack_skip_weight(hi), pack_skip_weight(lo)
    return ph, pl, torch.cat((ph, pl), dim=0).contiguous()


def pack_head_weight_split(w):
    hi, lo = split_f32(w)
    return torch.stack((pack_head_weight(hi), pack_head_weight(lo))).contiguous()        # (2,4,48,8)


def conv3d_tc_split(kind, xs, w_split, cout, scale=None, shift=None, gate_blocked=None, residual_s2d=None, relu=False,
                    out_mode=BLOCKED, skip_split=None):
    """fp32-accurate layer on the bf16 tensor cores (bf16x3: x*w = x_hi*w_hi + x_lo*w_hi + x_hi*w_lo, fp32 accumulation).
    xs (and residual_s2d) are split tensors = the hi/lo halves stacked on the batch axis (batch 2B); w_split / skip_split come
    from pack_weight_split / pack_skip_weight_split.  Layers with an in-kernel configuration issue the three MMAs per K step
    into one TMEM accumulator (one launch).  The others take two launches of the plain kernel: (1) the 2B-batch input x w_hi,
    raw fp32 partial sums; (2) the hi half x w_lo plus the partial sums, then the usual epilogue.  The bf16 outputs are split
    tensors again (batch 2B); out_mode F32 returns the plain (B,Cout,...) fp32 result."""
    B = xs.shape[0] // 2
    w_hi, w_lo, w_both = w_split
    s_hi, s_lo, s_both = skip_split if skip_split is not None else (None, None, None)
    if w_both is not None:
        return conv3d_tc(kind, xs, w_both, cout, scale, shift, gate_blocked, residual_s2d, relu=relu, out_mode=out_mode,
                         skip_weight=s_both, out_split=out_mode != F32, in_split=True)
    part = conv3d_tc(kind, xs, w_hi, cout, residual_s2d=residual_s2d, skip_weight=s_hi, out_mode=F32B4)
    return conv3d_tc(kind, xs[:B], w_lo, cout, scale, shift, gate_blocked, None if residual_s2d is None else residual_s2d[:B],
                     relu=relu, out_mode=out_mode, skip_weight=s_lo, acc_in=part[:B], acc_in2=part[B:], out_split=out_mode != F32)


def conv3d_tc(kind, xb, w_tc, cout, scale=None, shift=None, gate_blocked=None, residual_s2d=None, relu=False, out_mode=BLOCKED,
              skip_weight=None, acc_in=None, acc_in2=None, out_split=False, in_split=False):
    """xb: blocked (kinds S1, K1, T2) or phase-split (kind S2) bf16 input.  Returns bf16 blocked / phase-split or fp32 NCDHW.
    acc_in / acc_in2 / out_split / in_split: the hooks of the bf16x3 split route (conv3d_tc_split)."""
    if kind == S2:
        dev = _require_bf16(xb, 7)
        B, _, C8, D2, H2, W2, _ = xb.shape
        D, H, W = 2 * D2, 2 * H2, 2 * W2
        Do, Ho, Wo = D2, H2, W2
    else:
        dev = _require_bf16(xb, 6)
        B, C8, D, H, W, _ = xb.shape
        Do, Ho, Wo = (2 * D, 2 * H, 2 * W) if kind == T2 else (D, H, W)
    nparts = 2 if in_split else 1
    if in_split:
        if B % 2:
            raise ValueError("conv3d_tc: a split input has an even batch (hi batches | lo batches)")
        B //= 2
    cin = C8 * 8
    n = ntile(kind, cin, cout)
    taps = 1 if kind == K1 else (9 if kind == C2D else 27)
    wshape = (9, nparts * C8, 3 * cout, 8) if kind == S1F else (-(-cout // max(n, 1)), taps, nparts * C8, n, 8)
    if n == 0 or w_tc.dtype != torch.bfloat16 or tuple(w_tc.shape) != wshape or not w_tc.is_contiguous():
        raise ValueError("conv3d_tc: weight must come from pack_weight(w, kind) for this layer")
    for t in (scale, shift, gate_blocked):
        if t is not None:
            _require_cuda(t)
    if gate_blocked is not None and tuple(gate_blocked.shape) != (B, cout // 8, Ho, Wo, 8):
        raise ValueError("conv3d_tc: gate must be fp32 (B,Cout/8,Ho,Wo,8) from gate_sigmoid_blocked")
    if residual_s2d is not None:
        _require_bf16(residual_s2d, 7)
        if kind != T2 or tuple(residual_s2d.shape) != (nparts * B, 8, cout // 8, D, H, W, 8):
            raise ValueError("conv3d_tc: residual must be phase-split (B,8,Cout/8,D,H,W,8) at the transposed layer's input dims")
    for a in (acc_in, acc_in2):
        if a is not None and (_require_cuda(a) != dev or a.dtype != torch.float32 or not a.is_contiguous()
                              or tuple(a.shape) != (B, cout // 4, Do, Ho, Wo, 4)):
            raise ValueError("conv3d_tc: acc_in must be a contiguous fp32 (B,Cout/4,Do,Ho,Wo,4) tensor (out_mode F32B4)")
    if acc_in2 is not None and acc_in is None:
        raise ValueError("conv3d_tc: acc_in2 needs acc_in")
    nb = 2 * B if out_split else B
    if out_mode in (F32, F32B4):
        if out_split:
            raise ValueError("conv3d_tc: a split output exists only for the bf16 layouts")
        out = torch.empty((B, cout, Do, Ho, Wo) if out_mode == F32 else (B, cout // 4, Do, Ho, Wo, 4), device=dev, dtype=torch.float32)
    elif out_mode == S2D:
        out = torch.empty((nb, 8, cout // 8, Do // 2, Ho // 2, Wo // 2, 8), device=dev, dtype=torch.bfloat16)
    else:
        out = torch.empty((nb, cout // 8, Do, Ho, Wo, 8), device=dev, dtype=torch.bfloat16)
    if skip_weight is not None and (residual_s2d is None or skip_weight.dtype != torch.bfloat16
                                    or tuple(skip_weight.shape) != (nparts * cout // 8, cout, 8) or not skip_weight.is_contiguous()):
        raise ValueError("conv3d_tc: skip_weight must come from pack_skip_weight and needs the skip input as residual_s2d")
    _call("ss_conv3d_tc_ex", dev, int(kind), _ptr(xb), _ptr(w_tc), _ptr(scale), _ptr(shift), _ptr(gate_blocked), _ptr(residual_s2d),
          _ptr(skip_weight), _ptr(acc_in), _ptr(acc_in2), _ptr(out), int(out_mode), int(out_split), int(in_split), B, cin, cout, D, H, W,
          int(relu))
    return out


# ---------------------------------------------------------------------------------------------------------------------
# 2-D decoder convolutions (csrc/conv2d_tc.cu): blocked bf16 (B, C/8, H, W, 8)
# ---------------------------------------------------------------------------------------------------------------------
CONV3, CONV1, DECONV4 = 0, 1, 2      # modes of ss_conv2d_tc
_DECONV_SLABS = (((0, 0), (0, 1, 2, 3)), ((-1, 0), (0, 1)), ((1, 0), (2, 3)), ((0, -1), (0, 2)), ((0, 1), (1, 3)),
                 ((-1, -1), (0,)), ((-1, 1), (1,)), ((1, -1), (2,)), ((1, 1), (3,)))
_DECONV_TAP = {(0, 0): 1, (0, -1): 3, (1, 0): 2, (1, 1): 0}       # (output phase bit, input shift) -> kernel index


def to_blocked2d(x):
    """fp32 NCHW -> bf16 (B, C/8, H, W, 8)."""
    B, C, H, W = x.shape
    return to_blocked_bf16(x.unsqueeze(2)).view(B, C // 8, H, W, 8)


def from_blocked2d(xb):
    B, C8, H, W, _ = xb.shape
    return from_blocked_bf16(xb.view(B, C8, 1, H, W, 8)).squeeze(2)


def ntile2d(mode, cin, cout):
    return _lib.load().ss_conv2d_tc_ntile(int(mode), int(cin), int(cout))


def pack_weight2d(w, mode):
    """Conv2d weight (Cout,Cin,k,k) for CONV3 / CONV1, ConvTranspose2d weight (Cin,Cout,4,4) for DECONV4 -> the bf16 layout
    ss_conv2d_tc documents (include/semstereo_b200.h)."""
    w = w.detach().float()
    if mode == DECONV4:
        cin, cout = w.shape[:2]
        n = ntile2d(mode, cin, cout)
        if n == 0 or tuple(w.shape[2:]) != (4, 4):
            raise NotImplementedError(f"conv2d_tc: no tensor-core configuration for ConvTranspose2d {tuple(w.shape)}")
        nt, ncb = -(-cout // n), cin // 64
        wp = w.new_zeros((cin, nt * n, 4, 4))
        wp[:, :cout] = w
        wp = wp.reshape(ncb, 8, 8, nt, n, 4, 4)                                    # (cb, chunk, c, nt, n, kh, kw)
        slabs = []
        for (sh, sw), phases in _DECONV_SLABS:
            rows = [wp[..., _DECONV_TAP[(ph >> 1, sh)], _DECONV_TAP[(ph & 1, sw)]] for ph in phases]      # each (cb,chunk,c,nt,n)
            r = torch.stack(rows, 4)                                               # (cb, chunk, c, nt, phase, n)
            slabs.append(r.permute(3, 0, 1, 4, 5, 2).reshape(nt, ncb, 8 * len(phases) * n * 8))   # (nt, cb, [chunk][phase*n][c])
        return torch.cat(slabs, 2).contiguous().to(torch.bfloat16)                 # (nt, ncb, 16*n*64)
    cout, cin = w.shape[:2]
    taps = w.shape[2] * w.shape[3]
    n = ntile2d(mode, cin, cout)
    if n == 0 or taps != (9 if mode == CONV3 else 1):
        raise NotImplementedError(f"conv2d_tc: no tensor-core configuration for Conv2d {tuple(w.shape)} in mode {mode}")
    nt, ncb = -(-cout // n), cin // 64
    wp = w.new_zeros((nt * n, cin, taps))
    wp[:cout] = w.reshape(cout, cin, taps)
    t = wp.reshape(nt, n, ncb, 8, 8, taps).permute(0, 2, 5, 3, 1, 4)                 # (nt, cb, tap, chunk, n, c)
    return t.contiguous().to(torch.bfloat16)


def pack_weight2d_split(w, mode, groups=None):
    """bf16x3 split of a 2-D conv as ONE GEMM over K-concat operands: every input tensor of the (virtual) channel concat is given
    in the [hi | lo | hi] form of to_blocked_tri, and the weight columns of each input group are laid out [w_hi | w_hi | w_lo]
    to match.  groups: channel counts of the concatenated inputs (default: one input).  w as for pack_weight2d."""
    w = w.detach().float()
    cdim = 0 if mode == DECONV4 else 1
    cin = w.shape[cdim]
    groups = [cin] if groups is None else list(groups)
    if sum(groups) != cin:
        raise ValueError("pack_weight2d_split: groups must sum to the input channel count")
    hi, lo = split_f32(w)
    parts, o = [], 0
    for gch in groups:
        sl = [slice(None)] * w.dim()
        sl[cdim] = slice(o, o + gch)
        parts += [hi[tuple(sl)], hi[tuple(sl)], lo[tuple(sl)]]
        o += gch
    return pack_weight2d(torch.cat(parts, cdim), mode)


NONE, RELU, SILU = 0, 1, 2          # act of conv2d_tc (1x1 mode) / dwconv3x3


def conv2d_tc(mode, x0, w_packed, cout, scale=None, shift=None, relu=False, out_f32=False, x1=None, act=None, residual=None):
    """x0 (and optionally x1, concatenated after it along channels) blocked bf16 (B,C/8,H,W,8).  Returns blocked bf16
    (B,Cout/8,OH,OW,8) or fp32 NCHW; OH,OW = H,W (conv) or 2H,2W (DECONV4).  1x1 mode only: act (NONE / RELU / SILU, overrides
    `relu`) and residual (bf16 blocked, shaped like the output, added after the activation)."""
    dev = _require_bf16(x0, 5)
    B, C80, H, W, _ = x0.shape
    c0, c1 = C80 * 8, 0
    if x1 is not None:
        _require_bf16(x1, 5)
        if x1.shape[0] != B or tuple(x1.shape[2:4]) != (H, W):
            raise ValueError("conv2d_tc: the two inputs of the channel concat must agree in batch and spatial size")
        c1 = x1.shape[1] * 8
    n = ntile2d(mode, c0 + c1, cout)
    if n == 0 or c0 % 64 or c1 % 64:
        raise NotImplementedError(f"conv2d_tc: mode {mode} with Cin=({c0},{c1}), Cout={cout} has no tensor-core configuration")
    nt, ncb = -(-cout // n), (c0 + c1) // 64
    want = (nt, ncb, 16 * n * 64) if mode == DECONV4 else (nt, ncb, 9 if mode == CONV3 else 1, 8, n, 8)
    if w_packed.dtype != torch.bfloat16 or tuple(w_packed.shape) != want or not w_packed.is_contiguous():
        raise ValueError("conv2d_tc: weight must come from pack_weight2d(w, mode) for this layer")
    for t in (scale, shift):
        if t is not None:
            _require_cuda(t)
    OH, OW = (2 * H, 2 * W) if mode == DECONV4 else (H, W)
    if out_f32:
        out = torch.empty((B, cout, OH, OW), device=dev, dtype=torch.float32)
    else:
        out = torch.empty((B, cout // 8, OH, OW, 8), device=dev, dtype=torch.bfloat16)
    if residual is not None:
        _require_bf16(residual, 5)
        if out_f32 or tuple(residual.shape) != tuple(out.shape):
            raise ValueError("conv2d_tc: the residual must be bf16 blocked with the shape of the (bf16) output")
    _call("ss_conv2d_tc_ex", dev, int(mode), _ptr(x0), c0, _ptr(x1), c1, _ptr(w_packed), _ptr(scale), _ptr(shift), _ptr(residual), _ptr(out),
          int(out_f32), B, cout, H, W, int(relu) if act is None else int(act))
    return out


# ---------------------------------------------------------------------------------------------------------------------
# backbone kernels (csrc/backbone.cu)
# ---------------------------------------------------------------------------------------------------------------------
def stem_conv(image, weight, scale, shift, cout_padded=64):
    """Conv2d(3,32,3,s2,p1) + folded BN + SiLU: fp32 (B,3,H,W) -> bf16 blocked (B,cout_padded/8,H/2,W/2,8), channels >= 32 zero."""
    dev = _require_cuda(image, weight, scale, shift)
    B, C, H, W = image.shape
    if C != 3 or tuple(weight.shape) != (32, 3, 3, 3):
        raise ValueError("stem_conv: (B,3,H,W) image and (32,3,3,3) weight expected")
    out = torch.empty((B, cout_padded // 8, H // 2, W // 2, 8), device=dev, dtype=torch.bfloat16)
    _call("ss_stem_conv3x3_s2", dev, _ptr(image), _ptr(weight), _ptr(scale), _ptr(shift), _ptr(out), B, H, W, int(cout_padded))
    return out


def dwconv3x3(xb, weight, scale, shift, stride=1, act=SILU):
    """Depthwise Conv2d 3x3 p1 (+ folded BN + act) on blocked bf16; weight fp32 (C,9)."""
    dev = _require_bf16(xb, 5)
    _require_cuda(weight, scale, shift)
    B, C8, H, W, _ = xb.shape
    if weight.numel() != C8 * 8 * 9:
        raise ValueError("dwconv3x3: weight (C,9) expected")
    out = torch.empty((B, C8, H // stride, W // stride, 8), device=dev, dtype=torch.bfloat16)
    _call("ss_dwconv3x3_blocked", dev, _ptr(xb), _ptr(weight), _ptr(scale), _ptr(shift), _ptr(out), B, C8 * 8, H, W, int(stride), int(act))
    return out


def groupnorm1(xb, gamma, beta, eps=1e-5):
    """nn.GroupNorm(1, C) on blocked bf16 (statistics in fp32)."""
    dev = _require_bf16(xb, 5)
    _require_cuda(gamma, beta)
    B, C8, H, W, _ = xb.shape
    ws = torch.empty(_lib.load().ss_groupnorm1_workspace_floats(B), device=dev, dtype=torch.float32)
    out = torch.empty_like(xb)
    _call("ss_groupnorm1_blocked", dev, _ptr(xb), _ptr(gamma), _ptr(beta), _ptr(out), _ptr(ws), B, C8 * 8, H, W, ctypes.c_float(eps))
    return out


def linear_attention(qkv_b, d):
    """MobileViTv2 separable self-attention core: qkv blocked (B, 2d/8 + 1, H, W, 8) [key | value | query in lane 0 of the last
    chunk] -> relu(value) * context, blocked (B, d/8, H, W, 8)."""
    dev = _require_bf16(qkv_b, 5)
    B, CH, H, W, _ = qkv_b.shape
    if CH != 2 * (d // 8) + 1:
        raise ValueError("linear_attention: qkv must have 2d/8 + 1 channel chunks")
    ws = torch.empty(_lib.load().ss_linear_attention_workspace_floats(B, d), device=dev, dtype=torch.float32)
    out = torch.empty((B, d // 8, H, W, 8), device=dev, dtype=torch.bfloat16)
    _call("ss_linear_attention_blocked", dev, _ptr(qkv_b), _ptr(out), _ptr(ws), B, int(d), H, W)
    return out


def bilinear_up2(x):
    """F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=False) for fp32 NCHW."""
    dev = _require_cuda(x)
    B, C, h, w = x.shape
    out = torch.empty((B, C, 2 * h, 2 * w), device=dev, dtype=torch.float32)
    _call("ss_bilinear_up2", dev, _ptr(x), _ptr(out), B * C, h, w)
    return out


def pointwise_blocked_small(xb, weight, bias=None):
    """Conv2d 1x1 (Cout <= 8) from blocked bf16 (B,C/8,H,W,8) to fp32 NCHW -- segmenthead.conv2."""
    dev = _require_bf16(xb, 5)
    _require_cuda(weight, bias)
    B, C8, H, W, _ = xb.shape
    cout = weight.shape[0]
    if weight.numel() != cout * C8 * 8 or (bias is not None and bias.numel() != cout):
        raise ValueError("pointwise_blocked_small: weight (Cout,C[,1,1]) and bias (Cout,) expected")
    out = torch.empty((B, cout, H, W), device=dev, dtype=torch.float32)
    _call("ss_pointwise_blocked_small", dev, _ptr(xb), _ptr(weight), _ptr(bias), _ptr(out), B, C8 * 8, cout, H, W)
    return out


def concat_stem_fused(cf_l_b, cf_r_b, disp_topk, att_topk, w_s1f, dmin, scale=None, shift=None, gate_blocked=None, relu=True,
                      out_mode=BLOCKED):
    """concat_volume_generator * att_topk -> concat_stem (+ gate) with the volume produced inside the kernel (SemStereo.py:316-320).
    cf_*_b: bf16 blocked (B,4,H,W,8) [or (B,4,1,H,W,8)]; disp_topk / att_topk fp32 (B,K,H,W) with integer samples in
    [dmin, dmin+31]; w_s1f = pack_weight(concat_stem weight, S1F).  Returns (B,4,K,H,W,8) blocked / phase-split / fp32 (B,32,K,H,W)."""
    cf_l_b = cf_l_b.view(cf_l_b.shape[0], 4, *cf_l_b.shape[-3:])
    cf_r_b = cf_r_b.view(cf_r_b.shape[0], 4, *cf_r_b.shape[-3:])
    dev = _require_bf16(cf_l_b, 5)
    _require_bf16(cf_r_b, 5)
    _require_cuda(disp_topk, att_topk, scale, shift, gate_blocked)
    B, _, H, W, _ = cf_l_b.shape
    K = disp_topk.shape[1]
    if att_topk.dim() == 5 and att_topk.shape[1] == 1:          # the model's att_topk keeps the cost volume's channel dim (B,1,K,H,W)
        att_topk = att_topk.squeeze(1)
    if cf_r_b.shape != cf_l_b.shape or tuple(disp_topk.shape) != (B, K, H, W) or att_topk.shape != disp_topk.shape:
        raise ValueError("concat_stem_fused: cf (B,4,H,W,8) and samples (B,K,H,W) expected")
    if w_s1f.dtype != torch.bfloat16 or tuple(w_s1f.shape) != (9, 8, 96, 8) or not w_s1f.is_contiguous():
        raise ValueError("concat_stem_fused: weight must come from pack_weight(w, S1F) of a Conv3d(64, 32, 3)")
    if gate_blocked is not None and tuple(gate_blocked.shape) != (B, 4, H, W, 8):
        raise ValueError("concat_stem_fused: gate must be fp32 (B,4,H,W,8) from gate_sigmoid_blocked")
    if out_mode == F32:
        out = torch.empty((B, 32, K, H, W), device=dev, dtype=torch.float32)
    elif out_mode == S2D:
        out = torch.empty((B, 8, 4, K // 2, H // 2, W // 2, 8), device=dev, dtype=torch.bfloat16)
    else:
        out = torch.empty((B, 4, K, H, W, 8), device=dev, dtype=torch.bfloat16)
    _call("ss_concat_stem_fused", dev, _ptr(cf_l_b), _ptr(cf_r_b), _ptr(disp_topk), _ptr(att_topk), _ptr(w_s1f), _ptr(scale), _ptr(shift),
          _ptr(gate_blocked), _ptr(out), int(out_mode), B, K, H, W, int(dmin), int(relu))
    return out
