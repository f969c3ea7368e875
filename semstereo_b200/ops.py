"""Thin, validating Python wrappers over the C-ABI kernels (include/semstereo_b200.h).

Each wrapper (1) checks device / dtype / contiguity / shape, (2) allocates the outputs with torch on the
input's device, (3) passes torch's current CUDA stream, (4) raises on a non-zero return code.  Tensors are
fp32 CUDA tensors; nothing here computes on the CPU and nothing falls back to torch operators.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

SIGNED, NORM = 1, 2


def _require_cuda(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("semstereo_b200 kernels need CUDA tensors: there is no CPU fallback on this path")
        if t.dtype != torch.float32:
            raise TypeError(f"expected float32, got {t.dtype}")
        if not t.is_contiguous():
            raise ValueError("expected a contiguous tensor")
    dev = ts[0].device
    for t in ts:
        if t is not None and t.device != dev:
            raise ValueError("all tensors must live on the same device")
    return dev


def _ptr(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


class LaunchRecorder:
    """Counts kernel launches made through the C-ABI and (optionally) brackets each with CUDA events on the launching
    stream, so a benchmark can time every kernel live inside its timed region.  Install with `record_launches`."""

    def __init__(self, timing: bool = True):
        self.timing, self.count, self.records = timing, 0, []   # records: (label, entry point, start_event, end_event)

    def durations_ms(self):
        """label -> list of per-launch durations (call after a device synchronize)."""
        out = {}
        for label, name, a, b in self.records:
            out.setdefault(label or name, []).append(a.elapsed_time(b))
        return out


_recorder = None
_label = None


def record_launches(rec):
    """Installs (or with None removes) the process-wide launch recorder; returns the previous one."""
    global _recorder
    prev, _recorder = _recorder, rec
    return prev


class label:
    """Context manager naming the launches made inside it (e.g. the conv layer a generic kernel is running)."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        global _label
        self.prev, _label = _label, self.name

    def __exit__(self, *a):
        global _label
        _label = self.prev


def _call(name, dev, *args):
    lib = _lib.load()
    rec = _recorder
    with torch.cuda.device(dev):
        if rec is not None:
            rec.count += 1
            if rec.timing:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
        rc = getattr(lib, name)(*args, _stream(dev))
        if rec is not None and rec.timing:
            b.record()
            rec.records.append((_label, name, a, b))
    _lib.check(rc, name)


# ---- K1 / K2 -----------------------------------------------------------------------------------------
def gwc_volume(left, right, maxdisp, num_groups, signed=True, norm=False):
    dev = _require_cuda(left, right)
    if left.dim() != 4 or left.shape != right.shape:
        raise ValueError("gwc_volume: left/right must be (B,C,H,W) of equal shape")
    B, C, H, W = left.shape
    D = 2 * maxdisp if signed else maxdisp
    out = torch.empty((B, num_groups, D, H, W), device=dev, dtype=torch.float32)
    flags = (SIGNED if signed else 0) | (NORM if norm else 0)
    _call("ss_gwc_volume", dev, _ptr(left), _ptr(right), _ptr(out), B, C, H, W, int(maxdisp), int(num_groups), flags)
    return out


def concat_volume(left, right, maxdisp, signed=True):
    dev = _require_cuda(left, right)
    if left.dim() != 4 or left.shape != right.shape:
        raise ValueError("concat_volume: left/right must be (B,C,H,W) of equal shape")
    B, C, H, W = left.shape
    D = 2 * maxdisp if signed else maxdisp
    out = torch.empty((B, 2 * C, D, H, W), device=dev, dtype=torch.float32)
    _call("ss_concat_volume", dev, _ptr(left), _ptr(right), _ptr(out), B, C, H, W, int(maxdisp), SIGNED if signed else 0)
    return out


# ---- K3 ----------------------------------------------------------------------------------------------
def patch_gate(volume, patch_w=None, gate_logits=None):
    dev = _require_cuda(volume, patch_w, gate_logits)
    B, G, D, H, W = volume.shape
    if patch_w is not None and patch_w.numel() != G * 9:
        raise ValueError("patch_gate: patch weight must have G*9 elements")
    if gate_logits is not None and tuple(gate_logits.shape) != (B, G, H, W):
        raise ValueError("patch_gate: gate logits must be (B,G,H,W)")
    out = torch.empty_like(volume)
    _call("ss_patch_gate", dev, _ptr(volume), _ptr(patch_w), _ptr(gate_logits), _ptr(out), B, G, D, H, W)
    return out


def pointwise_conv2d(x, weight, scale=None, shift=None, relu=False):
    dev = _require_cuda(x, weight, scale, shift)
    B, Cin, H, W = x.shape
    Cout = weight.shape[0]
    if weight.numel() != Cout * Cin:
        raise ValueError("pointwise_conv2d: weight must be (Cout,Cin[,1,1])")
    out = torch.empty((B, Cout, H, W), device=dev, dtype=torch.float32)
    _call("ss_pointwise_conv2d", dev, _ptr(x), _ptr(weight), _ptr(scale), _ptr(shift), _ptr(out), B, Cin, Cout, H * W, int(relu))
    return out


# ---- K4 ----------------------------------------------------------------------------------------------
def pack_conv3d_weight(w, transposed=False):
    """(Cout,Cin,k,k,k) [or ConvTranspose3d (Cin,Cout,k,k,k)] -> [k^3][Cin][Cout] contiguous."""
    if transposed:
        return w.permute(2, 3, 4, 0, 1).reshape(-1, w.shape[0], w.shape[1]).contiguous()
    return w.permute(2, 3, 4, 1, 0).reshape(-1, w.shape[1], w.shape[0]).contiguous()


def conv3d_f32(x, w_packed, scale=None, shift=None, residual=None, gate_logits=None, k=3, stride=1, transposed=False, relu=False):
    dev = _require_cuda(x, w_packed, scale, shift, residual, gate_logits)
    B, Cin, Di, Hi, Wi = x.shape
    taps, cin_w, Cout = w_packed.shape
    if taps != k ** 3 or cin_w != Cin:
        raise ValueError("conv3d_f32: packed weight must be [k^3][Cin][Cout]")
    if transposed:
        Do, Ho, Wo = 2 * Di, 2 * Hi, 2 * Wi
    else:
        pad = k // 2
        Do, Ho, Wo = [(n + 2 * pad - k) // stride + 1 for n in (Di, Hi, Wi)]
    out = torch.empty((B, Cout, Do, Ho, Wo), device=dev, dtype=torch.float32)
    if residual is not None and residual.shape != out.shape:
        raise ValueError("conv3d_f32: residual must match the output shape")
    if gate_logits is not None and tuple(gate_logits.shape) != (B, Cout, Ho, Wo):
        raise ValueError("conv3d_f32: gate logits must be (B,Cout,Ho,Wo)")
    _call("ss_conv3d_f32", dev, _ptr(x), _ptr(w_packed), _ptr(scale), _ptr(shift), _ptr(residual), _ptr(gate_logits), _ptr(out),
          B, Cin, Cout, Di, Hi, Wi, int(k), int(stride), int(transposed), int(relu))
    return out


def conv3d_cout1_f32(x, weight):
    dev = _require_cuda(x, weight)
    B, Cin, D, H, W = x.shape
    if tuple(weight.shape) != (1, Cin, 3, 3, 3):
        raise ValueError("conv3d_cout1_f32: weight must be (1,Cin,3,3,3)")
    out = torch.empty((B, 1, D, H, W), device=dev, dtype=torch.float32)
    _call("ss_conv3d_cout1_f32", dev, _ptr(x), _ptr(weight), _ptr(out), B, Cin, D, H, W)
    return out


# ---- K5 ----------------------------------------------------------------------------------------------
def window_needs_mask(shape, block):
    """True when BOTH H and W of a (B,C,D,H,W[,8]) volume need padding to the window: the reference's masked branch."""
    return bool((-shape[3]) % block[1]) and bool((-shape[4]) % block[2])


def window_pad(t, block, both=False):
    """Zero-pads axes 3 (H) and 4 (W) of a (B,C,D,H,W) / blocked (B,C/8,D,H,W,8) tensor to multiples of the window, as
    attention_block.forward does before its qkv Linear (submodule_other.py:809-812).  Returns (tensor, H0, W0).
    Padding on ONE of the two axes: the reference's mask `mask[:, -pad_b:, :] = 1; mask[:, :, -pad_r:] = 1` is all ones when exactly
    one pad is 0 (`-0:` selects everything), so nothing is masked and the padded tokens (which carry the qkv bias) simply take part in
    their window's softmax: the unmodified kernels on the padded volume, cropped afterwards, ARE that computation.
    Padding on BOTH axes needs the -1000 score mask between padded and real tokens: only window_attention3d (the masked fp32 core,
    ss_window_attention_core_f32_masked) computes that, so callers of the unmasked kernels get a refusal unless they pass both=True."""
    D, H, W = t.shape[2], t.shape[3], t.shape[4]
    if D % block[0]:
        raise NotImplementedError(f"window attention: depth {D} is not a multiple of the window depth {block[0]} (the reference does not pad D)")
    pb, pr = (-H) % block[1], (-W) % block[2]
    if pb and pr and not both:
        raise NotImplementedError(f"window attention: H = {H} and W = {W} both need padding to the window {tuple(block)[1:]}: the "
                                  "reference's masked branch (submodule_other.py:822-829) runs through ops.window_attention3d only")
    if not (pb or pr):
        return t, H, W
    shape = list(t.shape)
    shape[3], shape[4] = H + pb, W + pr
    out = t.new_zeros(shape)
    out[:, :, :, :H, :W] = t
    return out, H, W


def window_crop(t, H0, W0):
    """Inverse of window_pad on the output (submodule_other.py:835-836)."""
    return t if (t.shape[3], t.shape[4]) == (H0, W0) else t[:, :, :, :H0, :W0].contiguous()


def window_attention_core_masked(qkv, block, H0, W0, num_heads=16):
    """The softmax core of the masked branch: qkv (B,3C,D,H,W) fp32 of the PADDED volume -> (B,C,D,H,W) fp32 (submodule_other.py:822-831)."""
    dev = _require_cuda(qkv)
    B, C3, D, H, W = qkv.shape
    out = torch.empty((B, C3 // 3, D, H, W), device=dev, dtype=torch.float32)
    _call("ss_window_attention_core_f32_masked", dev, _ptr(qkv), _ptr(out), B, C3 // 3, D, H, W, int(block[0]), int(block[1]), int(block[2]),
          int(num_heads), int(H0), int(W0))
    return out


def window_attention3d(x, wqkv_t, bqkv, wo_t, bo, block, num_heads=16):
    dev = _require_cuda(x, wqkv_t, bqkv, wo_t, bo)
    if window_needs_mask(x.shape, block):
        # both axes padded (submodule_other.py:809-836): pad -> qkv Linear (padded tokens = bias) -> masked core -> crop -> final conv
        xp, H0, W0 = window_pad(x, block, both=True)
        B, C, D, H, W = xp.shape
        qkv = pointwise_conv2d(xp.view(B, C, D * H, W), wqkv_t.t().contiguous(), None, bqkv).view(B, 3 * C, D, H, W)
        att = window_crop(window_attention_core_masked(qkv, block, H0, W0, num_heads), H0, W0)
        return pointwise_conv2d(att.view(B, C, D * H0, W0), wo_t.t().contiguous(), None, bo).view(B, C, D, H0, W0)
    x, H0, W0 = window_pad(x, block)
    B, C, D, H, W = x.shape
    out = torch.empty_like(x)
    _call("ss_window_attention3d", dev, _ptr(x), _ptr(wqkv_t), _ptr(bqkv), _ptr(wo_t), _ptr(bo), _ptr(out), B, C, D, H, W,
          int(block[0]), int(block[1]), int(block[2]), int(num_heads))
    return window_crop(out, H0, W0)            # the final 1x1x1 conv is position-wise: cropping after it equals cropping before it


# ---- K6 / K7 / K8 ------------------------------------------------------------------------------------
def att_stats(cost_att, beta, gamma, dmin):
    dev = _require_cuda(cost_att, beta, gamma)
    B, one, D8, H8, W8 = cost_att.shape
    if one != 1:
        raise ValueError("att_stats: cost_att must be (B,1,D,H,W)")
    att_up = torch.empty((B, 1, 2 * D8, 2 * H8, 2 * W8), device=dev, dtype=torch.float32)
    mu = torch.empty((B, 2 * H8, 2 * W8), device=dev, dtype=torch.float32)
    gate = torch.empty((B, 1, 2 * H8, 2 * W8), device=dev, dtype=torch.float32)
    _call("ss_att_stats", dev, _ptr(cost_att), _ptr(beta), _ptr(gamma), _ptr(att_up), _ptr(mu), _ptr(gate), B, D8, H8, W8, float(dmin))
    return att_up, mu, gate


def sample_strength(feat_l, feat_r, mu, gate):
    dev = _require_cuda(feat_l, feat_r, mu, gate)
    B, C, H, W = feat_l.shape
    if feat_r.shape != feat_l.shape or mu.numel() != B * H * W or gate.numel() != B * H * W:
        raise ValueError("sample_strength: shape mismatch")
    out = torch.empty((B, 5, H, W), device=dev, dtype=torch.float32)
    _call("ss_sample_strength", dev, _ptr(feat_l), _ptr(feat_r), _ptr(mu), _ptr(gate), _ptr(out), B, C, H, W)
    return out


def topk_select(att_up, strength, k, disp_offset, want_indices=True, want_prob=False):
    dev = _require_cuda(att_up, strength)
    B, one, nb, H, W = att_up.shape
    if one != 1 or tuple(strength.shape) != (B, 5, H, W):
        raise ValueError("topk_select: att_up must be (B,1,D,H,W) and strength (B,5,H,W)")
    ind = torch.empty((B, 1, k, H, W), device=dev, dtype=torch.int64) if want_indices else None
    att_topk = torch.empty((B, 1, k, H, W), device=dev, dtype=torch.float32)
    disp_topk = torch.empty((B, k, H, W), device=dev, dtype=torch.float32)
    pred = torch.empty((B, H, W), device=dev, dtype=torch.float32)
    prob = torch.empty((B, 1, nb, H, W), device=dev, dtype=torch.float32) if want_prob else None
    _call("ss_topk_select", dev, _ptr(att_up), _ptr(strength), ctypes.c_void_p(0 if ind is None else ind.data_ptr()),
          _ptr(att_topk), _ptr(disp_topk), _ptr(pred), _ptr(prob), B, nb, int(k), H, W, float(disp_offset))
    return ind, att_topk, disp_topk, pred, prob


# ---- K9 / K10 / K11 / K12 ----------------------------------------------------------------------------
def sparse_concat_volume(cf_l, cf_r, disp_topk, att_topk=None):
    dev = _require_cuda(cf_l, cf_r, disp_topk, att_topk)
    B, C, H, W = cf_l.shape
    K = disp_topk.shape[1]
    if cf_r.shape != cf_l.shape or tuple(disp_topk.shape) != (B, K, H, W):
        raise ValueError("sparse_concat_volume: shape mismatch")
    if att_topk is not None and att_topk.numel() != B * K * H * W:
        raise ValueError("sparse_concat_volume: att_topk must have B*K*H*W elements")
    out = torch.empty((B, 2 * C, K, H, W), device=dev, dtype=torch.float32)
    _call("ss_sparse_concat_volume", dev, _ptr(cf_l), _ptr(cf_r), _ptr(disp_topk), _ptr(att_topk), _ptr(out), B, C, K, H, W)
    return out


def regression_topk(cost, disp_samples, k):
    dev = _require_cuda(cost, disp_samples)
    B, D, H, W = cost.shape
    if disp_samples.shape != cost.shape:
        raise ValueError("regression_topk: cost and disparity_samples must have the same shape")
    out = torch.empty((B, 1, H, W), device=dev, dtype=torch.float32)
    _call("ss_regression_topk", dev, _ptr(cost), _ptr(disp_samples), _ptr(out), B, D, int(k), H, W)
    return out


def ssr_param_count(num_classes=6):
    return _lib.load().ss_ssr_param_count(int(num_classes))


def ssr_upsample(depth_low, spx, label, packed_host):
    """packed_host: CPU float32 tensor of ssr_param_count() folded parameters (see hotpath.pack_ssr)."""
    dev = _require_cuda(depth_low, spx, label)
    B, one, h, w = depth_low.shape
    nc = spx.shape[1]
    if one != 1 or tuple(spx.shape) != (B, nc, 4 * h, 4 * w) or label.shape != spx.shape:
        raise ValueError("ssr_upsample: depth_low (B,1,h,w), weights/pred_label (B,nc,4h,4w)")
    if packed_host.is_cuda or packed_host.dtype != torch.float32 or packed_host.numel() != ssr_param_count(nc):
        raise ValueError("ssr_upsample: packed parameters must be a CPU float32 tensor of ss_ssr_param_count() elements")
    out = torch.empty((B, 4 * h, 4 * w), device=dev, dtype=torch.float32)
    _call("ss_ssr_upsample", dev, _ptr(depth_low), _ptr(spx), _ptr(label), _ptr(out),
          ctypes.c_void_p(packed_host.contiguous().data_ptr()), B, h, w, nc)
    return out


def ssr_upsample2(depth_low_a, depth_low_b, spx, label, packed_host):
    """Both SSR_upsample calls of the model (SemStereo.py:312, :324) in one pass; returns (out_a, out_b)."""
    dev = _require_cuda(depth_low_a, depth_low_b, spx, label)
    B, one, h, w = depth_low_a.shape
    nc = spx.shape[1]
    if one != 1 or depth_low_b.shape != depth_low_a.shape or tuple(spx.shape) != (B, nc, 4 * h, 4 * w) or label.shape != spx.shape:
        raise ValueError("ssr_upsample2: depth_low_a/b (B,1,h,w), weights/pred_label (B,nc,4h,4w)")
    if packed_host.is_cuda or packed_host.dtype != torch.float32 or packed_host.numel() != ssr_param_count(nc):
        raise ValueError("ssr_upsample2: packed parameters must be a CPU float32 tensor of ss_ssr_param_count() elements")
    out_a = torch.empty((B, 4 * h, 4 * w), device=dev, dtype=torch.float32)
    out_b = torch.empty_like(out_a)
    _call("ss_ssr_upsample2", dev, _ptr(depth_low_a), _ptr(depth_low_b), _ptr(spx), _ptr(label), _ptr(out_a), _ptr(out_b),
          ctypes.c_void_p(packed_host.contiguous().data_ptr()), B, h, w, nc)
    return out_a, out_b


def context_upsample(depth_low, up_weights):
    dev = _require_cuda(depth_low, up_weights)
    B, one, h, w = depth_low.shape
    if one != 1 or tuple(up_weights.shape) != (B, 9, 4 * h, 4 * w):
        raise ValueError("context_upsample: depth_low (B,1,h,w), up_weights (B,9,4h,4w)")
    out = torch.empty((B, 4 * h, 4 * w), device=dev, dtype=torch.float32)
    _call("ss_context_upsample", dev, _ptr(depth_low), _ptr(up_weights), _ptr(out), B, h, w)
    return out


def disparity_regression(prob, dmin):
    dev = _require_cuda(prob)
    if prob.dim() != 4:
        raise AssertionError("disparity_regression expects (B,D,H,W)")
    B, D, H, W = prob.shape
    out = torch.empty((B, H, W), device=dev, dtype=torch.float32)
    _call("ss_disparity_regression", dev, _ptr(prob), _ptr(out), B, D, H, W, float(dmin))
    return out


def disparity_variance(prob, disparity, dmin):
    dev = _require_cuda(prob, disparity)
    if prob.dim() != 4:
        raise AssertionError("disparity_variance expects (B,D,H,W)")
    B, D, H, W = prob.shape
    if disparity.numel() != B * H * W:
        raise ValueError("disparity_variance: disparity must be (B,1,H,W)")
    out = torch.empty((B, 1, H, W), device=dev, dtype=torch.float32)
    _call("ss_disparity_variance", dev, _ptr(prob), _ptr(disparity), _ptr(out), B, D, H, W, float(dmin))
    return out


def propagation(x):
    """(B,1,H,W) -> (B,5,H,W)  or  (B,1,D,H,W) -> (B,5,D,H,W)."""
    dev = _require_cuda(x)
    if x.shape[1] != 1:
        raise ValueError("propagation expects a single channel")
    if x.dim() == 4:
        B, _, H, W = x.shape
        out = torch.empty((B, 5, H, W), device=dev, dtype=torch.float32)
        D = 1
    else:
        B, _, D, H, W = x.shape
        out = torch.empty((B, 5, D, H, W), device=dev, dtype=torch.float32)
    _call("ss_propagation", dev, _ptr(x), _ptr(out), B, D, H, W)
    return out


def spatial_transformer_grid(x, y, disp_samples, want_x_rep=True):
    dev = _require_cuda(x, y, disp_samples)
    B, C, H, W = y.shape
    K = disp_samples.shape[1]
    if tuple(disp_samples.shape) != (B, K, H, W) or x.shape != y.shape:
        raise ValueError("spatial_transformer_grid: shape mismatch")
    yw = torch.empty((B, C, K, H, W), device=dev, dtype=torch.float32)
    xr = torch.empty((B, C, K, H, W), device=dev, dtype=torch.float32) if want_x_rep else None
    _call("ss_spatial_transformer_grid", dev, _ptr(x), _ptr(y), _ptr(disp_samples), _ptr(yw), _ptr(xr), B, C, K, H, W)
    return yw, xr



# ---- backward (VJP) of the volume / regression operators (csrc/backward.cu) ----------------------------------------
def gwc_volume_backward(left, right, grad_volume, maxdisp, num_groups, signed=True, norm=False):
    dev = _require_cuda(left, right, grad_volume)
    B, C, H, W = left.shape
    D = 2 * maxdisp if signed else maxdisp
    if right.shape != left.shape or tuple(grad_volume.shape) != (B, num_groups, D, H, W):
        raise ValueError("gwc_volume_backward: left/right (B,C,H,W) and grad_volume (B,G,D,H,W) expected")
    gl, gr = torch.empty_like(left), torch.empty_like(right)
    flags = (SIGNED if signed else 0) | (NORM if norm else 0)
    _call("ss_gwc_volume_backward", dev, _ptr(left), _ptr(right), _ptr(grad_volume), _ptr(gl), _ptr(gr), B, C, H, W, int(maxdisp),
          int(num_groups), flags)
    return gl, gr


def concat_volume_backward(grad_volume, maxdisp, signed=True):
    dev = _require_cuda(grad_volume)
    B, C2, D, H, W = grad_volume.shape
    if C2 % 2 or D != (2 * maxdisp if signed else maxdisp):
        raise ValueError("concat_volume_backward: grad_volume must be (B,2C,D,H,W) with D matching maxdisp")
    gl = torch.empty((B, C2 // 2, H, W), device=dev, dtype=torch.float32)
    gr = torch.empty_like(gl)
    _call("ss_concat_volume_backward", dev, _ptr(grad_volume), _ptr(gl), _ptr(gr), B, C2 // 2, H, W, int(maxdisp), SIGNED if signed else 0)
    return gl, gr


def disparity_regression_backward(grad_out, D, dmin):
    dev = _require_cuda(grad_out)
    B, H, W = grad_out.shape
    gp = torch.empty((B, D, H, W), device=dev, dtype=torch.float32)
    _call("ss_disparity_regression_backward", dev, _ptr(grad_out), _ptr(gp), B, int(D), H, W, float(dmin))
    return gp


def regression_topk_backward(cost, disp_samples, grad_pred, k):
    dev = _require_cuda(cost, disp_samples, grad_pred)
    B, D, H, W = cost.shape
    if disp_samples.shape != cost.shape or grad_pred.numel() != B * H * W:
        raise ValueError("regression_topk_backward: cost/samples (B,D,H,W), grad_pred (B,1,H,W) expected")
    gc, gs = torch.empty_like(cost), torch.empty_like(cost)
    _call("ss_regression_topk_backward", dev, _ptr(cost), _ptr(disp_samples), _ptr(grad_pred), _ptr(gc), _ptr(gs), B, D, int(k), H, W)
    return gc, gs


def context_upsample_backward(depth_low, up_weights, grad_out):
    dev = _require_cuda(depth_low, up_weights, grad_out)
    B, one, h, w = depth_low.shape
    if one != 1 or tuple(up_weights.shape) != (B, 9, 4 * h, 4 * w) or tuple(grad_out.shape) != (B, 4 * h, 4 * w):
        raise ValueError("context_upsample_backward: depth_low (B,1,h,w), up_weights (B,9,4h,4w), grad_out (B,4h,4w)")
    gd, gw = torch.empty_like(depth_low), torch.empty_like(up_weights)
    _call("ss_context_upsample_backward", dev, _ptr(depth_low), _ptr(up_weights), _ptr(grad_out), _ptr(gd), _ptr(gw), B, h, w)
    return gd, gw


def propagation_backward(grad_out):
    dev = _require_cuda(grad_out)
    if grad_out.shape[1] != 5:
        raise ValueError("propagation_backward expects (B,5,[D,]H,W)")
    if grad_out.dim() == 4:
        B, _, H, W = grad_out.shape
        D, gin = 1, torch.empty((B, 1, H, W), device=dev, dtype=torch.float32)
    else:
        B, _, D, H, W = grad_out.shape
        gin = torch.empty((B, 1, D, H, W), device=dev, dtype=torch.float32)
    _call("ss_propagation_backward", dev, _ptr(grad_out), _ptr(gin), B, D, H, W)
    return gin


def disparity_variance_backward(prob, disparity, grad_out, dmin):
    dev = _require_cuda(prob, disparity, grad_out)
    B, D, H, W = prob.shape
    gp, gm = torch.empty_like(prob), torch.empty((B, 1, H, W), device=dev, dtype=torch.float32)
    _call("ss_disparity_variance_backward", dev, _ptr(prob), _ptr(disparity), _ptr(grad_out), _ptr(gp), _ptr(gm), B, D, H, W, float(dmin))
    return gp, gm


def spatial_transformer_grid_backward(y, disp_samples, grad_y_warped, grad_x_rep=None):
    """Returns (grad_x or None, grad_y, grad_disp)."""
    dev = _require_cuda(y, disp_samples, grad_y_warped, grad_x_rep)
    B, C, H, W = y.shape
    K = disp_samples.shape[1]
    if tuple(grad_y_warped.shape) != (B, C, K, H, W):
        raise ValueError("spatial_transformer_grid_backward: grad_y_warped must be (B,C,K,H,W)")
    gy = torch.zeros_like(y)
    gd = torch.empty_like(disp_samples)
    gx = torch.empty_like(y) if grad_x_rep is not None else None
    _call("ss_spatial_transformer_grid_backward", dev, _ptr(y), _ptr(disp_samples), _ptr(grad_y_warped), _ptr(grad_x_rep), _ptr(gx),
          _ptr(gy), _ptr(gd), B, C, K, H, W)
    return gx, gy, gd
