"""oracle/ops.py — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Closed-form CPU restatements (torch fp32) of the free functions and small modules of the
reference operator surface.  Every function cites the reference lines it restates
(paths relative to /root/reference).  `signed=True` follows `models/submodule.py`
(disparities -M..M-1, depth 2M); `signed=False` follows `models/submodule_.py`
(0..M-1, depth M).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

EPS_NORM = 1e-5  # added to the L2 norm, outside the sqrt (models/submodule.py:218)


def disparities(maxdisp: int, signed: bool):
    """Disparity bin values, in volume order (submodule.py:167 / submodule_.py:161)."""
    return list(range(-maxdisp, maxdisp)) if signed else list(range(0, maxdisp))


def shift_w(t: torch.Tensor, d: int) -> torch.Tensor:
    """out[..., x] = t[..., x - d] where 0 <= x - d < W, else 0."""
    W = t.shape[-1]
    out = torch.zeros_like(t)
    if d == 0:
        out.copy_(t)
    elif 0 < d < W:
        out[..., d:] = t[..., : W - d]
    elif -W < d < 0:
        out[..., : W + d] = t[..., -d:]
    return out


def valid_w(W: int, d: int, dtype=torch.float32) -> torch.Tensor:
    """1.0 where 0 <= x - d < W (the columns the reference's slices write)."""
    x = torch.arange(W)
    return ((x - d >= 0) & (x - d < W)).to(dtype)


# ------------------------------------------------------------------------------------
# K1 / K2 : volume builders
# ------------------------------------------------------------------------------------
def gwc_volume(ref, tgt, maxdisp, num_groups, signed=True, norm=False):
    """build_gwc_volume (submodule.py:198-211, submodule_.py:188-198) and
    build_gwc_volume_norm (submodule.py:224-238, submodule_.py:211-221).

    vol[b,g,k,y,x] = mean_{c in g} L[b,c,y,x] * R[b,c,y,x-d_k]   (0 where x-d_k outside [0,W))
    norm=True: L,R are first divided by (||.||_2 over the channel group + 1e-5)
    (groupwise_correlation_norm, submodule.py:213-221).  Normalising once and then
    shifting is bit-identical to the reference's normalise-per-shift on CPU.
    """
    B, C, H, W = ref.shape
    assert C % num_groups == 0
    cg = C // num_groups
    l = ref.reshape(B, num_groups, cg, H, W)
    r = tgt.reshape(B, num_groups, cg, H, W)
    if norm:
        l = l / (torch.norm(l, 2, 2, True) + EPS_NORM)
        r = r / (torch.norm(r, 2, 2, True) + EPS_NORM)
    ds = disparities(maxdisp, signed)
    vol = ref.new_zeros(B, num_groups, len(ds), H, W)
    for k, d in enumerate(ds):
        vol[:, :, k] = (l * shift_w(r, d)).mean(dim=2)
    return vol


def norm_correlation_volume(ref, tgt, maxdisp, signed=True):
    """build_norm_correlation_volume (submodule.py:244-255): gwc_volume_norm with one group."""
    return gwc_volume(ref, tgt, maxdisp, 1, signed=signed, norm=True)


def concat_volume(ref, tgt, maxdisp, signed=True):
    """build_concat_volume.  Signed (submodule.py:173-187): both halves are zero where
    x-d is out of range.  Unsigned (submodule_.py:166-178): the LEFT half is not masked."""
    B, C, H, W = ref.shape
    ds = disparities(maxdisp, signed)
    vol = ref.new_zeros(B, 2 * C, len(ds), H, W)
    for k, d in enumerate(ds):
        m = valid_w(W, d, ref.dtype)
        vol[:, :C, k] = ref * m if signed else ref
        vol[:, C:, k] = shift_w(tgt, d)
    return vol


# ------------------------------------------------------------------------------------
# K12 : regression / variance
# ------------------------------------------------------------------------------------
def disparity_regression(p, maxdisp, signed=True):
    """sum_k p[:,k] * d_k  (submodule.py:164-170, submodule_.py:159-163)."""
    assert p.dim() == 4
    dv = torch.tensor(disparities(maxdisp, signed), dtype=p.dtype).view(1, -1, 1, 1)
    return (p * dv).sum(1)


def disparity_variance(p, maxdisp, disparity, signed=True):
    """sum_k p[:,k] * (d_k - mu)^2, mu (B,1,H,W)  (submodule.py:257-263)."""
    assert p.dim() == 4
    dv = torch.tensor(disparities(maxdisp, signed), dtype=p.dtype).view(1, -1, 1, 1)
    return (p * (dv - disparity) ** 2).sum(1, keepdim=True)


# ------------------------------------------------------------------------------------
# K7/K8 helpers : one-hot propagation taps
# ------------------------------------------------------------------------------------
# (dy, dx) of the five taps, in output-channel order: one_hot_filter[i,0,ky,kx] with the
# replicate pad of 1 means out_i[y,x] = in[clamp(y+ky-1), clamp(x+kx-1)]
# (submodule.py:297-303, 366-371).
PROP_TAPS = ((-1, -1), (0, 0), (1, 1), (1, -1), (-1, 1))


def _clamped_tap(t, dy, dx):
    H, W = t.shape[-2:]
    ys = (torch.arange(H) + dy).clamp(0, H - 1)
    xs = (torch.arange(W) + dx).clamp(0, W - 1)
    return t.index_select(-2, ys).index_select(-1, xs)


def propagation(x):
    """Propagation.forward (submodule.py:290-307): (B,1,H,W) -> (B,5,H,W)."""
    assert x.shape[1] == 1
    return torch.cat([_clamped_tap(x, dy, dx) for dy, dx in PROP_TAPS], dim=1)


def propagation_prob(v):
    """Propagation_prob.forward (submodule.py:361-377): (B,1,D,H,W) -> (B,5,D,H,W);
    clamp-to-edge in H and W only, never in D."""
    assert v.shape[1] == 1
    return torch.cat([_clamped_tap(v, dy, dx) for dy, dx in PROP_TAPS], dim=1)


# ------------------------------------------------------------------------------------
# a11 : SpatialTransformer_grid
# ------------------------------------------------------------------------------------
def warp_coords(H, W, disp, dtype=torch.float32):
    """Pixel-space sampling coordinates exactly as the reference's fp32 round trip produces
    them: normalise (submodule.py:279-280) then grid_sample's align_corners=True
    un-normalise ((g+1)/2*(size-1))."""
    B, K = disp.shape[:2]
    yy = torch.arange(H, dtype=dtype).view(1, 1, H, 1).expand(B, K, H, W)
    xx = torch.arange(W, dtype=dtype).view(1, 1, 1, W).expand(B, K, H, W)
    gx = (xx - disp) / ((W - 1.0) / 2.0) - 1.0
    gy = yy / ((H - 1.0) / 2.0) - 1.0
    ix = ((gx + 1.0) / 2.0) * (W - 1)
    iy = ((gy + 1.0) / 2.0) * (H - 1)
    return ix, iy


def bilinear_zeros(img, ix, iy):
    """grid_sample(mode='bilinear', padding_mode='zeros') at pixel coordinates.
    img (B,C,H,W); ix,iy (B,K,H,W) -> (B,C,K,H,W)."""
    B, C, H, W = img.shape
    K = ix.shape[1]
    x0 = torch.floor(ix)
    y0 = torch.floor(iy)
    x1 = x0 + 1
    y1 = y0 + 1
    w_nw = (x1 - ix) * (y1 - iy)
    w_ne = (ix - x0) * (y1 - iy)
    w_sw = (x1 - ix) * (iy - y0)
    w_se = (ix - x0) * (iy - y0)
    flat = img.reshape(B, C, H * W)

    def corner(xc, yc, wgt):
        inside = (xc >= 0) & (xc <= W - 1) & (yc >= 0) & (yc <= H - 1)
        idx = (yc.clamp(0, H - 1) * W + xc.clamp(0, W - 1)).long().reshape(B, 1, -1).expand(B, C, -1)
        v = flat.gather(2, idx).reshape(B, C, K, H, W)
        return v * (wgt * inside.to(img.dtype)).unsqueeze(1)

    return corner(x0, y0, w_nw) + corner(x1, y0, w_ne) + corner(x0, y1, w_sw) + corner(x1, y1, w_se)


def spatial_transformer_grid(x, y, disp_samples):
    """SpatialTransformer_grid (submodule.py:265-288): returns (y_warped, x_repeated)."""
    B, C, H, W = y.shape
    ix, iy = warp_coords(H, W, disp_samples, x.dtype)
    y_warped = bilinear_zeros(y, ix, iy)
    x_rep = x.unsqueeze(2).expand(-1, -1, disp_samples.shape[1], -1, -1).contiguous()
    return y_warped, x_rep


# ------------------------------------------------------------------------------------
# interpolation restatements (PyTorch semantics, align_corners=False)
# ------------------------------------------------------------------------------------
def _lin_index(n_in, n_out, dtype=torch.float32):
    """src = max(0, (dst+.5)*in/out - .5); i0=floor, i1=min(i0+1,in-1), lam=src-i0."""
    scale = n_in / n_out
    src = ((torch.arange(n_out, dtype=dtype) + 0.5) * scale - 0.5).clamp_min(0.0)
    i0 = src.floor().long().clamp_max(n_in - 1)
    i1 = (i0 + 1).clamp_max(n_in - 1)
    lam = src - i0.to(dtype)
    return i0, i1, lam


def linear_upsample(t, dim, n_out):
    """1-D linear resize of `t` along `dim` (separable building block of bi/tri-linear)."""
    i0, i1, lam = _lin_index(t.shape[dim], n_out, t.dtype)
    shape = [1] * t.dim()
    shape[dim] = n_out
    lam = lam.view(shape)
    return t.index_select(dim, i0) * (1 - lam) + t.index_select(dim, i1) * lam


def trilinear_upsample(v, size):
    """F.interpolate(v, size, mode='trilinear') as called at SemStereo.py:279."""
    for dim, n in zip((2, 3, 4), size):
        v = linear_upsample(v, dim, n)
    return v


def bilinear_upsample(t, size):
    """F.interpolate(t, size, mode='bilinear') as called at submodule.py:424."""
    for dim, n in zip((2, 3), size):
        t = linear_upsample(t, dim, n)
    return t


# ------------------------------------------------------------------------------------
# K10 : regression_topk
# ------------------------------------------------------------------------------------
def topk_desc_stable(v, k, dim):
    """Indices of the k largest along `dim`, ties broken toward the LOWER index, in
    descending-value order.  (The reference's Tensor.sort is unstable; this is the
    tie rule the CUDA path implements — SURVEY.md section 8c.)"""
    idx = v.sort(dim=dim, descending=True, stable=True)[1]
    return idx.narrow(dim, 0, k)


def regression_topk(cost, disp_samples, k):
    """regression_topk (submodule.py:434-442): softmax over the k largest costs, expectation
    of their disparity samples.  cost, disp_samples (B,D,H,W) -> (B,1,H,W)."""
    ind = topk_desc_stable(cost, k, 1)
    c = torch.gather(cost, 1, ind)
    p = F.softmax(c, 1)
    d = torch.gather(disp_samples, 1, ind)
    return (d * p).sum(1, keepdim=True)


# ------------------------------------------------------------------------------------
# K11 : context_upsample / SSR_upsample
# ------------------------------------------------------------------------------------
def context_upsample(depth_low, up_weights):
    """context_upsample (submodule_.py:311-323).  depth_low (B,1,h,w), up_weights (B,9,4h,4w).
    out[y,x] = sum_{ky,kx} w[ky*3+kx,y,x] * depth_low[y//4+ky-1, x//4+kx-1] (zero padded)."""
    B, C, h, w = depth_low.shape
    assert C == 1
    pad = F.pad(depth_low, (1, 1, 1, 1))
    ys = torch.arange(4 * h) // 4
    xs = torch.arange(4 * w) // 4
    out = depth_low.new_zeros(B, 4 * h, 4 * w)
    for ky in range(3):
        for kx in range(3):
            tap = pad[:, 0].index_select(1, ys + ky).index_select(2, xs + kx)
            out = out + tap * up_weights[:, ky * 3 + kx]
    return out


def bn_eval_affine(p, prefix, eps=1e-5):
    """Eval-mode BatchNorm as per-channel (scale, shift) from running statistics."""
    scale = p[prefix + ".weight"] / torch.sqrt(p[prefix + ".running_var"] + eps)
    shift = p[prefix + ".bias"] - p[prefix + ".running_mean"] * scale
    return scale, shift


def _bn(x, p, prefix):
    return F.batch_norm(x, p[prefix + ".running_mean"], p[prefix + ".running_var"],
                        p[prefix + ".weight"], p[prefix + ".bias"], False, 0.0, 1e-5)


def ssr_upsample(depth_low, weights, pred_label, p, prefix="ssr_upsample"):
    """SSR_upsample.forward (submodule.py:421-431), eval-mode BN.
    depth_low (B,1,h,w); weights = spx_pred (B,6,4h,4w); pred_label (B,6,4h,4w) -> (B,4h,4w)."""
    B, _, h, w = depth_low.shape
    lab = F.softmax(pred_label, dim=1)
    d_up = bilinear_upsample(depth_low, (4 * h, 4 * w))
    d = _bn(d_up, p, prefix + ".conv.0")
    d = F.conv2d(d, p[prefix + ".conv.1.weight"], p[prefix + ".conv.1.bias"], padding=1)
    d = _bn(d, p, prefix + ".conv.2")
    g = F.conv2d(lab * weights, p[prefix + ".conv1.0.weight"], p[prefix + ".conv1.0.bias"])
    g = torch.sigmoid(_bn(g, p, prefix + ".conv1.1"))
    g = F.conv2d(g * weights, p[prefix + ".conv2.0.weight"], p[prefix + ".conv2.0.bias"])
    g = torch.sigmoid(_bn(g, p, prefix + ".conv2.1"))
    res = F.conv2d(d * g, p[prefix + ".conv3.weight"], p[prefix + ".conv3.bias"])
    return (d_up + res).squeeze(1)


# ------------------------------------------------------------------------------------
# K4 / K5 : 3D conv blocks and the windowed attention
# ------------------------------------------------------------------------------------
def conv3d_bn(x, p, conv, bn, stride=1, pad=1, relu=False):
    """convbn_3d (submodule_other.py:845-848) / BasicConv is_3d (submodule.py:89-116), eval BN."""
    y = F.conv3d(x, p[conv + ".weight"], None, stride=stride, padding=pad)
    if bn is not None:
        y = _bn(y, p, bn)
    return F.relu(y) if relu else y


def deconv3d_bn(x, p, conv, bn):
    """ConvTranspose3d(k3,s2,p1,op1,bias=False)+BN3d (SemStereo.py:124-130)."""
    y = F.conv_transpose3d(x, p[conv + ".weight"], None, stride=2, padding=1, output_padding=1)
    return _bn(y, p, bn)


def window_attention3d(x, p, prefix, num_heads, block):
    """attention_block.forward (submodule_other.py:805-837), including its padded / masked branch (:809-812, :822-829, :835-836):
    H and W are zero-padded (bottom / right) to multiples of the window BEFORE the qkv Linear (so a padded token carries the bias),
    scores between a padded and a real token get -1000, and the padding is cropped before the final 1x1x1 conv.
    Faithful to two quirks of the reference: D is never padded (it must divide), and the mask is built with
    `mask[:, -pad_b:, :] = 1; mask[:, :, -pad_r:] = 1` -- when exactly one of pad_b / pad_r is 0, `-0:` selects EVERYTHING, the mask
    is all ones and nothing is masked (the padded tokens then take part in the softmax of their window like real ones).
    qkv output channel = which*C + head*hd + j; tokens of one (bd,bh,bw) window attend to each other; output channel = head*hd + j;
    then 1x1x1 conv with bias; no residual."""
    B, C, D, H0, W0 = x.shape
    bd, bh, bw = block
    if D % bd:
        raise ValueError("window_attention3d: D must be a multiple of the window depth (the reference does not pad it)")
    pad_r, pad_b = (bw - W0 % bw) % bw, (bh - H0 % bh) % bh
    x = F.pad(x, (0, pad_r, 0, pad_b))
    H, W = H0 + pad_b, W0 + pad_r
    nd, nh, nw = D // bd, H // bh, W // bw
    hd = C // num_heads
    t = x.reshape(B, C, nd, bd, nh, bh, nw, bw).permute(0, 2, 4, 6, 3, 5, 7, 1)
    t = t.reshape(B, nd * nh * nw, bd * bh * bw, C)                      # (B, win, tok, C)
    qkv = F.linear(t, p[prefix + ".qkv_3d.weight"], p[prefix + ".qkv_3d.bias"])
    qkv = qkv.reshape(B, -1, bd * bh * bw, 3, num_heads, hd).permute(3, 0, 1, 4, 2, 5)
    q, k, v = qkv[0], qkv[1], qkv[2]                                      # (B, win, head, tok, hd)
    s = (q @ k.transpose(-2, -1)) * (hd ** -0.5)
    if pad_r > 0 or pad_b > 0:
        mask = torch.zeros((H, W), device=x.device)
        mask[H - pad_b if pad_b else 0:, :] = 1                           # `-0:` == everything
        mask[:, W - pad_r if pad_r else 0:] = 1
        mask = mask.reshape(nh, bh, nw, bw).permute(0, 2, 1, 3).reshape(nh * nw, bh * bw)          # (hw windows, hw tokens)
        am = mask.unsqueeze(1) - mask.unsqueeze(2)
        am = torch.where(am != 0, torch.full_like(am, -1000.0), torch.zeros_like(am))             # (nh*nw, bh*bw, bh*bw)
        am = am.repeat(nd, bd, bd)                                                                  # (win, tok, tok): same for every depth
        s = s + am.unsqueeze(0).unsqueeze(2)
    a = torch.softmax(s, dim=-1)
    o = a @ v                                                             # (B, win, head, tok, hd)
    o = o.reshape(B, nd, nh, nw, num_heads, bd, bh, bw, hd).permute(0, 4, 8, 1, 5, 2, 6, 3, 7)
    o = o.reshape(B, C, D, H, W)[:, :, :, :H0, :W0]
    return F.conv3d(o, p[prefix + ".final1x1.weight"], p[prefix + ".final1x1.bias"])


def hourglass(x, p, prefix, block):
    """hourglass / hourglass2 .forward (SemStereo.py:134-143, 173-182)."""
    c1 = conv3d_bn(x, p, prefix + ".conv1.0.0", prefix + ".conv1.0.1", 2, 1, True)
    c2 = conv3d_bn(c1, p, prefix + ".conv2.0.0", prefix + ".conv2.0.1", 1, 1, True)
    c3 = conv3d_bn(c2, p, prefix + ".conv3.0.0", prefix + ".conv3.0.1", 2, 1, True)
    c4 = conv3d_bn(c3, p, prefix + ".conv4.0.0", prefix + ".conv4.0.1", 1, 1, True)
    c4 = window_attention3d(c4, p, prefix + ".attention_block", 16, block)
    r2 = conv3d_bn(c2, p, prefix + ".redir2.0", prefix + ".redir2.1", 1, 0, False)
    c5 = F.relu(deconv3d_bn(c4, p, prefix + ".conv5.0", prefix + ".conv5.1") + r2)
    r1 = conv3d_bn(x, p, prefix + ".redir1.0", prefix + ".redir1.1", 1, 0, False)
    return F.relu(deconv3d_bn(c5, p, prefix + ".conv6.0", prefix + ".conv6.1") + r1)


def classifier(x, p, prefix):
    """classif / classif_att_ (SemStereo.py:228-234): convbn_3d+ReLU, then Conv3d(32->1)."""
    y = conv3d_bn(x, p, prefix + ".0.0", prefix + ".0.1", 1, 1, True)
    return F.conv3d(y, p[prefix + ".2.weight"], None, padding=1)


def patch_conv(vol, p, name="patch"):
    """`patch`: depthwise Conv3d (1,3,3), groups=C, no bias (SemStereo.py:219,274)."""
    return F.conv3d(vol, p[name + ".weight"], None, padding=(0, 1, 1), groups=vol.shape[1])


def channel_att_logits(im, p, prefix):
    """2-D part of channelAtt (SemStereo.py:93-95,100): BasicConv 1x1 (+BN+ReLU) -> Conv2d 1x1 (+bias)."""
    y = F.conv2d(im, p[prefix + ".im_att.0.conv.weight"])
    y = F.relu(_bn(y, p, prefix + ".im_att.0.bn"))
    return F.conv2d(y, p[prefix + ".im_att.1.weight"], p[prefix + ".im_att.1.bias"])


def channel_att(cv, im, p, prefix):
    """channelAtt.forward (SemStereo.py:98-103): sigmoid(gate)[:, :, None] * cv."""
    return torch.sigmoid(channel_att_logits(im, p, prefix)).unsqueeze(2) * cv


def concat_feature(f4, p, prefix="concat_feature"):
    """concat_feature (SemStereo.py:221-223, called at :314-315): BasicConv 3x3 (conv+BN+ReLU) -> Conv2d 3x3 (no bias)."""
    y = F.conv2d(f4, p[prefix + ".0.conv.weight"], None, padding=1)
    y = F.relu(_bn(y, p, prefix + ".0.bn"))
    return F.conv2d(y, p[prefix + ".1.weight"], None, padding=1)
