"""Tensor-core attention_block (qkv 1x1 -> per-head softmax core -> final 1x1, bf16 blocked layout) vs the fp32 oracle."""
import pytest
import torch

from oracle import ops as oo
from semstereo_b200.params import make_params

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

if torch.cuda.is_available():
    from semstereo_b200 import ops_tc as tc


@pytest.mark.parametrize("block,D,H,W", [((4, 4, 4), 4, 8, 8), ((6, 4, 4), 6, 8, 12), ((4, 4, 4), 8, 16, 8)])
def test_attention_block_tc(block, D, H, W):
    p = make_params(seed=2)
    x = torch.randn(2, 128, D, H, W, generator=torch.Generator().manual_seed(D + H))
    xq = x.to(torch.bfloat16).float()
    ref = oo.window_attention3d(xq, p, "hourglass.attention_block", 16, block)
    pre = "hourglass.attention_block."
    wq = tc.pack_weight(p[pre + "qkv_3d.weight"].reshape(384, 128, 1, 1, 1), tc.K1).to(DEV)
    wo = tc.pack_weight(p[pre + "final1x1.weight"], tc.K1).to(DEV)
    xb = tc.to_blocked_bf16(x.to(DEV))
    qkv = tc.conv3d_tc(tc.K1, xb, wq, 384, None, p[pre + "qkv_3d.bias"].to(DEV))
    core = tc.window_attention_core(qkv, block, 16)
    out = tc.conv3d_tc(tc.K1, core, wo, 128, None, p[pre + "final1x1.bias"].to(DEV), out_mode=tc.F32)
    torch.cuda.synchronize()
    err = (out.cpu() - ref).abs().max().item()
    assert err <= 3e-2 * ref.abs().max().item() + 1e-2, err


@pytest.mark.parametrize("block,B,D,H,W", [((4, 4, 4), 2, 4, 8, 8), ((6, 4, 4), 1, 12, 8, 12), ((2, 4, 4), 1, 4, 8, 8)])
def test_attention_core_vs_fp32_softmax(block, B, D, H, W):
    """The softmax(q k^T / sqrt(8)) v core alone (mma.sync kernel for 64 / 96-token windows, fp32 kernel otherwise) against an
    fp32 torch evaluation of the SAME bf16 q, k, v: differences = bf16 rounding of the probabilities and of the output."""
    g = torch.Generator().manual_seed(sum(block) + D)
    qkv = (1.5 * torch.randn(B, 384, D, H, W, generator=g)).to(torch.bfloat16).float()
    got = tc.from_blocked_bf16(tc.window_attention_core(tc.to_blocked_bf16(qkv.to(DEV)), block, 16)).cpu()
    bd, bh, bw = block
    t = qkv.view(B, 3, 16, 8, D // bd, bd, H // bh, bh, W // bw, bw).permute(1, 0, 4, 6, 8, 2, 5, 7, 9, 3)
    q, k, v = [x.reshape(B, D // bd, H // bh, W // bw, 16, bd * bh * bw, 8) for x in t]      # (..., head, token, hd)
    att = torch.softmax(q @ k.transpose(-1, -2) * 8 ** -0.5, dim=-1) @ v
    ref = att.view(B, D // bd, H // bh, W // bw, 16, bd, bh, bw, 8).permute(0, 4, 8, 1, 5, 2, 6, 3, 7).reshape(B, 128, D, H, W)
    err = (got - ref).abs().max().item()
    assert err <= 1.5e-2 * ref.abs().max().item(), err
