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
