"""Tensor-core (tcgen05/TMEM/TMA) conv3d vs torch fp32 conv on bf16-rounded operands.  Tolerance: the kernel accumulates in
fp32, so against an fp32 conv of the SAME bf16-rounded inputs/weights only summation order and the bf16 rounding of the
output differ: |err| <= 2^-8 * |y| + 1e-3 (bf16 output) or 1e-3 * max|y| (fp32 output)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

if torch.cuda.is_available():
    from semstereo_b200 import ops


def bf(t):
    return t.to(torch.bfloat16).float()


def test_blocked_roundtrip():
    x = torch.randn(2, 32, 3, 5, 7, generator=torch.Generator().manual_seed(0))
    xb = ops.to_blocked_bf16(x.to(DEV))
    assert tuple(xb.shape) == (2, 4, 3, 5, 7, 8)
    assert torch.equal(xb.cpu().float(), bf(x).view(2, 4, 8, 3, 5, 7).permute(0, 1, 3, 4, 5, 2))
    assert torch.equal(ops.from_blocked_bf16(xb).cpu(), bf(x))


@pytest.mark.parametrize("Cin,Cout,B,D,H,W", [(32, 32, 1, 4, 16, 8), (32, 32, 2, 5, 20, 12), (64, 32, 1, 6, 32, 24), (64, 64, 1, 4, 16, 16),
                                              (128, 128, 1, 4, 8, 8), (128, 128, 2, 6, 16, 16), (64, 32, 1, 24, 64, 64), (32, 32, 1, 16, 128, 128)])
@pytest.mark.parametrize("out_f32", [False, True])
def test_conv3d_tc_s1(Cin, Cout, B, D, H, W, out_f32):
    g = torch.Generator().manual_seed(Cin + Cout + D)
    x = torch.randn(B, Cin, D, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, 3, generator=g) / (27 * Cin) ** 0.5
    scale, shift = torch.rand(Cout, generator=g) + 0.5, 0.3 * torch.randn(Cout, generator=g)
    gate = torch.randn(B, Cout, H, W, generator=g)
    y = F.conv3d(bf(x), bf(w), None, padding=1)
    ref = F.relu(y * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1)) * torch.sigmoid(gate).unsqueeze(2)
    xb = ops.to_blocked_bf16(x.to(DEV))
    wt = ops.pack_conv3d_weight_tc(w).to(DEV)
    out = ops.conv3d_tc(xb, wt, scale.to(DEV), shift.to(DEV), gate.to(DEV), relu=True, out_f32=out_f32)
    torch.cuda.synchronize()
    got = out.cpu() if out_f32 else ops.from_blocked_bf16(out).cpu()
    tol = 1e-3 * ref.abs().max().item() + (0 if out_f32 else 1) * (2.0 ** -8) * ref.abs()
    bad = ((got - ref).abs() > tol + 1e-3).float().mean().item()
    assert bad == 0.0, f"{bad:.4%} of outputs outside tolerance; max err {(got - ref).abs().max().item():.4f}"
    plain = ops.conv3d_tc(xb, wt, out_f32=True)
    assert (plain.cpu() - y).abs().max().item() <= 2e-3 * max(1.0, y.abs().max().item())
