"""2-D decoder convolutions on tensor cores (csrc/conv2d_tc.cu) vs torch fp32 convs on the same bf16-rounded operands:
only the summation order (and the bf16 rounding of a bf16 output) differ."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

if torch.cuda.is_available():
    from semstereo_b200 import ops_tc as tc


def bf(t):
    return t.to(torch.bfloat16).float()


def check(got, ref, bf16_out):
    tol = 2e-3 * ref.abs().max().item() + (2.0 ** -8) * ref.abs() * (1 if bf16_out else 0) + 1e-4
    bad = ((got - ref).abs() > tol).float().mean().item()
    assert bad == 0.0, f"{bad:.4%} of outputs outside tolerance; max err {(got - ref).abs().max().item():.4f}"


def test_blocked2d_roundtrip():
    x = torch.randn(2, 64, 12, 20, generator=torch.Generator().manual_seed(0))
    xb = tc.to_blocked2d(x.to(DEV))
    assert tuple(xb.shape) == (2, 8, 12, 20, 8)
    assert torch.equal(tc.from_blocked2d(xb).cpu(), bf(x))


# Cin pieces of the virtual concat, Cout, B, H, W
CONV_CASES = [((64, 0), 32, 1, 16, 8), ((64, 64), 128, 2, 20, 24), ((128, 128), 256, 1, 32, 16), ((384, 384), 768, 1, 16, 16),
              ((256, 256), 512, 1, 24, 40), ((128, 0), 32, 1, 64, 64), ((64, 0), 6, 1, 16, 16), ((192, 64), 64, 1, 18, 10)]


@pytest.mark.parametrize("cins,Cout,B,H,W", CONV_CASES)
@pytest.mark.parametrize("k", [3, 1])
def test_conv2d(cins, Cout, B, H, W, k):
    g = torch.Generator().manual_seed(sum(cins) + Cout + k)
    cin = sum(cins)
    x = torch.randn(B, cin, H, W, generator=g)
    w = torch.randn(Cout, cin, k, k, generator=g) / (k * k * cin) ** 0.5
    scale, shift = torch.rand(Cout, generator=g) + 0.5, 0.3 * torch.randn(Cout, generator=g)
    y = F.conv2d(bf(x), bf(w), None, padding=k // 2)
    ref = F.relu(y * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    mode = tc.CONV3 if k == 3 else tc.CONV1
    x0 = tc.to_blocked2d(x[:, :cins[0]].contiguous().to(DEV))
    x1 = tc.to_blocked2d(x[:, cins[0]:].contiguous().to(DEV)) if cins[1] else None
    wt = tc.pack_weight2d(w, mode).to(DEV)
    plain = tc.conv2d_tc(mode, x0, wt, Cout, out_f32=True, x1=x1)
    torch.cuda.synchronize()
    check(plain.cpu(), y, False)
    if Cout % 8 == 0:
        out = tc.conv2d_tc(mode, x0, wt, Cout, scale.to(DEV), shift.to(DEV), relu=True, x1=x1)
        torch.cuda.synchronize()
        check(tc.from_blocked2d(out).cpu(), ref, True)


DECONV_CASES = [((64, 0), 64, 1, 16, 8), ((128, 0), 6, 1, 16, 24), ((256, 0), 64, 2, 12, 20), ((512, 0), 384, 1, 8, 8),
                ((768, 0), 256, 1, 16, 16), ((512, 0), 128, 1, 32, 32), ((64, 64), 128, 1, 20, 12)]


@pytest.mark.parametrize("cins,Cout,B,H,W", DECONV_CASES)
def test_deconv2d_k4s2(cins, Cout, B, H, W):
    g = torch.Generator().manual_seed(sum(cins) + Cout)
    cin = sum(cins)
    x = torch.randn(B, cin, H, W, generator=g)
    w = torch.randn(cin, Cout, 4, 4, generator=g) / (4 * cin) ** 0.5
    scale, shift = torch.rand(Cout, generator=g) + 0.5, 0.3 * torch.randn(Cout, generator=g)
    y = F.conv_transpose2d(bf(x), bf(w), None, stride=2, padding=1)
    ref = F.relu(y * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    x0 = tc.to_blocked2d(x[:, :cins[0]].contiguous().to(DEV))
    x1 = tc.to_blocked2d(x[:, cins[0]:].contiguous().to(DEV)) if cins[1] else None
    wt = tc.pack_weight2d(w, tc.DECONV4).to(DEV)
    plain = tc.conv2d_tc(tc.DECONV4, x0, wt, Cout, None, shift.to(DEV), out_f32=True, x1=x1)
    torch.cuda.synchronize()
    check(plain.cpu(), y + shift.view(1, -1, 1, 1), False)
    if Cout % 8 == 0:
        out = tc.conv2d_tc(tc.DECONV4, x0, wt, Cout, scale.to(DEV), shift.to(DEV), relu=True, x1=x1)
        torch.cuda.synchronize()
        check(tc.from_blocked2d(out).cpu(), ref, True)


def test_bilinear_up2():
    x = torch.randn(2, 6, 20, 36, generator=torch.Generator().manual_seed(3))
    ref = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
    got = tc.bilinear_up2(x.to(DEV)).cpu()
    assert (got - ref).abs().max().item() <= 1e-6
