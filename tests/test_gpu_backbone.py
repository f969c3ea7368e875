"""MobileViTv2 backbone on the B200 kernels against the HuggingFace implementation of the architecture run on
the CPU in fp32 (oracle/backbone.py) with the same seeded state_dict; kernel-level checks of the non-GEMM pieces against torch.
bf16 activations end to end: tolerance 4 % of max|ref| / 3 % of mean|ref| (measured: 0.5 % at x2 growing to ~2 % at x16 / x32)."""
import pytest
import torch
import torch.nn.functional as F

from semstereo_b200.params import make_backbone_params, make_images

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

if torch.cuda.is_available():
    from semstereo_b200 import ops_tc as tc
    from semstereo_b200.backbone import MobileViTv2Backbone, SemStereoB200


def bf(t):
    return t.to(torch.bfloat16).float()


def rel(a, b):
    d = (a.double() - b.double()).abs()
    return (d.max() / b.abs().max()).item(), (d.mean() / b.abs().mean()).item()


def test_stem_and_depthwise_kernels():
    g = torch.Generator().manual_seed(0)
    img = torch.randn(2, 3, 64, 96, generator=g)
    w = torch.randn(32, 3, 3, 3, generator=g) / 27 ** 0.5
    s, t = torch.rand(32, generator=g) + 0.5, 0.2 * torch.randn(32, generator=g)
    ref = F.silu(F.conv2d(img, w, None, 2, 1) * s.view(1, -1, 1, 1) + t.view(1, -1, 1, 1))
    out = tc.stem_conv(img.to(DEV), w.to(DEV), s.to(DEV), t.to(DEV), 64)
    got = tc.from_blocked2d(out).cpu()
    assert got.shape[1] == 64 and float(got[:, 32:].abs().max()) == 0.0
    assert (got[:, :32] - ref).abs().max().item() <= 2.0 ** -8 * ref.abs().max().item() + 1e-3
    for stride in (1, 2):
        x = torch.randn(2, 64, 32, 48, generator=g)
        wd = torch.randn(64, 1, 3, 3, generator=g) / 3
        s, t = torch.rand(64, generator=g) + 0.5, 0.2 * torch.randn(64, generator=g)
        ref = F.silu(F.conv2d(bf(x), wd, None, stride, 1, groups=64) * s.view(1, -1, 1, 1) + t.view(1, -1, 1, 1))
        got = tc.from_blocked2d(tc.dwconv3x3(tc.to_blocked2d(x.to(DEV)), wd.reshape(64, 9).to(DEV), s.to(DEV), t.to(DEV), stride, tc.SILU)).cpu()
        assert (got - ref).abs().max().item() <= 2.0 ** -8 * ref.abs().max().item() + 1e-3


def test_groupnorm_and_linear_attention_kernels():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 128, 16, 24, generator=g) * 2 + 0.3
    gam, bet = torch.rand(128, generator=g) + 0.5, 0.2 * torch.randn(128, generator=g)
    ref = F.group_norm(bf(x), 1, gam, bet, 1e-5)
    got = tc.from_blocked2d(tc.groupnorm1(tc.to_blocked2d(x.to(DEV)), gam.to(DEV), bet.to(DEV))).cpu()
    assert (got - ref).abs().max().item() <= 2.0 ** -7 * ref.abs().max().item()
    # separable attention: the reference formulation with unfold (transformers MobileViTV2LinearSelfAttention), on bf16-rounded qkv
    d, H, W = 64, 8, 12
    k, v, q = torch.randn(2, d, H, W, generator=g), torch.randn(2, d, H, W, generator=g), 3 * torch.randn(2, 1, H, W, generator=g)
    qkv = torch.cat((k, v, q, torch.zeros(2, 7, H, W)), 1)
    unf = lambda t: F.unfold(bf(t), 2, stride=2).reshape(2, t.shape[1], 4, -1)      # noqa: E731
    score = F.softmax(unf(q), dim=-1)
    ctx = (unf(k) * score).sum(-1, keepdim=True)
    ref = F.fold((F.relu(unf(v)) * ctx).reshape(2, d * 4, -1), (H, W), 2, stride=2)
    got = tc.from_blocked2d(tc.linear_attention(tc.to_blocked2d(qkv.to(DEV)), d)).cpu()
    assert (got - ref).abs().max().item() <= 2.0 ** -7 * ref.abs().max().item() + 1e-3


def test_pointwise_conv_with_silu_and_residual():
    g = torch.Generator().manual_seed(2)
    x, r = torch.randn(2, 128, 24, 16, generator=g), torch.randn(2, 192, 24, 16, generator=g)
    w = torch.randn(192, 128, generator=g) / 128 ** 0.5
    s, t = torch.rand(192, generator=g) + 0.5, 0.2 * torch.randn(192, generator=g)
    y = F.conv2d(bf(x), bf(w).view(192, 128, 1, 1)) * s.view(1, -1, 1, 1) + t.view(1, -1, 1, 1)
    wp = tc.pack_weight2d(w.view(192, 128, 1, 1), tc.CONV1).to(DEV)
    xb, rb = tc.to_blocked2d(x.to(DEV)), tc.to_blocked2d(r.to(DEV))
    a = tc.from_blocked2d(tc.conv2d_tc(tc.CONV1, xb, wp, 192, s.to(DEV), t.to(DEV), act=tc.SILU)).cpu()
    assert (a - F.silu(y)).abs().max().item() <= 2.0 ** -7 * y.abs().max().item() + 2e-3
    b = tc.from_blocked2d(tc.conv2d_tc(tc.CONV1, xb, wp, 192, s.to(DEV), t.to(DEV), residual=rb)).cpu()
    assert (b - (y + bf(r))).abs().max().item() <= 2.0 ** -7 * (y + r).abs().max().item() + 2e-3


@pytest.mark.parametrize("B,H,W", [(2, 128, 128), (1, 256, 192)])
def test_backbone_matches_the_huggingface_architecture(B, H, W):
    from oracle import backbone as ob
    p = make_backbone_params(seed=4)
    img, _ = make_images(5, B, H, W)
    ref = ob.forward(p, img)
    m = MobileViTv2Backbone()
    m.load_state_dict(p, strict=True)
    got = m.to(DEV)(img.to(DEV), as_f32=True)
    torch.cuda.synchronize()
    for i, (a, b) in enumerate(zip(got, ref)):
        assert tuple(a.shape) == tuple(b.shape)
        rmax, rmean = rel(a.cpu(), b)
        print(f"stage {i}: {tuple(b.shape)} max rel {rmax:.4f} mean rel {rmean:.4f}")
        assert rmax <= 4e-2 and rmean <= 3e-2, (i, rmax, rmean)      # bf16 activations through up to ~45 layers: measured 0.5 % (x2) .. 2 % (x32)


@pytest.mark.parametrize("H,W", [(128, 256), (192, 192)])          # 192: multiples of 64 only -> padded + masked attention windows
def test_full_model_images_to_disparity(H, W):
    """SemStereoB200 = backbone + decoder + path: images in, what SemStereo.forward returns out; against the oracles chained on the CPU."""
    from oracle import backbone as ob, decoder as od, hotpath as oh
    from semstereo_b200.params import make_decoder_params, make_params
    sd = dict(make_params(seed=1, peaked=20.0))
    sd.update(make_decoder_params(seed=2))
    pb = make_backbone_params(seed=4)
    sd.update({"feature." + k: v for k, v in pb.items()})
    left, right = make_images(6, 1, H, W)
    fl, fr = ob.forward(pb, left), ob.forward(pb, right)
    d = od.forward(sd, fl, fr, right_label=False)
    ref = oh.forward(sd, {k: d[k] for k in ("f8_l", "f8_r", "f4_l", "f4_r", "spx_pred", "pred_label")}, 64, signed=True)
    model = SemStereoB200(64)
    model.load_state_dict(sd, strict=True)
    assert set(k for k in model.state_dict() if not k.endswith("num_batches_tracked")) == set(sd)
    out = model.to(DEV)(left.to(DEV), right.to(DEV))
    torch.cuda.synchronize()
    disp, label = model.as_model_outputs(out)
    assert tuple(disp[0].shape) == (1, H, W) and tuple(label.shape) == (1, 6, H, W)
    rmax, rmean = rel(label.cpu(), d["pred_label"])
    med = (out["pred_up"].cpu() - ref["pred_up"]).abs().median().item()
    print(f"full model: pred_label max rel {rmax:.4f}, median |pred_up - oracle| {med:.4f} px (1/4-res units)")
    assert rmax <= 6e-2 and med <= 0.15
