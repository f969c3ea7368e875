"""bf16x3 split route (fp32-accurate products on the bf16 tensor cores) -- ops_tc.conv3d_tc_split and precision="split".

Operator level: every layer kind of the attention branch against a float64 torch convolution of the UNROUNDED fp32 operands.
A value is carried as hi = bf16(x), lo = bf16(x - hi): |x - hi - lo| <= 2^-17 |x|, the dropped lo*lo product is 2^-18, so a
sum of K products is off by at most ~2^-16 * sum|x||w|; measured 10-100x smaller (errors are random).  Tolerance in the tests:
2e-5 * max|y| (plain bf16 operands give 4e-3).

Whole path (the benchmarked default mode): sample selection `ind_k` equals the fp32 oracle's on >= 99.9 % of the pixels and the
kept probabilities agree to 2e-5 where it does (the selection is decided by the attention
branch alone)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import hotpath as oh
from semstereo_b200.params import make_inputs, make_params

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ORDER = ("f8_l", "f8_r", "f4_l", "f4_r", "cf_l", "cf_r", "spx_pred", "pred_label")

if torch.cuda.is_available():
    from semstereo_b200 import ops_tc as tc
    from semstereo_b200.hotpath import DisparityHotPath


def join(xs):
    """split tensor (2B, ...) bf16 -> fp32 hi + lo (B, ...)."""
    B = xs.shape[0] // 2
    return xs[:B].float() + xs[B:].float()


def from_split(xs):
    B = xs.shape[0] // 2
    return tc.from_blocked_bf16(xs[:B].contiguous()).cpu().double() + tc.from_blocked_bf16(xs[B:].contiguous()).cpu().double()


def packs(w, kind, two_launch=False):
    """pack_weight_split on the device; two_launch drops the in-kernel packing so that the two-launch route is taken."""
    hi, lo, both = tc.pack_weight_split(w.to(DEV), kind)
    return (hi, lo, None if two_launch else both)


def close(got, ref, rel=2e-5):
    err = (got.double() - ref.double()).abs().max().item()
    assert err <= rel * ref.abs().max().item() + 1e-7, f"max err {err:.3e} vs max|ref| {ref.abs().max().item():.3e}"


def test_split_converters():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 32, 4, 6, 8, generator=g) * 3
    for s2d in (False, True):
        xs = tc.to_blocked_bf16(x.to(DEV), s2d=s2d, split=True)
        assert xs.shape[0] == 4
        hi = tc.to_blocked_bf16(x.to(DEV), s2d=s2d)
        assert torch.equal(xs[:2].cpu(), hi.cpu())
        lo_ref = tc.to_blocked_bf16((x - x.bfloat16().float()).to(DEV), s2d=s2d)
        assert torch.equal(xs[2:].cpu(), lo_ref.cpu())
    rec = from_split(tc.to_blocked_bf16(x.to(DEV), split=True))
    assert (rec - x.double()).abs().max().item() <= 2.0 ** -16 * x.abs().max().item()


@pytest.mark.parametrize("W", [20, 18])
def test_patch_gate_split(W):
    from semstereo_b200 import ops
    p = make_params(seed=4)
    g = torch.Generator().manual_seed(5)
    vol, logits = torch.randn(2, 32, 6, 10, W, generator=g), torch.randn(2, 32, 10, W, generator=g)
    pw = p["patch.weight"].reshape(32, 9).to(DEV)
    full = ops.patch_gate(vol.to(DEV), pw, logits.to(DEV))
    got = tc.patch_gate_blocked(vol.to(DEV), pw, logits.to(DEV), split=True)
    want = tc.to_blocked_bf16(full, s2d=True, split=True)
    assert torch.equal(got.cpu(), want.cpu())


@pytest.mark.parametrize("kind,Cin,Cout,B,D,H,W", [("s1f", 32, 32, 2, 5, 20, 12), ("s1f", 64, 64, 1, 8, 32, 16), ("s1", 128, 128, 2, 4, 16, 16),
                                                  ("s2", 32, 64, 2, 8, 40, 24), ("s2", 64, 128, 1, 4, 32, 16), ("s1f", 32, 32, 1, 16, 64, 64)])
@pytest.mark.parametrize("out", ["blocked", "f32", "s2d"])
@pytest.mark.parametrize("two_launch", [False, True])
def test_split_conv(kind, Cin, Cout, B, D, H, W, out, two_launch):
    g = torch.Generator().manual_seed(Cin + Cout + D)
    x = torch.randn(B, Cin, D, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, 3, generator=g) / (27 * Cin) ** 0.5
    scale, shift = torch.rand(Cout, generator=g) + 0.5, 0.3 * torch.randn(Cout, generator=g)
    stride = 2 if kind == "s2" else 1
    y = F.conv3d(x.double(), w.double(), None, stride=stride, padding=1)
    ref = F.relu(y * scale.double().view(1, -1, 1, 1, 1) + shift.double().view(1, -1, 1, 1, 1))
    k = {"s1f": tc.S1F, "s1": tc.S1, "s2": tc.S2}[kind]
    xs = tc.to_blocked_bf16(x.to(DEV), s2d=(kind == "s2"), split=True)
    if two_launch:      # the partial-sum hooks are compiled only into the kernels of the layers without an in-kernel configuration
        pytest.skip("covered by the two_launch=False case" if not tc.split_supported(k, Cin, Cout) else "layer runs the in-kernel route")
    ws = packs(w, k, two_launch)
    if out == "s2d" and any(v % 2 for v in ref.shape[2:]):
        pytest.skip("phase-split output needs even output dims")
    mode = {"blocked": tc.BLOCKED, "f32": tc.F32, "s2d": tc.S2D}[out]
    for _ in range(2):
        o = tc.conv3d_tc_split(k, xs, ws, Cout, scale.to(DEV), shift.to(DEV), relu=True, out_mode=mode)
        torch.cuda.synchronize()
        if out == "f32":
            close(o.cpu(), ref)
        elif out == "blocked":
            close(from_split(o), ref)
        else:       # the phase-split split output is the permutation of the blocked one
            b = tc.conv3d_tc_split(k, xs, ws, Cout, scale.to(DEV), shift.to(DEV), relu=True)
            Bo = b.shape[0] // 2
            assert torch.equal(o[:Bo].cpu(), tc.blocked_to_s2d(b[:Bo].contiguous()).cpu())
            assert torch.equal(o[Bo:].cpu(), tc.blocked_to_s2d(b[Bo:].contiguous()).cpu())


@pytest.mark.parametrize("Cin,Cout,B,D,H,W", [(128, 64, 1, 2, 16, 8), (64, 32, 2, 4, 24, 16), (64, 32, 1, 8, 64, 64), (64, 32, 2, 3, 20, 12)])
@pytest.mark.parametrize("two_launch", [False, True])
def test_split_transposed_with_fused_skip(Cin, Cout, B, D, H, W, two_launch):
    g = torch.Generator().manual_seed(Cin + D + 7)
    x = torch.randn(B, Cin, D, H, W, generator=g)
    skip = torch.randn(B, Cout, 2 * D, 2 * H, 2 * W, generator=g)
    w = torch.randn(Cin, Cout, 3, 3, 3, generator=g) / (27 * Cin / 8) ** 0.5
    wr = torch.randn(Cout, Cout, generator=g) / Cout ** 0.5
    t = 0.3 * torch.randn(Cout, generator=g)
    ref = F.relu(F.conv_transpose3d(x.double(), w.double(), None, stride=2, padding=1, output_padding=1)
                 + F.conv3d(skip.double(), wr.double().view(Cout, Cout, 1, 1, 1)) + t.double().view(1, -1, 1, 1, 1))
    if two_launch:
        pytest.skip("covered by the two_launch=False case" if not tc.split_supported(tc.T2, Cin, Cout) else "layer runs the in-kernel route")
    for _ in range(2):
        o = tc.conv3d_tc_split(tc.T2, tc.to_blocked_bf16(x.to(DEV), split=True), packs(w, tc.T2, two_launch), Cout, None, t.to(DEV),
                               residual_s2d=tc.to_blocked_bf16(skip.to(DEV), s2d=True, split=True),
                               skip_split=tc.pack_skip_weight_split(wr.to(DEV)), relu=True)
        torch.cuda.synchronize()
        close(from_split(o), ref)


@pytest.mark.parametrize("B,D,H,W", [(2, 6, 24, 20), (1, 16, 64, 64)])
def test_split_head(B, D, H, W):
    g = torch.Generator().manual_seed(B + D + H)
    x = torch.randn(B, 32, D, H, W, generator=g)
    w = torch.randn(1, 32, 3, 3, 3, generator=g) / (27 * 32) ** 0.5
    ref = F.conv3d(x.double(), w.double(), None, padding=1)
    for _ in range(2):
        o = tc.conv3d_tc_head(tc.to_blocked_bf16(x.to(DEV), split=True), tc.pack_head_weight_split(w).to(DEV), in_split=True)
        torch.cuda.synchronize()
        close(o.cpu(), ref)
    # the two-launch form (partial sums through the acc_in hooks) gives the same result
    xs = tc.to_blocked_bf16(x.to(DEV), split=True)
    hi, lo = tc.split_f32(w)
    part = tc.conv3d_tc_head(xs, tc.pack_head_weight(hi).to(DEV))
    o2 = tc.conv3d_tc_head(xs[:B], tc.pack_head_weight(lo).to(DEV), part[:B], part[B:])
    close(o2.cpu(), ref)


def run(m, inp, keep=True):
    out = m(*[inp[k].to(DEV) for k in ORDER], keep=keep)
    torch.cuda.synchronize()
    return {k: v.cpu() for k, v in out.items() if v is not None}


def selection_agreement(out, ref):
    same = (out["ind_k"] == ref["ind_k"]).all(dim=2)                                # (B,1,H,W)
    m = same.squeeze(1)
    e_att = (out["att_topk"].reshape(ref["att_topk"].shape)[:, 0].permute(0, 2, 3, 1)[m]
             - ref["att_topk"][:, 0].permute(0, 2, 3, 1)[m]).abs().max().item()
    e_pred = (out["pred_att"].reshape(ref["pred_att"].shape)[m] - ref["pred_att"][m]).abs().max().item()
    return same.float().mean().item(), e_att, e_pred


@pytest.mark.parametrize("signed,maxdisp,B,H,W,peaked", [(True, 64, 2, 128, 128, 20.0), (False, 128, 1, 256, 128, 20.0), (True, 64, 1, 256, 256, 1.0)])
def test_split_mode_selects_the_oracle_samples(signed, maxdisp, B, H, W, peaked):
    p = make_params(seed=9, peaked=peaked, gamma=0.1)
    inp = make_inputs(11, B, H, W)
    ref = oh.forward(p, inp, maxdisp, signed=signed, keep=True)
    m = DisparityHotPath(maxdisp, False, signed, precision="split")
    m.load_state_dict(p, strict=True)
    out = run(m.to(DEV), inp)
    agree, e_att, e_pred = selection_agreement(out, ref)
    cerr = (out["cost_att"] - ref["cost_att"]).abs().max().item()
    print(f"split mode {H}x{W} peaked={peaked}: ind_k agreement {agree:.6f}, cost_att err {cerr:.2e}, att_topk err {e_att:.2e}, pred_att err {e_pred:.2e}")
    assert agree >= 0.999
    assert cerr <= 5e-5 * max(1.0, ref["cost_att"].abs().max().item())
    assert e_att <= 2e-5
    assert e_pred <= 1e-3
    # bf16 aggregation on top of identical samples: statistical tolerance (stated separately, north_star)
    d = (out["pred_up"] - ref["pred_up"]).abs()
    assert d.median().item() <= 0.05


@pytest.mark.parametrize("Cin,Cout,B,H,W", [(256, 128, 2, 16, 24), (128, 32, 1, 40, 8), (128, 384, 1, 128, 32)])
def test_pointwise_split(Cin, Cout, B, H, W):
    """1x1 conv as ONE GEMM over the [hi | lo | hi] K-concat form (channelAtt gate convs, attention projections)."""
    g = torch.Generator().manual_seed(Cin + Cout)
    x = torch.randn(B, Cin, H, W, generator=g) * 2
    w = torch.randn(Cout, Cin, generator=g) / Cin ** 0.5
    scale, shift = torch.rand(Cout, generator=g) + 0.5, 0.3 * torch.randn(Cout, generator=g)
    ref = F.relu(F.conv2d(x.double(), w.double().view(Cout, Cin, 1, 1)) * scale.double().view(1, -1, 1, 1) + shift.double().view(1, -1, 1, 1))
    xt = tc.to_blocked_tri(x.to(DEV))
    hi = tc.to_blocked2d(x.to(DEV))
    assert torch.equal(xt[:, :Cin // 8].cpu(), hi.cpu()) and torch.equal(xt[:, 2 * Cin // 8:].cpu(), hi.cpu())
    o = tc.pointwise_split(xt, tc.pack_pointwise_split(w).to(DEV), Cout, scale.to(DEV), shift.to(DEV), relu=True)
    torch.cuda.synchronize()
    close(o.cpu(), ref)


@pytest.mark.parametrize("B,D,H,W,block", [(2, 4, 8, 8, (4, 4, 4)), (1, 6, 8, 12, (6, 4, 4))])
def test_attention_core_f32(B, D, H, W, block):
    from oracle import ops as oo
    g = torch.Generator().manual_seed(D + H)
    C = 128
    x = torch.randn(B, C, D, H, W, generator=g)
    p = {"a.qkv_3d.weight": torch.randn(3 * C, C, generator=g) / C ** 0.5, "a.qkv_3d.bias": 0.1 * torch.randn(3 * C, generator=g),
         "a.final1x1.weight": torch.randn(C, C, 1, 1, 1, generator=g) / C ** 0.5, "a.final1x1.bias": 0.1 * torch.randn(C, generator=g)}
    ref = oo.window_attention3d(x, p, "a", 16, block)
    qkv = tc.pointwise_split(tc.to_blocked_tri(x.view(B, C, D * H, W).to(DEV)), tc.pack_pointwise_split(p["a.qkv_3d.weight"]).to(DEV), 3 * C,
                             None, p["a.qkv_3d.bias"].to(DEV))
    att = tc.window_attention_core_f32(qkv.view(B, 3 * C, D, H, W), block, 16)
    out = tc.pointwise_split(att.view(B, 3 * C // 8, D * H, W, 8), tc.pack_pointwise_split(p["a.final1x1.weight"]).to(DEV), C, None,
                             p["a.final1x1.bias"].to(DEV))
    torch.cuda.synchronize()
    close(out.view(B, C, D, H, W).cpu(), ref, rel=3e-5)
