"""Tensor-core (tcgen05/TMEM/TMA) 3-D convolutions vs torch fp32 convs on bf16-rounded operands.
The kernels accumulate in fp32, so against an fp32 conv of the SAME bf16-rounded inputs/weights only the summation order and
the bf16 rounding of the output differ: |err| <= 2^-8*|y| + 2e-3*max|y| (bf16 output), 2e-3*max|y| (fp32 output)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

if torch.cuda.is_available():
    from semstereo_b200 import ops_tc as tc


def bf(t):
    return t.to(torch.bfloat16).float()


def s2d_ref(x):
    """(B,C,D,H,W) -> (B,8,C/8,D/2,H/2,W/2,8) phase-split blocked, reference permutation."""
    B, C, D, H, W = x.shape
    t = x.view(B, C // 8, 8, D // 2, 2, H // 2, 2, W // 2, 2).permute(0, 4, 6, 8, 1, 3, 5, 7, 2)
    return t.reshape(B, 8, C // 8, D // 2, H // 2, W // 2, 8)


def check(got, ref, bf16_out):
    tol = 2e-3 * ref.abs().max().item() + (2.0 ** -8) * ref.abs() * (1 if bf16_out else 0) + 1e-4
    bad = ((got - ref).abs() > tol).float().mean().item()
    assert bad == 0.0, f"{bad:.4%} of outputs outside tolerance; max err {(got - ref).abs().max().item():.4f}"


def test_layout_converters():
    x = torch.randn(2, 32, 4, 6, 8, generator=torch.Generator().manual_seed(0))
    xb = tc.to_blocked_bf16(x.to(DEV))
    assert tuple(xb.shape) == (2, 4, 4, 6, 8, 8)
    assert torch.equal(xb.cpu().float(), bf(x).view(2, 4, 8, 4, 6, 8).permute(0, 1, 3, 4, 5, 2))
    assert torch.equal(tc.from_blocked_bf16(xb).cpu(), bf(x))
    xs = tc.to_blocked_bf16(x.to(DEV), s2d=True)
    assert torch.equal(xs.cpu().float(), s2d_ref(bf(x)))
    assert torch.equal(tc.blocked_to_s2d(xb).cpu(), xs.cpu())


S1_CASES = [(32, 32, 1, 4, 16, 8), (32, 32, 2, 5, 20, 12), (64, 32, 1, 6, 32, 24), (64, 64, 1, 4, 16, 16), (128, 128, 1, 4, 8, 8),
            (128, 128, 2, 6, 16, 16), (64, 32, 1, 24, 64, 64), (32, 32, 1, 16, 128, 128), (32, 1, 1, 6, 16, 24)]


@pytest.mark.parametrize("Cin,Cout,B,D,H,W", S1_CASES)
@pytest.mark.parametrize("out_f32", [False, True])
def test_conv_k3_s1(Cin, Cout, B, D, H, W, out_f32):
    if Cout % 8 and not out_f32:
        pytest.skip("bf16 blocked output needs Cout % 8 == 0")
    g = torch.Generator().manual_seed(Cin + Cout + D)
    x = torch.randn(B, Cin, D, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, 3, generator=g) / (27 * Cin) ** 0.5
    scale, shift = torch.rand(Cout, generator=g) + 0.5, 0.3 * torch.randn(Cout, generator=g)
    gate = torch.randn(B, Cout, H, W, generator=g)
    y = F.conv3d(bf(x), bf(w), None, padding=1)
    ref = F.relu(y * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1)) * torch.sigmoid(gate).unsqueeze(2)
    xb = tc.to_blocked_bf16(x.to(DEV))
    wt = tc.pack_weight(w, tc.S1).to(DEV)
    gb = tc.gate_sigmoid_blocked(gate.to(DEV)) if Cout % 8 == 0 else None
    if gb is None:
        ref = F.relu(y * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1))
    out = tc.conv3d_tc(tc.S1, xb, wt, Cout, scale.to(DEV), shift.to(DEV), gb, relu=True, out_mode=tc.F32 if out_f32 else tc.BLOCKED)
    torch.cuda.synchronize()
    check(out.cpu() if out_f32 else tc.from_blocked_bf16(out).cpu(), ref, not out_f32)
    plain = tc.conv3d_tc(tc.S1, xb, wt, Cout, out_mode=tc.F32)
    check(plain.cpu(), y, False)
    if Cout % 8 == 0 and D % 2 == 0 and H % 2 == 0 and W % 2 == 0:      # phase-split output == permutation of the blocked one
        a = tc.conv3d_tc(tc.S1, xb, wt, Cout, scale.to(DEV), shift.to(DEV), relu=True, out_mode=tc.S2D)
        b = tc.blocked_to_s2d(tc.conv3d_tc(tc.S1, xb, wt, Cout, scale.to(DEV), shift.to(DEV), relu=True))
        assert torch.equal(a.cpu(), b.cpu())


@pytest.mark.parametrize("C,B,D,H,W", [(32, 2, 4, 16, 24), (64, 1, 6, 20, 12)])
def test_conv_k1(C, B, D, H, W):
    g = torch.Generator().manual_seed(C + D)
    x = torch.randn(B, C, D, H, W, generator=g)
    w = torch.randn(C, C, 1, 1, 1, generator=g) / C ** 0.5
    scale, shift = torch.rand(C, generator=g) + 0.5, 0.3 * torch.randn(C, generator=g)
    ref = F.conv3d(bf(x), bf(w)) * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1)
    out = tc.conv3d_tc(tc.K1, tc.to_blocked_bf16(x.to(DEV)), tc.pack_weight(w, tc.K1).to(DEV), C, scale.to(DEV), shift.to(DEV))
    check(tc.from_blocked_bf16(out).cpu(), ref, True)
    # position-wise layer on a phase-split tensor viewed as batch*8 keeps the phase-split layout
    xs = tc.to_blocked_bf16(x.to(DEV), s2d=True)
    outs = tc.conv3d_tc(tc.K1, tc.s2d_as_batch(xs), tc.pack_weight(w, tc.K1).to(DEV), C, scale.to(DEV), shift.to(DEV))
    assert torch.equal(outs.view(xs.shape).cpu(), tc.blocked_to_s2d(out).cpu())


@pytest.mark.parametrize("Cin,Cout,B,D,H,W", [(32, 64, 1, 4, 16, 16), (32, 64, 2, 8, 40, 24), (64, 128, 1, 4, 32, 16), (64, 128, 2, 12, 16, 32),
                                              (32, 64, 1, 24, 64, 64)])
@pytest.mark.parametrize("out_f32", [False, True])
def test_conv_k3_s2(Cin, Cout, B, D, H, W, out_f32):
    g = torch.Generator().manual_seed(Cin + Cout + D + 1)
    x = torch.randn(B, Cin, D, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, 3, generator=g) / (27 * Cin) ** 0.5
    scale, shift = torch.rand(Cout, generator=g) + 0.5, 0.3 * torch.randn(Cout, generator=g)
    ref = F.relu(F.conv3d(bf(x), bf(w), None, stride=2, padding=1) * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1))
    xs = tc.to_blocked_bf16(x.to(DEV), s2d=True)
    out = tc.conv3d_tc(tc.S2, xs, tc.pack_weight(w, tc.S2).to(DEV), Cout, scale.to(DEV), shift.to(DEV), relu=True,
                       out_mode=tc.F32 if out_f32 else tc.BLOCKED)
    torch.cuda.synchronize()
    check(out.cpu() if out_f32 else tc.from_blocked_bf16(out).cpu(), ref, not out_f32)


@pytest.mark.parametrize("W", [20, 18, 128])          # W % 4 == 0 takes the 4-pixel-per-thread patch-gate kernel
def test_blocked_producers(W):
    from oracle import ops as oo
    from semstereo_b200 import ops
    from semstereo_b200.params import make_params
    p = make_params(seed=4)
    g = torch.Generator().manual_seed(5)
    vol, logits = torch.randn(2, 32, 6, 10, W, generator=g), torch.randn(2, 32, 10, W, generator=g)
    ref = torch.sigmoid(logits).unsqueeze(2) * oo.patch_conv(vol, p)
    got = tc.patch_gate_blocked(vol.to(DEV), p["patch.weight"].reshape(32, 9).to(DEV), logits.to(DEV))
    assert torch.equal(got.cpu(), tc.to_blocked_bf16(ops.patch_gate(vol.to(DEV), p["patch.weight"].reshape(32, 9).to(DEV), logits.to(DEV)), s2d=True).cpu())
    assert (got.cpu().float() - s2d_ref(ref)).abs().max() <= 2e-2
    gb = tc.gate_sigmoid_blocked(logits.to(DEV)).cpu()
    assert (gb - torch.sigmoid(logits).view(2, 4, 8, 10, W).permute(0, 1, 3, 4, 2)).abs().max() <= 1e-6
    cfl, cfr = torch.randn(2, 32, 9, 20, generator=g), torch.randn(2, 32, 9, 20, generator=g)
    d = torch.randint(-6, 7, (2, 24, 9, 20), generator=g).float()
    a = torch.rand(2, 24, 9, 20, generator=g)
    v32 = ops.sparse_concat_volume(cfl.to(DEV), cfr.to(DEV), d.to(DEV), a.to(DEV))
    vb = tc.sparse_concat_volume_blocked(cfl.to(DEV), cfr.to(DEV), d.to(DEV), a.to(DEV))
    assert torch.equal(vb.cpu(), tc.to_blocked_bf16(v32).cpu())


@pytest.mark.parametrize("Cin,Cout,B,D,H,W", [(128, 64, 1, 2, 16, 8), (128, 64, 2, 3, 20, 12), (64, 32, 1, 4, 16, 16), (64, 32, 2, 6, 24, 40),
                                              (64, 32, 1, 12, 64, 64)])
@pytest.mark.parametrize("with_res", [False, True])
def test_conv_transposed(Cin, Cout, B, D, H, W, with_res):
    g = torch.Generator().manual_seed(Cin + Cout + D + 2)
    x = torch.randn(B, Cin, D, H, W, generator=g)
    w = torch.randn(Cin, Cout, 3, 3, 3, generator=g) / (27 * Cin / 8) ** 0.5
    scale, shift = torch.rand(Cout, generator=g) + 0.5, 0.3 * torch.randn(Cout, generator=g)
    y = F.conv_transpose3d(bf(x), bf(w), None, stride=2, padding=1, output_padding=1)
    res = torch.randn(y.shape, generator=g)
    ref = y * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1)
    rs = None
    if with_res:
        ref = ref + bf(res)
        rs = tc.to_blocked_bf16(res.to(DEV), s2d=True)
    ref = F.relu(ref)
    out = tc.conv3d_tc(tc.T2, tc.to_blocked_bf16(x.to(DEV)), tc.pack_weight(w, tc.T2).to(DEV), Cout, scale.to(DEV), shift.to(DEV),
                       residual_s2d=rs, relu=True)
    torch.cuda.synchronize()
    check(tc.from_blocked_bf16(out).cpu(), ref, True)


@pytest.mark.parametrize("B,D,H,W", [(1, 4, 16, 8), (2, 6, 24, 20), (1, 24, 64, 64), (1, 1, 16, 16), (1, 2, 40, 8), (3, 16, 32, 48), (1, 40, 16, 16)])
def test_conv_head_taps_as_n(B, D, H, W):
    g = torch.Generator().manual_seed(B + D + H)
    x = torch.randn(B, 32, D, H, W, generator=g)
    w = torch.randn(1, 32, 3, 3, 3, generator=g) / (27 * 32) ** 0.5
    ref = F.conv3d(bf(x), bf(w), None, padding=1)
    xb, wt = tc.to_blocked_bf16(x.to(DEV)), tc.pack_head_weight(w).to(DEV)
    for _ in range(2):          # twice: no dependence on TMEM state left by the previous launch
        out = tc.conv3d_tc_head(xb, wt)
        torch.cuda.synchronize()
        check(out.cpu(), ref, False)


@pytest.mark.parametrize("Cin,Cout,B,D,H,W", [(128, 64, 1, 2, 16, 8), (64, 32, 2, 4, 24, 16), (64, 32, 1, 12, 64, 64)])
def test_conv_transposed_with_fused_skip_conv(Cin, Cout, B, D, H, W):
    """relu(bn(deconv(x)) + bn(redir(skip))) with the 1x1 redir conv fused into the transposed layer (SemStereo.py:141-142)."""
    g = torch.Generator().manual_seed(Cin + D + 7)
    x = torch.randn(B, Cin, D, H, W, generator=g)
    skip = torch.randn(B, Cout, 2 * D, 2 * H, 2 * W, generator=g)
    w = torch.randn(Cin, Cout, 3, 3, 3, generator=g) / (27 * Cin / 8) ** 0.5
    wr = torch.randn(Cout, Cout, 1, 1, 1, generator=g) / Cout ** 0.5
    s1, t1 = torch.rand(Cout, generator=g) + 0.5, 0.3 * torch.randn(Cout, generator=g)
    s2, t2 = torch.rand(Cout, generator=g) + 0.5, 0.3 * torch.randn(Cout, generator=g)
    wf = w * s1.view(1, -1, 1, 1, 1)
    wrf = wr * s2.view(-1, 1, 1, 1, 1)
    ref = F.relu(F.conv_transpose3d(bf(x), bf(wf), None, stride=2, padding=1, output_padding=1) + F.conv3d(bf(skip), bf(wrf))
                 + (t1 + t2).view(1, -1, 1, 1, 1))
    out = tc.conv3d_tc(tc.T2, tc.to_blocked_bf16(x.to(DEV)), tc.pack_weight(wf, tc.T2).to(DEV), Cout, None, (t1 + t2).to(DEV),
                       residual_s2d=tc.to_blocked_bf16(skip.to(DEV), s2d=True), skip_weight=tc.pack_skip_weight(wr, s2).to(DEV), relu=True)
    torch.cuda.synchronize()
    check(tc.from_blocked_bf16(out).cpu(), ref, True)


@pytest.mark.parametrize("Cin,Cout,B,H,W", [(128, 64, 1, 16, 8), (128, 64, 2, 40, 24), (64, 32, 1, 32, 32), (64, 32, 2, 20, 44),
                                            (128, 64, 1, 256, 256)])
def test_conv2d_3x3(Cin, Cout, B, H, W):
    """kind 4: a 2-D 3x3 convolution (concat_feature, SemStereo.py:221-223) as a depth-1 volume with 9 taps."""
    g = torch.Generator().manual_seed(Cin + H)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / (9 * Cin) ** 0.5
    scale, shift = torch.rand(Cout, generator=g) + 0.5, 0.3 * torch.randn(Cout, generator=g)
    y = F.conv2d(bf(x), bf(w), None, padding=1)
    ref = F.relu(y * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    xb = tc.to_blocked_bf16(x.unsqueeze(2).to(DEV))
    wt = tc.pack_weight(w, tc.C2D).to(DEV)
    out = tc.conv3d_tc(tc.C2D, xb, wt, Cout, scale.to(DEV), shift.to(DEV), relu=True)
    plain = tc.conv3d_tc(tc.C2D, xb, wt, Cout, out_mode=tc.F32)
    torch.cuda.synchronize()
    check(tc.from_blocked_bf16(out).cpu().squeeze(2), ref, True)
    check(plain.cpu().squeeze(2), y, False)


@pytest.mark.parametrize("Cin,Cout,B,D,H,W", [(32, 32, 1, 4, 16, 8), (64, 32, 2, 5, 20, 12), (64, 32, 1, 24, 64, 64), (32, 32, 1, 16, 128, 128),
                                              (32, 64, 1, 6, 16, 24), (64, 64, 1, 12, 32, 32), (64, 64, 2, 20, 40, 24), (64, 32, 1, 40, 16, 16),
                                              (32, 32, 3, 1, 16, 16), (64, 32, 1, 2, 24, 8)])
def test_conv_k3_s1_depth_folded(Cin, Cout, B, D, H, W):
    """kind 5 (depth taps folded into N, TMEM accumulator ring) gives what kind 0 gives; D = 40 wraps the 16-block ring twice."""
    g = torch.Generator().manual_seed(Cin + Cout + D)
    x = torch.randn(B, Cin, D, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, 3, generator=g) / (27 * Cin) ** 0.5
    scale, shift = torch.rand(Cout, generator=g) + 0.5, 0.3 * torch.randn(Cout, generator=g)
    gate = torch.randn(B, Cout, H, W, generator=g)
    y = F.conv3d(bf(x), bf(w), None, padding=1)
    ref = F.relu(y * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1)) * torch.sigmoid(gate).unsqueeze(2)
    xb = tc.to_blocked_bf16(x.to(DEV))
    wt = tc.pack_weight(w, tc.S1F).to(DEV)
    gb = tc.gate_sigmoid_blocked(gate.to(DEV))
    for _ in range(2):          # twice: the second launch must not depend on TMEM state left by the first
        out = tc.conv3d_tc(tc.S1F, xb, wt, Cout, scale.to(DEV), shift.to(DEV), gb, relu=True)
        plain = tc.conv3d_tc(tc.S1F, xb, wt, Cout, out_mode=tc.F32)
        torch.cuda.synchronize()
        check(tc.from_blocked_bf16(out).cpu(), ref, True)
        check(plain.cpu(), y, False)


@pytest.mark.parametrize("B,H,W,dmin", [(1, 16, 8, -16), (2, 40, 24, -16), (1, 64, 64, 0), (1, 128, 128, -16)])
def test_concat_stem_with_the_volume_generated_in_kernel(B, H, W, dmin):
    """ss_concat_stem_fused == sparse_concat_volume_blocked + kind-5 conv (volume never written to HBM).  The fused kernel reads
    cf as bf16 (one more rounding than the fp32 cf of the two-kernel route) -> compared at the bf16 tolerance of this file."""
    g = torch.Generator().manual_seed(H + W)
    K = 24
    cfl, cfr = torch.randn(B, 32, H, W, generator=g), torch.randn(B, 32, H, W, generator=g)
    ind = torch.stack([torch.randperm(32, generator=g)[:K].sort()[0] for _ in range(B * H * W)]).view(B, H, W, K).permute(0, 3, 1, 2)
    disp = (ind + dmin).float().contiguous()
    att = torch.rand(B, K, H, W, generator=g)
    w = torch.randn(32, 64, 3, 3, 3, generator=g) / (27 * 64) ** 0.5
    scale, shift = torch.rand(32, generator=g) + 0.5, 0.3 * torch.randn(32, generator=g)
    gate = torch.randn(B, 32, H, W, generator=g)
    # reference: volume from bf16-rounded cf (what the fused kernel multiplies), integer shift, then an fp32 conv
    xs = torch.arange(W).view(1, 1, 1, W) - disp.long()                                   # (B,K,H,W) source column
    okx = (xs >= 0) & (xs < W)
    right = torch.gather(bf(cfr).unsqueeze(2).expand(B, 32, K, H, W), 4, xs.clamp(0, W - 1).unsqueeze(1).expand(B, 32, K, H, W))
    vol = torch.cat((bf(cfl).unsqueeze(2).expand(B, 32, K, H, W), right * okx.unsqueeze(1)), 1) * att.unsqueeze(1)
    y = F.conv3d(bf(vol), bf(w), None, padding=1)
    ref = F.relu(y * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1)) * torch.sigmoid(gate).unsqueeze(2)
    wt = tc.pack_weight(w, tc.S1F).to(DEV)
    gb = tc.gate_sigmoid_blocked(gate.to(DEV))
    for _ in range(2):
        out = tc.concat_stem_fused(tc.to_blocked2d(cfl.to(DEV)), tc.to_blocked2d(cfr.to(DEV)), disp.to(DEV), att.to(DEV), wt, dmin,
                                   scale.to(DEV), shift.to(DEV), gb, relu=True)
        plain = tc.concat_stem_fused(tc.to_blocked2d(cfl.to(DEV)), tc.to_blocked2d(cfr.to(DEV)), disp.to(DEV), att.to(DEV), wt, dmin,
                                     relu=False, out_mode=tc.F32)
        torch.cuda.synchronize()
        check(tc.from_blocked_bf16(out).cpu(), ref, True)
        check(plain.cpu(), y, False)
