"""Op-level parity of the CUDA kernels (through the C-ABI) against the oracle and the reference goldens.
Runs on the B200 box:  python -m pytest tests -m gpu"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ops as oo
from oracle.make_golden import op_inputs
from semstereo_b200.params import make_params

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from semstereo_b200 import ops, submodule as sub_s, submodule_ as sub_u

DEV = "cuda:0"


def cu(t):
    return t.to(DEV).contiguous()


def err(a, b):
    a = a.detach().float().cpu()
    b = b.detach().float().cpu() if torch.is_tensor(b) else torch.from_numpy(np.asarray(b)).float()
    assert a.shape == b.shape, (a.shape, b.shape)
    return (a - b).abs().max().item() if a.numel() else 0.0


def rnd(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,C,H,W,M,G", [(1, 16, 5, 24, 4, 4), (2, 32, 3, 37, 6, 8), (1, 64, 4, 128, 8, 32), (1, 256, 2, 128, 8, 32),
                                         (1, 128, 2, 256, 16, 32), (1, 24, 3, 70, 5, 3), (1, 8, 2, 9, 12, 2), (1, 128, 1, 64, 48, 8)])
@pytest.mark.parametrize("signed", [True, False])
@pytest.mark.parametrize("norm", [False, True])
def test_gwc_volume(B, C, H, W, M, G, signed, norm):
    l, r = rnd(B, C, H, W, seed=1), rnd(B, C, H, W, seed=2)
    ref = oo.gwc_volume(l, r, M, G, signed, norm)
    got = ops.gwc_volume(cu(l), cu(r), M, G, signed, norm)
    assert err(got, ref) <= 2e-6
    # the never-written region of the reference volume is exactly zero here too
    assert bool(((ref == 0) <= (got.cpu() == 0)).all())


def test_gwc_norm_single_group_and_symmetry():
    l, r = rnd(1, 48, 3, 40, seed=3), rnd(1, 48, 3, 40, seed=4)
    assert err(ops.gwc_volume(cu(l), cu(r), 6, 1, True, True), oo.norm_correlation_volume(l, r, 6, True)) <= 2e-6
    # correlation symmetry: V(L,R)[d][x] == V(R,L)[-d][x-d], bit-exact (same products, same order)
    M = 6
    a = ops.gwc_volume(cu(l), cu(r), M, 4, True, False).cpu()
    b = ops.gwc_volume(cu(r), cu(l), M, 4, True, False).cpu()
    for d in range(-M + 1, M):
        lo, hi = max(0, d), min(40, 40 + d)
        assert torch.equal(a[:, :, d + M, :, lo:hi], b[:, :, -d + M, :, lo - d:hi - d])


@pytest.mark.parametrize("B,C,H,W,M", [(2, 16, 5, 24, 4), (1, 5, 3, 37, 6), (1, 32, 2, 256, 16)])
@pytest.mark.parametrize("signed", [True, False])
def test_concat_volume_bit_exact(B, C, H, W, M, signed):
    l, r = rnd(B, C, H, W, seed=5), rnd(B, C, H, W, seed=6)
    assert torch.equal(ops.concat_volume(cu(l), cu(r), M, signed).cpu(), oo.concat_volume(l, r, M, signed))


@pytest.mark.parametrize("flavour", ["signed", "unsigned"])
def test_surface_against_reference_goldens(golden_dir, flavour):
    """The reference-named functions reproduce the recorded outputs of the unmodified reference."""
    g = dict(np.load(os.path.join(golden_dir, f"ops_{flavour}.npz")))
    d = {k: cu(v) for k, v in op_inputs().items()}
    s = sub_s if flavour == "signed" else sub_u
    M, G = 4, 4
    assert err(s.build_gwc_volume(d["ref"], d["tgt"], M, G), g["gwc"]) <= 2e-6
    assert err(s.build_gwc_volume_norm(d["ref"], d["tgt"], M, G), g["gwc_norm"]) <= 2e-6
    assert err(s.build_concat_volume(d["ref"], d["tgt"], M), g["concat"]) == 0.0
    assert err(s.build_norm_correlation_volume(d["ref"], d["tgt"], M), g["normcorr"]) <= 2e-6
    nb = 4 if flavour == "signed" else 8
    assert err(s.disparity_regression(d["prob32"], nb), g["regress"]) <= 2e-6
    assert err(s.disparity_variance(d["prob32"], nb, d["mu"]), g["variance"]) <= 2e-5
    assert err(s.Propagation()(d["disp1"]), g["prop"]) == 0.0
    assert err(s.Propagation_prob()(d["vol1"]), g["prop_prob"]) == 0.0
    yw, xr = s.SpatialTransformer_grid(d["feat_l"], d["feat_r"], d["disp_real"])
    assert err(yw, g["stn_real"]) <= 5e-6
    assert torch.equal(xr.cpu(), op_inputs()["feat_l"].unsqueeze(2).expand(-1, -1, 5, -1, -1))
    assert err(s.SpatialTransformer_grid(d["feat_l"], d["feat_r"], d["disp_int"])[0], g["stn_int"]) <= 5e-6
    assert err(s.regression_topk(d["cost24"], d["samples24"], 2), g["topk2"]) <= 5e-6
    assert err(s.regression_topk(d["cost24"], d["samples24"], 3), g["topk3"]) <= 5e-6
    p = make_params(seed=2)
    ssr = s.SSR_upsample(6).eval()
    ssr.load_state_dict({k[len("ssr_upsample."):]: v for k, v in p.items() if k.startswith("ssr_upsample.")})
    assert err(ssr.to(DEV)(d["depth_low"], d["spx"], d["label"]), g["ssr"]) <= 5e-5
    if flavour == "unsigned":
        assert err(s.context_upsample(d["depth_low"], d["up9"]), g["context_up"]) <= 2e-6
    for tag, block in (("444", (4, 4, 4)), ("644", (6, 4, 4))):
        a = s.attention_block(128, 16, block).eval()
        a.load_state_dict({k[len("hourglass.attention_block."):]: v for k, v in p.items() if k.startswith("hourglass.attention_block.")})
        assert err(a.to(DEV)(d["att_in_" + tag]), g["att_" + tag]) <= 5e-5


def test_groupwise_helpers():
    l, r = rnd(2, 16, 5, 24, seed=7), rnd(2, 16, 5, 24, seed=8)
    ref = (l * r).view(2, 4, 4, 5, 24).mean(2)
    assert err(sub_s.groupwise_correlation(cu(l), cu(r), 4), ref) <= 2e-6
    lg, rg = l.view(2, 4, 4, 5, 24), r.view(2, 4, 4, 5, 24)
    refn = ((lg / (lg.norm(2, 2, True) + 1e-5)) * (rg / (rg.norm(2, 2, True) + 1e-5))).mean(2)
    assert err(sub_s.groupwise_correlation_norm(cu(l), cu(r), 4), refn) <= 2e-6
    assert tuple(sub_s.norm_correlation(cu(l), cu(r)).shape) == (2, 1, 5, 24)


# ----------------------------------------------------------------------------------------------------
def test_patch_gate_and_pointwise():
    p = make_params(seed=4)
    vol, im = rnd(2, 32, 6, 9, 20, seed=9), rnd(2, 256, 9, 20, seed=10)
    logits = oo.channel_att_logits(im, p, "corr_feature_att_8")
    ref = torch.sigmoid(logits).unsqueeze(2) * oo.patch_conv(vol, p)
    w0 = p["corr_feature_att_8.im_att.0.conv.weight"].reshape(128, 256)
    s0, t0 = oo.bn_eval_affine(p, "corr_feature_att_8.im_att.0.bn")
    y = ops.pointwise_conv2d(cu(im), cu(w0), cu(s0), cu(t0), relu=True)
    lg = ops.pointwise_conv2d(y, cu(p["corr_feature_att_8.im_att.1.weight"].reshape(32, 128)), None,
                              cu(p["corr_feature_att_8.im_att.1.bias"]))
    assert err(lg, logits) <= 2e-5
    got = ops.patch_gate(cu(vol), cu(p["patch.weight"].reshape(32, 9)), lg)
    assert err(got, ref) <= 2e-5
    assert err(ops.patch_gate(cu(vol), None, lg), torch.sigmoid(logits).unsqueeze(2) * vol) <= 2e-5
    assert err(ops.patch_gate(cu(vol), cu(p["patch.weight"].reshape(32, 9)), None), oo.patch_conv(vol, p)) <= 2e-5


CONV_CASES = [  # Cin, Cout, D, H, W, k, stride, transposed
    (32, 64, 8, 12, 16, 3, 2, False), (64, 64, 4, 6, 8, 3, 1, False), (64, 128, 4, 6, 8, 3, 2, False),
    (128, 128, 2, 4, 4, 3, 1, False), (128, 64, 2, 3, 4, 3, 2, True), (64, 32, 4, 6, 8, 3, 2, True),
    (32, 32, 8, 12, 16, 1, 1, False), (64, 32, 6, 8, 12, 3, 1, False), (32, 32, 5, 7, 9, 3, 1, False), (32, 64, 5, 7, 9, 3, 2, False),
]


@pytest.mark.parametrize("Cin,Cout,D,H,W,k,stride,transposed", CONV_CASES)
def test_conv3d_f32(Cin, Cout, D, H, W, k, stride, transposed):
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, Cin, D, H, W, generator=g)
    wshape = (Cin, Cout, 3, 3, 3) if transposed else (Cout, Cin, k, k, k)
    w = torch.randn(*wshape, generator=g) / (Cin * k ** 3) ** 0.5
    scale, shift = torch.rand(Cout, generator=g) + 0.5, torch.randn(Cout, generator=g)
    if transposed:
        y = F.conv_transpose3d(x, w, None, stride=2, padding=1, output_padding=1)
    else:
        y = F.conv3d(x, w, None, stride=stride, padding=k // 2)
    res = torch.randn(y.shape, generator=g)
    gate = torch.randn(2, Cout, y.shape[3], y.shape[4], generator=g)
    ref = F.relu(y * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1) + res) * torch.sigmoid(gate).unsqueeze(2)
    got = ops.conv3d_f32(cu(x), cu(ops.pack_conv3d_weight(w, transposed)), cu(scale), cu(shift), cu(res), cu(gate),
                         k=k, stride=stride, transposed=transposed, relu=True)
    assert err(got, ref) <= 2e-5 * max(1.0, ref.abs().max().item())
    plain = ops.conv3d_f32(cu(x), cu(ops.pack_conv3d_weight(w, transposed)), k=k, stride=stride, transposed=transposed)
    assert err(plain, y) <= 2e-5 * max(1.0, y.abs().max().item())


def test_conv3d_cout1():
    x, w = rnd(2, 32, 5, 7, 10, seed=12), rnd(1, 32, 3, 3, 3, seed=13) / 30
    assert err(ops.conv3d_cout1_f32(cu(x), cu(w)), F.conv3d(x, w, None, padding=1)) <= 2e-5


def test_hourglass_block_against_reference_golden(golden_dir):
    from semstereo_b200.hotpath import DisparityHotPath
    g = dict(np.load(os.path.join(golden_dir, "ops_signed.npz")))
    m = DisparityHotPath(64, precision="fp32")
    m.load_state_dict(make_params(seed=2))
    m.to(DEV)
    out = m._hourglass(m._packed(), "hourglass_att", cu(op_inputs()["hg_in"]))
    assert err(out.reshape(-1)[::7], g["hourglass_sub"]) <= 5e-4


def test_attention_stats_strength_topk_concat():
    p = make_params(seed=5)
    g = torch.Generator().manual_seed(14)
    B, D8, H8, W8 = 2, 16, 5, 9
    cost_att = 3 * torch.randn(B, 1, D8, H8, W8, generator=g)
    for signed, maxdisp in ((True, 64), (False, 128)):
        att, mu, gate = oh_stats(p, cost_att, maxdisp, signed)
        a2, m2, g2 = ops.att_stats(cu(cost_att), cu(p["beta"]), cu(p["gamma"]), float(-maxdisp // 4) if signed else 0.0)
        assert err(a2, att) <= 5e-6 and err(m2, mu) <= 2e-5 and err(g2, gate) <= 1e-5
        f4l, f4r = torch.randn(B, 128, 2 * H8, 2 * W8, generator=g), torch.randn(B, 128, 2 * H8, 2 * W8, generator=g)
        from oracle import hotpath as oh
        st = oh.sample_strength(f4l, f4r, mu, gate)
        st2 = ops.sample_strength(cu(f4l), cu(f4r), cu(mu), cu(gate))
        assert err(st2, st) <= 2e-5
        sel = oh.topk_select(att, st, maxdisp, signed)
        ind, atk, dtk, pa, prob = ops.topk_select(cu(att), cu(st), 24, (maxdisp // 4) if signed else 0, True, True)
        assert err(prob, sel["prob"]) <= 2e-6
        assert torch.equal(ind.cpu(), sel["ind_k"]), "top-k indices must be bit-exact on identical inputs"
        assert err(atk, sel["att_topk"]) <= 2e-6 and torch.equal(dtk.cpu(), sel["disp_topk"]) and err(pa, sel["pred_att"]) <= 2e-5
        cfl, cfr = torch.randn(B, 32, 2 * H8, 2 * W8, generator=g), torch.randn(B, 32, 2 * H8, 2 * W8, generator=g)
        vol = oh.sparse_concat_volume(cfl, cfr, sel["disp_topk"], sel["att_topk"])
        assert err(ops.sparse_concat_volume(cu(cfl), cu(cfr), dtk, atk), vol) <= 5e-6


def test_sample_strength_wide_rows():
    """Row widths that take the 128-bit staging path (W % 4 == 0): two full pixel blocks, a ragged second block, one short block;
    channel counts that are not a multiple of the 32-channel chunk; disparities that leave the +-24 column window (global-memory
    fallback) and the image (zero padding)."""
    from oracle import hotpath as oh
    g = torch.Generator().manual_seed(41)
    for C, H, W in ((128, 6, 256), (40, 5, 136), (24, 4, 20), (32, 3, 260)):
        f4l, f4r = torch.randn(2, C, H, W, generator=g), torch.randn(2, C, H, W, generator=g)
        mu = 12 * torch.randn(2, H, W, generator=g)
        gate = torch.rand(2, 1, H, W, generator=g)
        st = oh.sample_strength(f4l, f4r, mu, gate)
        st2 = ops.sample_strength(cu(f4l), cu(f4r), cu(mu), cu(gate))
        assert err(st2, st) <= 2e-5, (C, H, W)


def oh_stats(p, cost_att, maxdisp, signed):
    from oracle import hotpath as oh
    return oh.attention_stats(p, cost_att, maxdisp, (2 * cost_att.shape[3], 2 * cost_att.shape[4]), signed)


def test_topk_tie_rule_lower_index_wins():
    att = torch.zeros(1, 1, 32, 4, 4)                   # all bins tied -> the 24 lowest bins are kept
    st = torch.full((1, 5, 4, 4), 0.2)
    ind, atk, dtk, pa, _ = ops.topk_select(cu(att), cu(st), 24, 16.0)
    assert torch.equal(ind.cpu()[0, 0, :, 0, 0], torch.arange(24))
    assert torch.equal(dtk.cpu()[0, :, 1, 1], torch.arange(24).float() - 16)
    c = torch.zeros(1, 24, 3, 3)
    s = torch.arange(24).float().view(1, 24, 1, 1).expand(1, 24, 3, 3).contiguous()
    assert err(ops.regression_topk(cu(c), cu(s), 2), torch.full((1, 1, 3, 3), 0.5)) <= 1e-6


def test_error_behaviour():
    z = torch.zeros(1, 8, 4, 8, device=DEV)
    with pytest.raises(ValueError):
        ops.gwc_volume(z, z, 2, 3)                      # C % groups != 0  (reference: assert, submodule.py:192)
    with pytest.raises(AssertionError):
        sub_s.disparity_regression(torch.zeros(1, 8, 4, device=DEV), 4)
    with pytest.raises(RuntimeError):
        sub_s.disparity_regression(torch.zeros(1, 6, 4, 4, device=DEV), 4)   # bins != 2*maxdisp
    with pytest.raises(NotImplementedError):                    # D does not divide by the window (the reference does not pad D)
        ops.window_attention3d(torch.zeros(1, 128, 5, 8, 8, device=DEV), torch.zeros(128, 384, device=DEV), torch.zeros(384, device=DEV),
                               torch.zeros(128, 128, device=DEV), torch.zeros(128, device=DEV), (4, 4, 4))
    with pytest.raises(NotImplementedError):                    # the unmasked kernels refuse a volume that needs the score mask
        ops.window_pad(torch.zeros(1, 128, 4, 6, 10, device=DEV), (4, 4, 4))
    with pytest.raises(RuntimeError):
        ops.gwc_volume(torch.zeros(1, 8, 4, 8), torch.zeros(1, 8, 4, 8), 2, 2)   # CPU tensors: no fallback


def test_ssr_upsample2_equals_two_calls():
    """The fused two-map SSR_upsample (SemStereo.py:312 + :324) is bit-identical to two single-map calls."""
    from semstereo_b200.hotpath import DisparityHotPath, pack_ssr
    from semstereo_b200.params import make_params
    m = DisparityHotPath(64, False, True, precision="fp32")
    m.load_state_dict(make_params(seed=4), strict=True)
    packed = pack_ssr(m.ssr_upsample)
    g = torch.Generator().manual_seed(8)
    B, h, w = 2, 24, 40
    da, db = (8 * torch.randn(B, 1, h, w, generator=g)).to(DEV), (8 * torch.randn(B, 1, h, w, generator=g)).to(DEV)
    spx, lab = torch.randn(B, 6, 4 * h, 4 * w, generator=g).to(DEV), (2 * torch.randn(B, 6, 4 * h, 4 * w, generator=g)).to(DEV)
    oa, ob = ops.ssr_upsample2(da, db, spx, lab, packed)
    assert torch.equal(oa, ops.ssr_upsample(da, spx, lab, packed))
    assert torch.equal(ob, ops.ssr_upsample(db, spx, lab, packed))


def test_torch_library_ops_match_the_wrappers():
    import semstereo_b200.torch_ops  # noqa: F401
    g = torch.Generator().manual_seed(12)
    l, r = torch.randn(1, 64, 8, 32, generator=g).to(DEV), torch.randn(1, 64, 8, 32, generator=g).to(DEV)
    assert torch.equal(torch.ops.semstereo_b200.gwc_volume(l, r, 4, 8, True, True), ops.gwc_volume(l, r, 4, 8, True, True))
    assert torch.equal(torch.ops.semstereo_b200.concat_volume(l, r, 4, False), ops.concat_volume(l, r, 4, False))
    c, d = torch.randn(1, 24, 8, 32, generator=g).to(DEV), torch.randn(1, 24, 8, 32, generator=g).to(DEV)
    assert torch.equal(torch.ops.semstereo_b200.regression_topk(c, d, 2), ops.regression_topk(c, d, 2))


def test_window_attention_on_padded_windows(golden_dir):
    """attention_block on H / W that are not multiples of the window, against outputs of the unmodified reference module
    (tests/golden/att_padded.npz): padding on one axis masks nothing (see ops.window_pad) and runs on the unmasked kernels; padding
    on both axes takes the masked core (-1000 between padded and real tokens, ss_window_attention_core_f32_masked)."""
    from oracle.make_golden_attpad import CASES
    g = dict(np.load(os.path.join(golden_dir, "att_padded.npz")))
    p = make_params(seed=2)
    wq = p["hourglass.attention_block.qkv_3d.weight"].t().contiguous()
    wo = p["hourglass.attention_block.final1x1.weight"].reshape(128, 128).t().contiguous()
    args = (cu(wq), cu(p["hourglass.attention_block.qkv_3d.bias"]), cu(wo), cu(p["hourglass.attention_block.final1x1.bias"]))
    for name, (block, shape) in CASES.items():
        x = cu(torch.from_numpy(g["in_" + name]))
        assert err(ops.window_attention3d(x, *args, block, 16), g["out_" + name]) <= 5e-5, name
        if name.startswith("both"):     # and against the oracle's restatement of the same branch
            want = oo.window_attention3d(torch.from_numpy(g["in_" + name]), p, "hourglass.attention_block", 16, block)
            assert err(ops.window_attention3d(x, *args, block, 16), want) <= 5e-5, name
