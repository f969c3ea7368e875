"""Pins the oracle (oracle/ops.py, oracle/hotpath.py) against outputs of the unmodified
reference recorded by oracle/make_golden.py (tests/golden/*.npz).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import hotpath as oh
from oracle import ops as oo
from oracle.make_golden import checksums, op_inputs
from semstereo_b200.params import make_inputs, make_params


def load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name + ".npz")))


def close(a, b, tol, what=""):
    a = a.detach().numpy() if torch.is_tensor(a) else np.asarray(a)
    err = float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)))) if a.size else 0.0
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert err <= tol, f"{what}: max abs err {err:.3e} > {tol:.1e}"
    return err


@pytest.mark.parametrize("flavour", ["signed", "unsigned"])
def test_ops_match_reference(golden_dir, flavour):
    g = load(golden_dir, "ops_" + flavour)
    d = op_inputs()
    for k, v in checksums(d).items():
        assert abs(v - g[k]) <= 1e-9 * max(1.0, abs(g[k])), f"input RNG drift in {k}"
    s = flavour == "signed"
    M, G = 4, 4
    # volume builders: closed form reproduces the reference bit for bit
    close(oo.gwc_volume(d["ref"], d["tgt"], M, G, s, False), g["gwc"], 0.0, "gwc")
    close(oo.gwc_volume(d["ref"], d["tgt"], M, G, s, True), g["gwc_norm"], 1e-7, "gwc_norm")
    close(oo.concat_volume(d["ref"], d["tgt"], M, s), g["concat"], 0.0, "concat")
    close(oo.norm_correlation_volume(d["ref"], d["tgt"], M, s), g["normcorr"], 1e-7, "normcorr")
    nb = 4 if s else 8
    close(oo.disparity_regression(d["prob32"], nb, s), g["regress"], 1e-6, "regress")
    close(oo.disparity_variance(d["prob32"], nb, d["mu"], s), g["variance"], 1e-5, "variance")
    close(oo.propagation(d["disp1"]), g["prop"], 0.0, "prop")
    close(oo.propagation_prob(d["vol1"]), g["prop_prob"], 0.0, "prop_prob")
    yw, xr = oo.spatial_transformer_grid(d["feat_l"], d["feat_r"], d["disp_real"])
    close(yw, g["stn_real"], 2e-6, "stn_real")
    assert abs(xr.double().abs().sum().item() - g["stn_xrep_chk"]) < 1e-6
    close(oo.spatial_transformer_grid(d["feat_l"], d["feat_r"], d["disp_int"])[0], g["stn_int"], 2e-6, "stn_int")
    close(oo.regression_topk(d["cost24"], d["samples24"], 2), g["topk2"], 1e-6, "topk2")
    close(oo.regression_topk(d["cost24"], d["samples24"], 3), g["topk3"], 1e-6, "topk3")
    p = make_params(seed=2)
    close(oo.ssr_upsample(d["depth_low"], d["spx"], d["label"], p), g["ssr"], 2e-5, "ssr")
    if not s:
        close(oo.context_upsample(d["depth_low"], d["up9"]), g["context_up"], 1e-6, "context_up")
    for tag, block in (("444", (4, 4, 4)), ("644", (6, 4, 4))):
        close(oo.window_attention3d(d["att_in_" + tag], p, "hourglass.attention_block", 16, block),
              g["att_" + tag], 2e-5, "att_" + tag)
    if s:
        hg = oo.hourglass(d["hg_in"], p, "hourglass_att", (4, 4, 4))
        close(hg.reshape(-1)[::7], g["hourglass_sub"], 2e-4, "hourglass")


def test_interpolation_restatements_match_torch():
    g = torch.Generator().manual_seed(0)
    v = torch.randn(1, 1, 16, 8, 12, generator=g)
    ref = torch.nn.functional.interpolate(v, [32, 16, 24], mode="trilinear")
    assert (oo.trilinear_upsample(v, (32, 16, 24)) - ref).abs().max() < 1e-6
    t = torch.randn(2, 1, 7, 9, generator=g)
    ref = torch.nn.functional.interpolate(t, (28, 36), mode="bilinear")
    assert (oo.bilinear_upsample(t, (28, 36)) - ref).abs().max() < 1e-6


CASES = [("us3d_peaked", 64, True, 20.0, False), ("us3d_flat", 64, True, 1.0, False),
         ("us3d_attonly", 64, True, 20.0, True), ("whu_peaked", 128, False, 20.0, False),
         ("whu_attonly", 128, False, 20.0, True)]


@pytest.mark.parametrize("name,maxdisp,signed,peaked,att_only", CASES)
def test_hotpath_matches_reference_forward(golden_dir, name, maxdisp, signed, peaked, att_only):
    """Whole path vs the reference model's own forward (SemStereo.py / SemStereo_WHU.py:273-324)."""
    g = load(golden_dir, name)
    H, W, seed = int(g["meta"][3]), int(g["meta"][4]), int(g["meta"][5])
    p = make_params(seed=1, peaked=peaked)
    inp = make_inputs(seed, 1, H, W)
    for k, v in inp.items():
        assert abs(v.double().abs().sum().item() - g["chk_" + k]) <= 1e-9 * g["chk_" + k], f"RNG drift in {k}"
    chk_p = sum(v.double().abs().sum().item() for v in p.values())          # the weights too (goldens regenerated in round 2)
    assert abs(chk_p - float(g["chk_params"])) <= 1e-9 * float(g["chk_params"]), "RNG drift in the seeded parameters"
    out = oh.forward(p, inp, maxdisp, signed=signed, att_weights_only=att_only, keep=True)
    close(out["corr_volume"].reshape(-1)[::5], g["corr_volume_sub"], 1e-6, "corr_volume")
    close(out["cost_att"], g["cost_att"], 5e-4 * peaked, "cost_att")
    # top-k selection: exact wherever the reference probabilities are untied at the k boundary
    ref_ind = torch.from_numpy(g["ind_k"].astype(np.int64))
    prob = torch.from_numpy(g["prob"])
    srt = prob.sort(2, descending=True)[0]
    gap = (srt[:, :, 23] - srt[:, :, 24]).abs()                      # (B,1,H,W)
    untied = gap > 1e-6 * srt[:, :, 0]
    same = (out["ind_k"] == ref_ind).all(dim=2)
    assert bool(same[untied].all()), "top-k indices differ on untied pixels"
    frac_untied = untied.float().mean().item()
    assert frac_untied > (0.95 if peaked > 1 else 0.5)
    m = (same & untied).squeeze(1)                                     # (B,H,W) pixels to compare downstream
    close(out["att_topk"][:, 0].permute(0, 2, 3, 1)[m], g["att_topk"][:, 0].transpose(0, 2, 3, 1)[m.numpy()], 1e-5, "att_topk")
    close(out["pred_att"][m], g["pred_att"][:, 0][m.numpy()], 1e-3, "pred_att")
    if m.all():
        close(out["pred_att_up"], g["pred_att_up"], 1e-3, "pred_att_up")
    if not att_only:
        if m.all():
            close(out["volume"].reshape(-1)[::37], g["volume_sub"], 1e-5, "volume")
            close(out["cost"], g["cost"], 2e-3 * peaked, "cost")
            close(out["pred_up"], g["pred_up"], 1e-3, "pred_up")
            close(out["pred_up"] * 4, g["model_out"], 4e-3, "model_out")


def test_decoder_oracle_matches_reference_modules(golden_dir):
    """oracle/decoder.py (+ oracle/hotpath.py after it) vs the reference's own FeatUp / segmenthead / chal_* / spx* modules and
    its final disparity (tests/golden/decoder_us3d.npz, written by oracle/make_golden_decoder.py from /root/reference)."""
    from oracle import decoder as od
    from semstereo_b200.params import make_backbone_features, make_decoder_params
    g = np.load(os.path.join(golden_dir, "decoder_us3d.npz"))
    p = dict(make_params(seed=1, peaked=20.0))
    p.update(make_decoder_params(seed=2))
    fl, fr = make_backbone_features(7, 1, 128, 128)
    out = od.forward(p, fl, fr)
    for k in ("f4_l", "f4_r", "f8_l", "f8_r", "spx_pred", "pred_label", "pred_label_r"):
        ref = torch.from_numpy(g[k])
        assert (out[k] - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item()), k
    up = od.feat_up(p, fl)
    for k, i, st in (("x2_up", 0, 8), ("x4_up", 1, 8), ("x16_up", 3, 16)):
        ref = torch.from_numpy(g[k])
        assert (up[i][:, ::st] - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item()), k
    full = oh.forward(p, {k: out[k] for k in ("f8_l", "f8_r", "f4_l", "f4_r", "spx_pred", "pred_label")}, 64, signed=True)
    diff = (full["pred_up"] * 4 - torch.from_numpy(g["model_disp"])).abs()
    assert diff.median().item() <= 1e-3 and (diff > 1e-2).float().mean().item() <= 0.02     # top-k ties may flip isolated pixels


def test_attention_block_padded_and_masked_windows(golden_dir):
    """attention_block on H / W that are not multiples of the window (submodule_other.py:809-812, 822-829, 835-836): pad on one side
    only (the reference's `-0:` slice makes the mask all ones: nothing is masked) and on both (masked), both window shapes;
    tests/golden/att_padded.npz holds outputs of the unmodified reference module (oracle/make_golden_attpad.py)."""
    from oracle.make_golden_attpad import CASES
    g = dict(np.load(os.path.join(golden_dir, "att_padded.npz")))
    p = make_params(seed=2)
    for name, (block, shape) in CASES.items():
        x = torch.from_numpy(g["in_" + name])
        assert tuple(x.shape) == shape
        close(oo.window_attention3d(x, p, "hourglass.attention_block", 16, block), g["out_" + name], 2e-5, name)
