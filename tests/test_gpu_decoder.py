"""Decoder2D / StereoHead (bf16 tensor-core 2-D decoder + disparity path) vs the CPU oracle (oracle/decoder.py, pinned to the
reference's own modules by tests/golden/decoder_us3d.npz) and vs that golden directly.  bf16 operands, fp32 accumulation:
tolerances are statistical and stated here; the per-layer arithmetic is checked exactly in test_gpu_conv2d.py."""
import os

import numpy as np
import pytest
import torch

from oracle import decoder as od
from oracle import hotpath as oh
from semstereo_b200.params import make_backbone_features, make_decoder_params, make_params

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

if torch.cuda.is_available():
    from semstereo_b200.decoder import Decoder2D, StereoHead


def params():
    p = dict(make_params(seed=1, peaked=20.0))
    p.update(make_decoder_params(seed=2))
    return p


def rel(got, ref):
    d = (got.double() - ref.double()).abs()
    return d.max().item() / ref.abs().max().item(), d.mean().item() / ref.abs().mean().item()


@pytest.mark.parametrize("B,H,W", [(1, 128, 128), (2, 128, 256), (1, 192, 192)])
def test_decoder_against_oracle(B, H, W):
    p = params()
    fl, fr = make_backbone_features(7, B, H, W)
    ref = od.forward(p, fl, fr)
    m = Decoder2D()
    m.load_state_dict(p, strict=True)
    m = m.to(DEV)
    out = m([t.to(DEV) for t in fl], [t.to(DEV) for t in fr], right_label=True)
    torch.cuda.synchronize()
    for k in ("f4_l", "f4_r", "f8_l", "f8_r", "spx_pred", "pred_label", "pred_label_r"):
        got = out[k].cpu()
        assert got.shape == ref[k].shape, k
        rmax, rmean = rel(got, ref[k])
        print(f"\n[decoder bf16] {k}: max err / max|ref| = {rmax:.4f}, mean err / mean|ref| = {rmean:.4f}")
        assert rmax <= 1e-2 and rmean <= 1e-2, (k, rmax, rmean)          # measured 0.4-0.5 % (bf16 operands through <= 10 layers)
    assert torch.equal(tc_from_blocked(out["f4_l_blocked"]), out["f4_l"].cpu())


def tc_from_blocked(xb):
    from semstereo_b200 import ops_tc as tc
    return tc.from_blocked2d(xb).cpu()


def test_decoder_against_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "decoder_us3d.npz"))
    p = params()
    fl, fr = make_backbone_features(7, 1, 128, 128)
    m = Decoder2D()
    m.load_state_dict(p, strict=True)
    out = m.to(DEV)([t.to(DEV) for t in fl], [t.to(DEV) for t in fr], right_label=True)
    for k in ("f4_l", "f8_r", "spx_pred", "pred_label", "pred_label_r"):
        rmax, rmean = rel(out[k].cpu(), torch.from_numpy(g[k]))
        assert rmax <= 1e-2 and rmean <= 1e-2, (k, rmax, rmean)          # measured 0.4-0.5 %


def test_stereo_head_against_oracle():
    """Backbone pyramids -> full-resolution disparity: decoder + path on the GPU vs oracle decoder + oracle path."""
    p = params()
    fl, fr = make_backbone_features(7, 1, 128, 256)
    d = od.forward(p, fl, fr)
    ref = oh.forward(p, {k: d[k] for k in ("f8_l", "f8_r", "f4_l", "f4_r", "spx_pred", "pred_label")}, 64, signed=True, keep=True)
    m = StereoHead(64)
    m.load_state_dict(p, strict=True)
    assert set(m.state_dict()) >= set(p)
    out = m.to(DEV)([t.to(DEV) for t in fl], [t.to(DEV) for t in fr], keep=True)
    torch.cuda.synchronize()
    agree = (out["ind_k"].cpu() == ref["ind_k"]).all(dim=2).float().mean().item()
    e = (out["pred_up"].cpu() - ref["pred_up"]).abs().flatten()
    print(f"\n[stereo head bf16] top-24 set agreement {agree:.4f}; pred_up median {e.median():.4f} p90 {e.quantile(0.9):.4f} "
          f"max {e.max():.3f} (1/4-res px)")
    assert agree >= 0.85
    assert e.median().item() <= 0.06 and e.quantile(0.9).item() <= 0.8
    disp, label = m.as_model_outputs(out)
    assert tuple(disp[0].shape) == (1, 128, 256) and tuple(label.shape) == (1, 6, 128, 256)


def test_patch_model_rebinds_forward():
    """patch_model on a stand-in with the reference model's attribute surface (the reference tree is not on the GPU box):
    backbone = its own torch module, everything after = StereoHead with the parameters read from model.state_dict()."""
    import torch.nn as nn
    from semstereo_b200.patch import patch_model
    from semstereo_b200.params import BACKBONE_CHANS

    class Backbone(nn.Module):
        def __init__(self):
            super().__init__()
            self.convs = nn.ModuleList(nn.Conv2d(3, c, 1) for c in BACKBONE_CHANS)

        def forward(self, x):
            return [conv(torch.nn.functional.avg_pool2d(x, s)) for conv, s in zip(self.convs, (2, 4, 8, 16, 32))]

    class StandIn(nn.Module):
        def __init__(self, p):
            super().__init__()
            self.maxdisp, self.att_weights_only, self.seg_if, self.stereo_if, self.num_classes = 64, False, True, True, 6
            self.feature = Backbone()
            self.extra = nn.ParameterDict()
            self._p = p

        def state_dict(self, *a, **k):
            sd = super().state_dict(*a, **k)
            sd.update(self._p)
            return sd

    torch.manual_seed(0)
    p = params()
    model = StandIn(p).to(DEV).eval()
    left, right = torch.randn(1, 3, 128, 128, device=DEV), torch.randn(1, 3, 128, 128, device=DEV)
    patch_model(model)
    disp, label = model(left, right)
    head = StereoHead(64)
    head.load_state_dict(p, strict=True)
    head = head.to(DEV)
    with torch.no_grad():
        out = head(model.feature(left), model.feature(right))
    assert torch.equal(disp[0], out["pred_up"] * 4) and torch.equal(label, out["pred_label"])
    model.train()
    with pytest.raises(NotImplementedError):
        model(left, right)


def test_stereo_head_full_size_properties_and_graph_replay():
    """BASELINE config #1 shape (1,1024,1024) through decoder + path: shapes, finiteness, label/disparity ranges that need no
    oracle run, and a CUDA-graph replay of the whole thing equal to the eager result bit for bit."""
    from semstereo_b200.graph import GraphedCall
    p = params()
    head = StereoHead(64)
    head.load_state_dict(p, strict=True)
    head = head.to(DEV)
    fl, fr = make_backbone_features(5, 1, 1024, 1024)
    st = {f"l{i}": t.to(DEV) for i, t in enumerate(fl)}
    st.update({f"r{i}": t.to(DEV) for i, t in enumerate(fr)})
    call = lambda s: head([s[f"l{i}"] for i in range(5)], [s[f"r{i}"] for i in range(5)])       # noqa: E731
    out = call(st)
    torch.cuda.synchronize()
    assert tuple(out["pred_up"].shape) == (1, 1024, 1024) and tuple(out["pred_label"].shape) == (1, 6, 1024, 1024)
    for k in ("pred_up", "pred_att_up", "pred_label", "disp_topk", "att_topk"):
        assert bool(torch.isfinite(out[k]).all()), k
    assert float(out["disp_topk"].min()) >= -16 and float(out["disp_topk"].max()) <= 15
    assert bool((out["disp_topk"][:, 1:] > out["disp_topk"][:, :-1]).all()), "kept bins must be strictly ascending"
    assert float(out["att_topk"].min()) >= 0 and float(out["att_topk"].sum(2 if out["att_topk"].dim() == 5 else 1).max()) <= 1 + 1e-5
    want = {k: out[k].clone() for k in ("pred_up", "pred_label")}
    g = GraphedCall(call, st)
    got = g.replay()
    torch.cuda.synchronize()
    assert torch.equal(got["pred_up"], want["pred_up"]) and torch.equal(got["pred_label"], want["pred_label"])


def test_loss_inputs_tuple_for_lrsc():
    """BASELINE config #4: WHU model (unsigned), forward outputs incl. pred_label_r in the layout the reference's training forward
    returns (SemStereo_WHU.py:329-337), so LRSC_loss / model_loss_train can be evaluated on top."""
    p = params()
    head = StereoHead(128, False, signed=False)
    head.load_state_dict(p, strict=True)
    fl, fr = make_backbone_features(9, 2, 128, 256)
    out = head.to(DEV)([t.to(DEV) for t in fl], [t.to(DEV) for t in fr], right_label=True)
    disp, label, label_r = head.as_loss_inputs(out)
    assert [tuple(t.shape) for t in disp] == [(2, 128, 256), (2, 32, 64), (2, 128, 256), (2, 32, 64)]
    assert tuple(label.shape) == tuple(label_r.shape) == (2, 6, 128, 256)
    assert float(disp[1].min()) >= 0 and float(disp[1].max()) <= 4 * 31          # unsigned: regression over bins 0..31, x4
    with pytest.raises(ValueError):
        head.as_loss_inputs(head([t.to(DEV) for t in fl], [t.to(DEV) for t in fr]))


def test_stereo_head_full_size_against_oracle():
    """Decoder + path at the headline size (1, 1024, 1024): the oracle (decoder + path, ~3 s on the host cores) is compared directly."""
    p = params()
    fl, fr = make_backbone_features(5, 1, 1024, 1024)
    d = od.forward(p, fl, fr, right_label=False)
    ref = oh.forward(p, {k: d[k] for k in ("f8_l", "f8_r", "f4_l", "f4_r", "spx_pred", "pred_label")}, 64, signed=True, keep=True)
    head = StereoHead(64)
    head.load_state_dict(p, strict=True)
    out = head.to(DEV)([t.to(DEV) for t in fl], [t.to(DEV) for t in fr], keep=True)
    torch.cuda.synchronize()
    for k in ("pred_label", "spx_pred", "f4_l", "f8_r"):
        rmax, rmean = rel(out[k].cpu(), d[k])
        assert rmax <= 1e-2 and rmean <= 1e-2, (k, rmax, rmean)          # measured 0.4-0.5 %
    agree = (out["ind_k"].cpu() == ref["ind_k"]).all(dim=2).float().mean().item()
    e = (out["pred_up"].cpu() - ref["pred_up"]).abs().flatten()
    print(f"\n[stereo head 1024x1024] top-24 agreement {agree:.4f}; pred_up median {e.median():.4f} p90 {e.quantile(0.9):.4f} (1/4-res px)")
    assert agree >= 0.85 and e.median().item() <= 0.06 and e.quantile(0.9).item() <= 0.8


# 192 x 192: a multiple of 64 that is not one of 128 -- the attention block of hourglass_att pads and masks its windows (H/32 = W/32 = 6)
@pytest.mark.parametrize("B,H,W", [(1, 128, 256), (1, 192, 192)])
def test_split_precision_decoder_and_head_select_the_oracle_samples(B, H, W):
    """precision="split": FeatUp + chal_1/2 with fp32-accurate bf16x3 products (K-concat GEMMs) -> the path's inputs match the
    oracle decoder to ~1e-5, and with the split attention branch behind it the top-24 sample sets equal the oracle's END TO END
    from the backbone features (StereoHead precision)."""
    p = params()
    fl, fr = make_backbone_features(11, B, H, W)
    d = od.forward(p, fl, fr, right_label=False)
    ref = oh.forward(p, {k: d[k] for k in ("f8_l", "f8_r", "f4_l", "f4_r", "spx_pred", "pred_label")}, 64, signed=True, keep=True)
    head = StereoHead(64, precision="split")
    head.load_state_dict(p, strict=True)
    out = head.to(DEV)([t.to(DEV) for t in fl], [t.to(DEV) for t in fr], keep=True)
    torch.cuda.synchronize()
    for k in ("f4_l", "f4_r", "f8_l", "f8_r"):
        rmax, _ = rel(out[k].cpu(), d[k])
        print(f"[decoder split] {k}: max err / max|ref| = {rmax:.2e}")
        assert rmax <= 5e-5, (k, rmax)
    agree = (out["ind_k"].cpu() == ref["ind_k"]).all(dim=2).float().mean().item()
    med = (out["pred_up"].cpu() - ref["pred_up"]).abs().median().item()
    print(f"[stereo head split] top-24 sample-set agreement {agree:.6f}; median |pred_up - oracle| {med:.4f} px (1/4-res units)")
    assert agree >= 0.999 and med <= 0.05
    with pytest.raises(ValueError):
        StereoHead(64, precision="fp32")
