"""oracle/make_golden_decoder.py — TEST INFRASTRUCTURE ONLY.

Pins oracle/decoder.py (and decoder + hot path together) against the reference's OWN modules: builds models/SemStereo.py's
SemStereo from /root/reference (timm stubbed, appendix D of SURVEY.md), loads the seeded parameters of
semstereo_b200.params (hot path + decoder) into it, replaces only `model.feature` (the timm backbone) by a
replay of seeded backbone pyramids, runs the reference forward and records what FeatUp / heads / chal_* / spx* produce and the
final disparity.  Writes tests/golden/decoder_us3d.npz (sub-sampled to stay small).  Run here only: the GPU box has no
/root/reference.

    python -m oracle.make_golden_decoder
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

from oracle.make_golden import REF, Replay, stub_timm
from semstereo_b200.params import make_backbone_features, make_decoder_params, make_params

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "decoder_us3d.npz")
H = W = 128
SEED = 7


def main():
    stub_timm()
    sys.path.insert(0, REF)
    from models.SemStereo import SemStereo  # noqa
    torch.manual_seed(0)
    model = SemStereo(64, False, True, True, 6).eval()
    p = dict(make_params(seed=1, peaked=20.0))
    p.update(make_decoder_params(seed=2))
    missing, unexpected = model.load_state_dict(p, strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith("feature.") or k.endswith("num_batches_tracked") or "one_hot_filter" in k for k in missing), missing
    fl, fr = make_backbone_features(SEED, 1, H, W)
    model.feature = Replay([list(fl), list(fr)])
    cap = {}

    def keep(name):
        def hook(m, i, o):
            cap.setdefault(name, []).append(o.detach().clone() if torch.is_tensor(o) else [t.detach().clone() for t in o[0]])
        return hook

    model.head_l.register_forward_hook(keep("pred_label"))
    model.head_r.register_forward_hook(keep("pred_label_r"))
    model.chal_1.register_forward_hook(keep("chal_1"))
    model.chal_2.register_forward_hook(keep("chal_2"))
    model.spx2.register_forward_hook(keep("spx_pred"))
    model.feature_up.register_forward_hook(keep("feature_up"))
    with torch.no_grad():
        disp, label = model(torch.zeros(1, 3, H, W), torch.zeros(1, 3, H, W))
    g = {
        "x2_up": cap["feature_up"][0][0][:, ::8], "x4_up": cap["feature_up"][0][1][:, ::8], "x16_up": cap["feature_up"][0][3][:, ::16],
        "f4_l": cap["chal_1"][0], "f4_r": cap["chal_1"][1], "f8_l": cap["chal_2"][0], "f8_r": cap["chal_2"][1],
        "spx_pred": cap["spx_pred"][0], "pred_label": cap["pred_label"][0], "pred_label_r": cap["pred_label_r"][0],
        "model_disp": disp[0],
    }
    assert torch.equal(label, cap["pred_label"][0])
    np.savez_compressed(OUT, **{k: v.numpy().astype(np.float32) for k, v in g.items()})
    print("wrote", OUT, {k: tuple(v.shape) for k, v in g.items()}, f"{os.path.getsize(OUT) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
