"""oracle/make_golden.py — TEST INFRASTRUCTURE ONLY.

Generates the fixtures under tests/golden/ by running the UNMODIFIED reference code
from /root/reference (imported by path, never copied) on seeded synthetic inputs.
Run in the authoring container only:   python -m oracle.make_golden
(spawns one subprocess per reference flavour: the signed `models/submodule.py` and the
unsigned `models/submodule_.py` share module names).

Inputs and weights are NOT stored: they are regenerated from seeds by
`semstereo_b200.params.make_params / make_inputs` and the small generators below; each
fixture stores a float64 checksum of every input so RNG drift is detected, plus the
reference outputs (sub-sampled where large).
"""
from __future__ import annotations

import argparse
import importlib
import importlib.util
import os
import subprocess
import sys
import types

import numpy as np
import torch
import torch.nn as nn

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from semstereo_b200.params import make_inputs, make_params  # noqa: E402


# ---------------------------------------------------------------------------------------
# seeded small-op inputs shared with tests/test_oracle_golden.py
# ---------------------------------------------------------------------------------------
def op_inputs():
    g = torch.Generator().manual_seed(7)
    r = lambda *s: torch.randn(*s, generator=g)
    d = {}
    d["ref"], d["tgt"] = r(2, 16, 5, 24), r(2, 16, 5, 24)
    d["prob32"] = torch.softmax(r(2, 8, 6, 10), 1)
    d["mu"] = r(2, 1, 6, 10)
    d["disp1"] = 3 * r(2, 1, 9, 11)
    d["vol1"] = r(2, 1, 6, 9, 11)
    d["feat_l"], d["feat_r"] = r(2, 8, 9, 20), r(2, 8, 9, 20)
    d["disp_real"] = 4 * r(2, 5, 9, 20)
    d["disp_int"] = torch.randint(-6, 7, (2, 7, 9, 20), generator=g).float()
    d["cost24"], d["samples24"] = r(2, 24, 7, 9), torch.sort(torch.randint(-8, 8, (2, 24, 7, 9), generator=g).float(), 1)[0]
    d["depth_low"], d["up9"] = 5 * r(2, 1, 6, 7), torch.softmax(r(2, 9, 24, 28), 1)
    d["spx"], d["label"] = r(2, 6, 24, 28), 2 * r(2, 6, 24, 28)
    d["att_in_444"] = r(1, 128, 4, 8, 8)
    d["att_in_644"] = r(1, 128, 6, 8, 4)
    d["hg_in"] = r(1, 32, 16, 32, 32)
    return d


def checksums(d):
    return {"chk_" + k: np.float64(v.double().abs().sum().item()) for k, v in d.items() if torch.is_tensor(v)}


def sub(t, step):
    return t.detach().reshape(-1)[::step].contiguous().numpy()


# ---------------------------------------------------------------------------------------
# reference loading (appendix D of SURVEY.md)
# ---------------------------------------------------------------------------------------
def stub_timm():
    timm = types.ModuleType("timm")

    class _Backbone(nn.Module):
        def __init__(self):
            super().__init__()
            self.stem = nn.Conv2d(3, 32, 3, 2, 1)
            chans = [(32, 64, 1), (64, 128, 2), (128, 256, 2), (256, 384, 2), (384, 512, 2)]
            for i, (ci, co, s) in enumerate(chans):
                setattr(self, f"stages_{i}", nn.Sequential(nn.Conv2d(ci, co, 3, s, 1)))

    timm.create_model = lambda *a, **k: _Backbone()
    sys.modules["timm"] = timm


def load_file(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


class Replay(nn.Module):
    """Returns preset tensors in call order (stands in for the out-of-scope 2-D modules)."""

    def __init__(self, outs):
        super().__init__()
        self.outs, self.i = list(outs), 0

    def forward(self, *a, **k):
        o = self.outs[self.i % len(self.outs)]
        self.i += 1
        return o


def run_model(flavour, maxdisp, peaked, att_only, H=128, W=256, seed=3):
    """Execute the reference model's own forward with the 2-D front end replaced by Replay."""
    stub_timm()
    if flavour == "us3d":
        sys.path.insert(0, REF)
        from models.SemStereo import SemStereo as Model  # noqa
        modname = "models.SemStereo"
    else:
        pkg = types.ModuleType("models")
        pkg.__path__ = [REF + "/models"]
        sys.modules["models"] = pkg
        load_file("models.submodule", REF + "/models/submodule_.py")
        sys.path.insert(0, REF)
        Model = importlib.import_module("models.SemStereo_WHU").SemStereo_WHU
        modname = "models.SemStereo_WHU"
    torch.manual_seed(0)
    model = Model(maxdisp, att_only, True, True, 6).eval()
    p = make_params(seed=1, peaked=peaked)
    missing, unexpected = model.load_state_dict(p, strict=False)
    assert not unexpected, unexpected
    inp = make_inputs(seed, 1, H, W)
    B = 1
    z = lambda c, s: torch.zeros(B, c, H // s, W // s)
    model.feature = Replay([[z(1, 2), inp["f4_l"], inp["f8_l"], z(1, 16), z(1, 32)],
                            [z(1, 2), inp["f4_r"], inp["f8_r"], z(1, 16), z(1, 32)]])
    model.feature_up = Replay([None])
    model.feature_up.forward = lambda l, r: (l, r)
    model.head_l, model.head_r = Replay([inp["pred_label"]]), Replay([inp["pred_label"]])
    for i in range(5):
        setattr(model, f"chal_{i}", nn.Identity())
    for n in ("spx32_16", "spx16_8", "spx8_4", "spx4_2"):
        setattr(model, n, Replay([z(1, 2)]))
    model.spx2 = Replay([inp["spx_pred"]])
    model.concat_feature = Replay([inp["cf_l"], inp["cf_r"]])

    cap = {}
    ns = sys.modules[modname]
    orig_gwc = ns.build_gwc_volume_norm
    ns.build_gwc_volume_norm = lambda *a: cap.setdefault("corr_volume", orig_gwc(*a))
    gathers = []
    orig_gather = torch.gather

    def rec_gather(inp_, dim, index, **k):
        o = orig_gather(inp_, dim, index, **k)
        gathers.append((inp_, index, o))
        return o

    hooks = []
    def hk(name):
        def hook(m, i, o):          # must return None (a returned value would replace the output)
            cap.setdefault(name, (i, o))
        return hook
    for name in ("classif_att_", "classif", "ssr_upsample", "hourglass_att", "hourglass", "concat_stem"):
        if hasattr(model, name):
            hooks.append(getattr(model, name).register_forward_hook(hk(name)))
    ssr_calls = []
    def ssr_hook(m, i, o):
        ssr_calls.append((i[0], o))
    hooks.append(model.ssr_upsample.register_forward_hook(ssr_hook))
    torch.gather = rec_gather
    try:
        with torch.no_grad():
            outs = model(torch.zeros(B, 3, H, W), torch.zeros(B, 3, H, W))
    finally:
        torch.gather = orig_gather
        ns.build_gwc_volume_norm = orig_gwc
        for h in hooks:
            h.remove()
    res = {"chk_" + k: np.float64(v.double().abs().sum().item()) for k, v in inp.items()}
    res["chk_params"] = np.float64(sum(v.double().abs().sum().item() for v in p.values()))
    res["corr_volume_sub"] = sub(cap["corr_volume"], 5)
    res["cost_att"] = cap["classif_att_"][1].numpy()
    res["hourglass_att_sub"] = sub(cap["hourglass_att"][1], 11)
    res["ind_k"] = gathers[0][1].numpy().astype(np.uint8)
    res["att_topk"] = gathers[0][2].numpy()
    res["prob"] = gathers[0][0].numpy()
    res["pred_att"] = ssr_calls[0][0].numpy()
    res["pred_att_up"] = ssr_calls[0][1].numpy()
    if not att_only:
        res["volume_sub"] = sub(cap["concat_stem"][0][0], 37)
        res["concat_stem_sub"] = sub(cap["concat_stem"][1], 19)
        res["hourglass_sub"] = sub(cap["hourglass"][1], 19)
        res["cost"] = cap["classif"][1].numpy()
        res["pred"] = ssr_calls[1][0].numpy()
        res["pred_up"] = ssr_calls[1][1].numpy()
        res["model_out"] = outs[0][0].numpy()      # = pred_up * 4
    else:
        res["model_out"] = outs[0][0].numpy()      # = pred_att_up * 4
    res["meta"] = np.array([maxdisp, peaked, int(att_only), H, W, seed], dtype=np.float64)
    return res


def run_ops(flavour):
    """Op-level fixtures from models/submodule.py (signed) or models/submodule_.py (unsigned)."""
    path = REF + ("/models/submodule.py" if flavour == "signed" else "/models/submodule_.py")
    m = load_file("ref_sub_" + flavour, path)
    d = op_inputs()
    res = checksums(d)
    M, G = 4, 4
    res["gwc"] = m.build_gwc_volume(d["ref"], d["tgt"], M, G).numpy()
    res["gwc_norm"] = m.build_gwc_volume_norm(d["ref"], d["tgt"], M, G).numpy()
    res["concat"] = m.build_concat_volume(d["ref"], d["tgt"], M).numpy()
    res["normcorr"] = m.build_norm_correlation_volume(d["ref"], d["tgt"], M).numpy()
    nb = 4 if flavour == "signed" else 8
    res["regress"] = m.disparity_regression(d["prob32"], nb).numpy()
    res["variance"] = m.disparity_variance(d["prob32"], nb, d["mu"]).numpy()
    res["prop"] = m.Propagation()(d["disp1"]).numpy()
    res["prop_prob"] = m.Propagation_prob()(d["vol1"]).numpy()
    yw, xr = m.SpatialTransformer_grid(d["feat_l"], d["feat_r"], d["disp_real"])
    res["stn_real"], res["stn_xrep_chk"] = yw.numpy(), np.float64(xr.double().abs().sum().item())
    res["stn_int"] = m.SpatialTransformer_grid(d["feat_l"], d["feat_r"], d["disp_int"])[0].numpy()
    res["topk2"] = m.regression_topk(d["cost24"], d["samples24"], 2).numpy()
    res["topk3"] = m.regression_topk(d["cost24"], d["samples24"], 3).numpy()
    p = make_params(seed=2)
    ssr = m.SSR_upsample(6).eval()
    ssr.load_state_dict({k[len("ssr_upsample."):]: v for k, v in p.items() if k.startswith("ssr_upsample.")})
    with torch.no_grad():
        res["ssr"] = ssr(d["depth_low"], d["spx"], d["label"]).numpy()
    if flavour == "unsigned":
        res["context_up"] = m.context_upsample(d["depth_low"], d["up9"]).numpy()
        blk = m.attention_block          # submodule_.py:15-61 (copy of submodule_other.py:790-837)
    else:
        stub_timm()
        other = load_file("ref_other", REF + "/models/submodule_other.py")
        blk = other.attention_block
    for tag, block in (("444", (4, 4, 4)), ("644", (6, 4, 4))):
        a = blk(128, 16, block).eval()
        a.load_state_dict({k[len("hourglass.attention_block."):]: v for k, v in p.items()
                           if k.startswith("hourglass.attention_block.")})
        with torch.no_grad():
            res["att_" + tag] = a(d["att_in_" + tag]).numpy()
    if flavour == "signed":
        sys.path.insert(0, REF)
        from models.SemStereo import hourglass  # noqa
        hg = hourglass(32).eval()
        hg.load_state_dict({k[len("hourglass_att."):]: v for k, v in p.items() if k.startswith("hourglass_att.")})
        with torch.no_grad():
            res["hourglass_sub"] = sub(hg(d["hg_in"]), 7)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--job", default="all")
    a = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    if a.job == "all":
        for job in ("ops_signed", "ops_unsigned", "us3d_peaked", "us3d_flat", "us3d_attonly", "whu_peaked", "whu_attonly"):
            subprocess.check_call([sys.executable, "-m", "oracle.make_golden", "--job", job],
                                  cwd=os.path.dirname(OUT.rstrip("/")).rsplit("/tests", 1)[0])
        return
    if a.job.startswith("ops_"):
        res = run_ops(a.job[4:])
    else:
        flavour, kind = a.job.split("_")
        res = run_model(flavour, 64 if flavour == "us3d" else 128,
                        peaked=1.0 if kind == "flat" else 20.0, att_only=(kind == "attonly"))
    np.savez_compressed(os.path.join(OUT, a.job + ".npz"), **res)
    print(a.job, {k: (v.shape if hasattr(v, "shape") and v.shape else float(v)) for k, v in res.items()})


if __name__ == "__main__":
    main()
