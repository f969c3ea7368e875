"""Level-1 drop-in (north_star: "SemStereo.py and SemStereo_WHU.py call it as a drop-in"): the UNMODIFIED reference model files are
imported with `semstereo_b200.submodule` / `submodule_` installed as `models.submodule` and `semstereo_b200.submodule_other` as
`models.submodule_other` (timm stubbed: it is not installed here), the models are constructed, and their
state_dict is checked against the parameter inventory and loaded.  Runs wherever the reference tree is available
(SEMSTEREO_REFERENCE, default /root/reference; absent on the GPU box -> skipped there).  The forward of that glue on the GPU is
covered by tests/test_gpu_surface_glue.py, which restates the glue so that it can travel."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("SEMSTEREO_REFERENCE", "/root/reference")

SCRIPT = textwrap.dedent("""
    import importlib, sys, types
    import torch, torch.nn as nn
    sys.path.insert(0, {root!r})
    ref, variant = {ref!r}, {variant!r}

    # timm stub: what Feature.__init__ touches (models/SemStereo.py:37-45)
    timm = types.ModuleType("timm")
    def create_model(name, pretrained=False, features_only=False, **kw):
        m = nn.Module()
        def stage(ci, co, s):
            return nn.Sequential(nn.Conv2d(ci, co, 3, s, 1), nn.BatchNorm2d(co), nn.SiLU())
        m.stem = stage(3, 32, 2)
        for i, (ci, co, s) in enumerate(((32, 64, 1), (64, 128, 2), (128, 256, 2), (256, 384, 2), (384, 512, 2))):
            setattr(m, f"stages_{{i}}", nn.Sequential(stage(ci, co, s)))
        return m
    timm.create_model = create_model
    sys.modules["timm"] = timm

    pkg = types.ModuleType("models"); pkg.__path__ = [ref + "/models"]; sys.modules["models"] = pkg     # skip models/__init__.py
    import semstereo_b200.submodule as signed, semstereo_b200.submodule_ as unsigned, semstereo_b200.submodule_other as other
    sys.modules["models.submodule"] = signed if variant == "us3d" else unsigned
    sys.modules["models.submodule_other"] = other
    mod = importlib.import_module("models.SemStereo" if variant == "us3d" else "models.SemStereo_WHU")
    cls = mod.SemStereo if variant == "us3d" else mod.SemStereo_WHU
    for att_only in (False, True):
        model = cls(64 if variant == "us3d" else 128, att_only, True, True, 6).eval()
    model = cls(64 if variant == "us3d" else 128, False, True, True, 6).eval()

    from semstereo_b200.params import make_params, make_decoder_params
    want = dict(make_params(seed=1, peaked=20.0)); want.update(make_decoder_params(seed=2))
    sd = {{k: v for k, v in model.state_dict().items() if not k.endswith("num_batches_tracked")}}
    ours = {{k for k in sd if not k.startswith("feature.")}}
    assert ours == set(want), (sorted(ours - set(want))[:5], sorted(set(want) - ours)[:5])
    assert all(tuple(sd[k].shape) == tuple(want[k].shape) for k in want)
    missing, unexpected = model.load_state_dict(want, strict=False)
    assert not unexpected and all(k.startswith("feature.") for k in missing)
    assert torch.equal(model.state_dict()["hourglass.conv5.0.weight"], want["hourglass.conv5.0.weight"])

    # the 3-D blocks and the stateful operators are the B200 modules, not torch re-implementations
    b200 = lambda m: type(m).__module__.startswith("semstereo_b200")
    assert b200(model.hourglass_att.conv1[0]) and b200(model.hourglass.attention_block) and b200(model.concat_stem)
    assert b200(model.classif[0]) and b200(model.ssr_upsample) and b200(model.propagation) and b200(model.propagation_prob)
    assert b200(model.feature_up.deconv32_16) and b200(model.head_l)
    for name in ("build_gwc_volume_norm", "disparity_regression", "disparity_variance", "SpatialTransformer_grid", "regression_topk"):
        assert getattr(mod, name).__module__.startswith("semstereo_b200"), name
    # no CUDA here: the first kernel-backed operator refuses CPU tensors instead of falling back
    if not torch.cuda.is_available():
        x = torch.randn(1, 3, 128, 128)
        try:
            model(x, x)
        except RuntimeError as e:
            assert "no CPU fallback" in str(e), e
        else:
            raise AssertionError("the drop-in ran on CPU tensors")
    print("DROPIN-OK", variant, len(sd))
""")


@pytest.mark.parametrize("variant", ["us3d", "whu"])
def test_reference_models_construct_against_the_shim(variant):
    if not os.path.isdir(os.path.join(REF, "models")):
        pytest.skip(f"reference tree not available at {REF}")
    r = subprocess.run([sys.executable, "-c", SCRIPT.format(root=ROOT, ref=REF, variant=variant)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DROPIN-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
