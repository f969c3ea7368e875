"""oracle/backbone.py -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference's `Feature` (models/SemStereo.py:33-56) is timm's `mobilevitv2_100` (third-party, README pins timm==0.6.5; not
installed here and not under /root/reference).  Its published architecture, MobileViTv2-1.0, is restated by HuggingFace
`transformers` (`MobileViTV2Model`, v5.5.0 in this image; width_multiplier 1.0 yields exactly the 64/128/256/384/512-channel
pyramid at strides 2/4/8/16/32 that `FeatUp` expects).  That implementation, run on the CPU in fp32, is
the oracle of semstereo_b200/backbone.py: same parameter names, so one state_dict feeds both.  Parity is therefore pinned to HF's
restatement of the architecture, not to timm's code (unavailable): stated in DESIGN.md.
"""
from __future__ import annotations

import torch


def build(state_dict=None):
    from transformers import MobileViTV2Config, MobileViTV2Model
    m = MobileViTV2Model(MobileViTV2Config(width_multiplier=1.0), expand_output=False).eval()
    if state_dict is not None:
        missing, unexpected = m.load_state_dict(state_dict, strict=False)
        bad = [k for k in missing if not k.endswith("num_batches_tracked")]
        assert not bad and not unexpected, (bad[:3], unexpected[:3])
    return m


@torch.no_grad()
def forward(state_dict, image):
    """image fp32 (B,3,H,W) -> [x2, x4, x8, x16, x32] fp32 NCHW: the five stage outputs (`Feature.forward`, SemStereo.py:47-56)."""
    m = build(state_dict)
    out = m(image, output_hidden_states=True)
    return list(out.hidden_states)
