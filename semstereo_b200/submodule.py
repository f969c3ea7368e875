"""Signed operator surface: drop-in for the reference's `models/submodule.py` (+ `attention_block` / `convbn_3d` of
`models/submodule_other.py:790-848`): disparities -maxdisp..maxdisp-1, volume depth 2*maxdisp.
`from semstereo_b200.submodule import *` gives SemStereo.py the names it star-imports (SemStereo.py:7-8)."""
from .surface import make_surface as _make

globals().update(_make(signed=True))
__all__ = [k for k in _make(signed=True)]
