"""Unsigned operator surface: drop-in for the reference's `models/submodule_.py` (the flavour SemStereo_WHU.py is
written against): disparities 0..maxdisp-1, volume depth maxdisp; adds `context_upsample`."""
from .surface import make_surface as _make

globals().update(_make(signed=False))
__all__ = [k for k in _make(signed=False)]
