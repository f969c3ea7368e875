"""Drop-in for the two names the reference models take from `models/submodule_other.py` (SemStereo.py:8 star-imports that file
AFTER models.submodule, so its `attention_block` (:790-837) and `convbn_3d` (:845-848) are the ones the hourglasses use).
Installing this module as `models.submodule_other` next to `semstereo_b200.submodule` as `models.submodule` makes every 3-D
block of the unmodified SemStereo.py / SemStereo_WHU.py run on the B200 kernels (level 1).  The rest of that
reference file (RAFT-style encoders, GRUs, ...) is dead code for these models and is not mirrored."""
from .surface import make_surface as _make

_ns = _make(signed=True)          # both names are independent of the disparity convention
attention_block = _ns["attention_block"]
convbn_3d = _ns["convbn_3d"]
__all__ = ["attention_block", "convbn_3d"]
