"""oracle/ — TEST INFRASTRUCTURE ONLY.  Not part of the product.

CPU (torch fp32) restatement of the disparity hot path of chenchen235/SemStereo:
cost-volume construction -> Conv3d hourglass aggregation -> disparity regression
(reference `models/submodule.py`, `models/submodule_.py`, `models/submodule_other.py:790-848`,
`models/SemStereo.py:89-182,241-244,273-324`).

Who may import this package: `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` -- as the CHECKER or as the
timed CPU baseline, never as the thing shipped.  `semstereo_b200/` must not import it
(tests/test_layout.py enforces that) and fails loudly when its CUDA library is missing.

Pinning status: the reference ships NO tests, golden vectors or known answers
(SURVEY.md section 4), so the reference's own tests pin nothing.  The oracle is pinned
instead against outputs of the UNMODIFIED reference code executed in the authoring
container (`oracle/make_golden.py`, which imports `/root/reference` by path); the
resulting vectors are committed under `tests/golden/` and `tests/test_oracle_golden.py`
replays them on every run.  `/root/reference` is never read at test/bench run time.

The arithmetic of the reference lives in PyTorch (README pins pytorch==1.12.1; this
image has torch 2.11.0).  Convolution / batch-norm / softmax / sort use the same ATen CPU
ops here; grid_sample, bilinear/trilinear interpolation, unfold+nearest, replication-pad
one-hot convolutions and the volume builders are restated in closed form.
"""
