"""Runs the default-precision hot path once (after one warm-up pass) at batch B on cuda:0: the command to wrap in ncu.
    ncu --set full --clock-control none --import-source on -k regex:<names> -s <skip> -c <n> -o gpurun_out/x python tools/ncu_path_once.py 8"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from semstereo_b200.hotpath import DisparityHotPath  # noqa: E402
from semstereo_b200.params import make_inputs, make_params  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
prec = sys.argv[2] if len(sys.argv) > 2 else "split"
passes = int(sys.argv[3]) if len(sys.argv) > 3 else 2
m = DisparityHotPath(64, False, True, precision=prec)
m.load_state_dict(make_params(seed=1, peaked=20.0), strict=True)
m = m.to("cuda:0")
inp = {k: v.to("cuda:0") for k, v in make_inputs(100, B, 1024, 1024).items() if k not in ("cf_l", "cf_r")}
order = ("f8_l", "f8_r", "f4_l", "f4_r", "cf_l", "cf_r", "spx_pred", "pred_label")
for _ in range(passes):
    out = m(*[inp.get(k) for k in order])
torch.cuda.synchronize()
print("ok", float(out["pred_up"].mean()))
