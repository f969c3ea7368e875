"""Runs the whole model (MobileViTv2 backbone + decoder + path, SemStereoB200) once after one warm-up pass at batch B on cuda:0:
the command to wrap in ncu for a launch list of the --stage full configuration.
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/x.csv python tools/ncu_full_once.py 2"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from semstereo_b200.backbone import SemStereoB200  # noqa: E402
from semstereo_b200.params import make_backbone_params, make_decoder_params, make_images, make_params  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
model = SemStereoB200(64, False, True)
sd = dict(make_params(seed=1, peaked=20.0))
sd.update(make_decoder_params(seed=2))
sd.update({"feature." + k: v for k, v in make_backbone_params(seed=4).items()})
model.load_state_dict(sd, strict=True)
model = model.to("cuda:0")
left, right = make_images(100, B, 1024, 1024)
left, right = left.to("cuda:0"), right.to("cuda:0")
for _ in range(2):
    out = model(left, right)
torch.cuda.synchronize()
print("ok", float(out["pred_up"].mean()))
