import sys; sys.path.insert(0, '/root/repo')
import torch
from semstereo_b200.params import make_inputs, make_params
from semstereo_b200.hotpath import DisparityHotPath
from semstereo_b200.pipeline import HostPipeline
DEV='cuda:0'
ORDER = ("f8_l", "f8_r", "f4_l", "f4_r", "cf_l", "cf_r", "spx_pred", "pred_label")
for prec in ('fp32','bf16'):
    m = DisparityHotPath(64, False, True, precision=prec); m.load_state_dict(make_params(seed=1, peaked=20.0), strict=True); m = m.to(DEV)
    batches = [{k: v.pin_memory() for k, v in make_inputs(30 + i, 1, 128, 128).items()} for i in range(5)]
    direct = [m(*[b[k].to(DEV) for k in ORDER])["pred_up"].cpu() for b in batches]
    direct2 = [m(*[b[k].to(DEV) for k in ORDER])["pred_up"].cpu() for b in batches]
    print(prec, 'direct repeatable', [torch.equal(a,b) for a,b in zip(direct,direct2)])
    pipe = HostPipeline(m, depth=2)
    got = [o.clone() for o in pipe.run(batches)]
    print(prec, 'pipe vs direct', [(a-b).abs().max().item() for a,b in zip(got,direct)])
    got2 = [o.clone() for o in pipe.run(batches)]
    print(prec, 'pipe2 vs direct', [(a-b).abs().max().item() for a,b in zip(got2,direct)])
