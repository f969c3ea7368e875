"""Multi-GPU check of dist.TiledHotPath over NCCL: torchrun --nproc-per-node N tools/check_tiles_nccl.py
One 2048x256 pair, N row bands (one per rank) with 384-row halos, stitched with an all-reduce; rank 0 compares with the untiled run."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from semstereo_b200 import dist as sd
from semstereo_b200.hotpath import DisparityHotPath
from semstereo_b200.params import make_inputs, make_params

rank, local, world = sd.init_from_env("nccl")
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
m = DisparityHotPath(64, False, True, precision="fp32")
m.load_state_dict(make_params(seed=1, peaked=20.0), strict=True)
m = m.to(dev)
inp = {k: v.to(dev) for k, v in make_inputs(31, 1, 2048, 256).items()}
tiled = sd.TiledHotPath(m, n_tiles=max(world, 2), halo=384)(inp)
torch.cuda.synchronize()
if rank == 0:
    full = m(*[inp[k] for k in ("f8_l", "f8_r", "f4_l", "f4_r", "cf_l", "cf_r", "spx_pred", "pred_label")])["pred_up"]
    d = (tiled - full).abs()
    print(f"[tiles over {world} ranks] max |diff| {float(d.max()):.3e}, pixels > 1e-3: {float((d > 1e-3).float().mean()):.2e}")
    assert float((d > 1e-3).float().mean()) <= 1e-3
if world > 1:
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()
