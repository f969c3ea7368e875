"""Decoder2D / StereoHead on the BASELINE config #3 shape (B pairs of 1024x1024): per-layer CUDA-event times, TFLOP/s per conv
layer, whole-decoder and decoder+path pairs/s.  Usage: python tools/bench_decoder.py [B]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from semstereo_b200 import ops
from semstereo_b200.decoder import StereoHead
from semstereo_b200.params import BACKBONE_CHANS, make_decoder_params, make_params

dev = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
H = W = 1024
p = dict(make_params(seed=1, peaked=20.0))
p.update(make_decoder_params(seed=2))
m = StereoHead(64)
m.load_state_dict(p, strict=True)
m = m.to(dev)
g = torch.Generator(device=dev).manual_seed(0)
fl = [torch.randn(B, c, H // s, W // s, device=dev, generator=g) for c, s in zip(BACKBONE_CHANS, (2, 4, 8, 16, 32))]
fr = [torch.randn(B, c, H // s, W // s, device=dev, generator=g) for c, s in zip(BACKBONE_CHANS, (2, 4, 8, 16, 32))]


def flops(name):
    """2*MACs per image for the conv layers of the decoder (H=W=1024)."""
    px = {32: 32 * 32, 16: 64 * 64, 8: 128 * 128, 4: 256 * 256, 2: 512 * 512, 1: 1024 * 1024}
    t = {"deconv32_16": (512, 384, 16), "deconv16_8": (768, 256, 8), "deconv8_4": (512, 128, 4), "deconv4_2": (256, 64, 2),
         "spx32_16": (256, 384, 16), "spx16_8": (768, 256, 8), "spx8_4": (512, 128, 4), "spx4_2": (256, 64, 2)}
    for k, (ci, co, s) in t.items():
        if name.endswith(k + ".conv1"):
            return 2 * px[s] * ci * co * 4
        if name.endswith(k + ".conv2"):
            return 2 * px[s] * (2 * co) * (2 * co) * 9
    if name.startswith("head_"):
        return 2 * px[2] * 128 * 32 * 9
    if name.startswith("chal_"):
        i = int(name[-1])
        return 2 * px[2 << i] * (128, 256, 512, 768, 512)[i] * (64, 128, 256, 384, 256)[i]
    if name == "spx2":
        return 2 * px[1] * 128 * 6 * 4
    return 0


for _ in range(3):
    m(fl, fr)
torch.cuda.synchronize()
steps = 5
rec = ops.LaunchRecorder(timing=True)
ops.record_launches(rec)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    m(fl, fr)
e1.record()
torch.cuda.synchronize()
ops.record_launches(None)
ms = e0.elapsed_time(e1) / steps
d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
d0.record()
for _ in range(steps):
    m.decoder(fl, fr)
d1.record()
torch.cuda.synchronize()
ms_dec = d0.elapsed_time(d1) / steps
durs = rec.durations_ms()
rows, dec_flops, dec_ms = [], 0.0, 0.0
for name, v in sorted(durs.items(), key=lambda kv: -sum(kv[1])):
    t = sum(v) / steps
    f = flops(name) * B * (len(v) / steps if name.startswith(("feature_up", "chal_", "head_")) else 1)
    # feature_up layers run twice per step (left + right) under one label; len(v)/steps launches share the label
    f = flops(name) * B * (len(v) / steps)
    rows.append(dict(name=name, ms_per_step=round(t, 4), launches=len(v) / steps, tflops=round(f / t / 1e9, 1) if f else None))
    if flops(name):
        dec_flops += f
        dec_ms += t
for r in rows[:40]:
    print(r)
res = dict(B=B, ms_per_step_head=round(ms, 3), pairs_per_s_head=round(B / ms * 1e3, 1), ms_per_step_decoder=round(ms_dec, 3),
           decoder_conv_tflops=round(dec_flops / dec_ms / 1e9, 1), decoder_conv_gflop_per_pair=round(dec_flops / B / 1e9, 1), kernels=rows)
print({k: v for k, v in res.items() if k != "kernels"})
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/bench_decoder.json", "w"), indent=1)
