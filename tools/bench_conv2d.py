"""Micro-benchmark of the tensor-core 2-D decoder convs on the real layer shapes at 1024x1024 (CUDA events, L2 flushed).
Usage: python tools/bench_conv2d.py [B] [name filter ...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from semstereo_b200 import ops_tc as tc

dev = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
# name, mode, (C0, C1), Cout, input H = W
LAYERS = [("deconv4_2.conv2", tc.CONV3, (64, 64), 128, 512), ("deconv8_4.conv2", tc.CONV3, (128, 128), 256, 256),
          ("deconv16_8.conv2", tc.CONV3, (256, 256), 512, 128), ("deconv32_16.conv2", tc.CONV3, (384, 384), 768, 64),
          ("deconv32_16.conv1", tc.DECONV4, (512, 0), 384, 32), ("deconv16_8.conv1", tc.DECONV4, (768, 0), 256, 64),
          ("deconv8_4.conv1", tc.DECONV4, (512, 0), 128, 128), ("deconv4_2.conv1", tc.DECONV4, (256, 0), 64, 256),
          ("head.conv1", tc.CONV3, (128, 0), 32, 512), ("chal_1", tc.CONV1, (256, 0), 128, 256), ("spx2", tc.DECONV4, (128, 0), 6, 512)]
if len(sys.argv) > 2:
    LAYERS = [l for l in LAYERS if any(a in l[0] for a in sys.argv[2:])]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
res = []
for name, mode, (c0, c1), co, hw in LAYERS:
    x0 = torch.randn(B, c0 // 8, hw, hw, 8, device=dev).to(torch.bfloat16)
    x1 = torch.randn(B, c1 // 8, hw, hw, 8, device=dev).to(torch.bfloat16) if c1 else None
    cin = c0 + c1
    if mode == tc.DECONV4:
        w, taps = torch.randn(cin, co, 4, 4) / (4 * cin) ** 0.5, 4
    else:
        k = 3 if mode == tc.CONV3 else 1
        w, taps = torch.randn(co, cin, k, k) / (k * k * cin) ** 0.5, k * k
    wp = tc.pack_weight2d(w, mode).to(dev)
    sc, sh = torch.ones(co, device=dev), torch.zeros(co, device=dev)
    f32 = co % 8 != 0
    run = lambda: tc.conv2d_tc(mode, x0, wp, co, sc, sh, relu=True, out_f32=f32, x1=x1)
    for _ in range(3):
        run()
    ts = []
    for _ in range(5):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record()
        torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts) // 2]
    opx = hw * hw * (4 if mode == tc.DECONV4 else 1)
    fl = 2 * taps * cin * co * opx * B
    r = dict(layer=name, B=B, ms=round(ms, 4), tflops=round(fl / ms / 1e9, 1))
    print(r); res.append(r)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/bench_conv2d.json", "w"), indent=1)
