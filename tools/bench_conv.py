"""Micro-benchmark of the tensor-core conv3d kernels on the real layer shapes (CUDA events, L2 flushed between runs)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from semstereo_b200 import ops

dev = "cuda:0"
LAYERS = [("concat_stem", 64, 32, 24, 256, 256), ("classif.0", 32, 32, 24, 256, 256), ("hourglass.conv2", 64, 64, 12, 128, 128),
          ("hourglass.conv4", 128, 128, 6, 64, 64), ("classif_att_.0", 32, 32, 16, 128, 128), ("hourglass_att.conv2", 64, 64, 8, 64, 64),
          ("hourglass_att.conv4", 128, 128, 4, 32, 32)]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
res = []
for B in (1, 8):
    for name, ci, co, D, H, W in LAYERS:
        x = torch.randn(B, ci // 8, D, H, W, 8, device=dev).to(torch.bfloat16)
        w = ops.pack_conv3d_weight_tc(torch.randn(co, ci, 3, 3, 3) / (27 * ci) ** 0.5).to(dev)
        sc, sh = torch.ones(co, device=dev), torch.zeros(co, device=dev)
        for _ in range(3):
            ops.conv3d_tc(x, w, sc, sh, relu=True)
        ts = []
        for _ in range(5):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ops.conv3d_tc(x, w, sc, sh, relu=True); b.record()
            torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        ms = sorted(ts)[len(ts) // 2]
        fl = 2 * 27 * ci * co * D * H * W * B
        r = dict(layer=name, B=B, ms=round(ms, 4), tflops=round(fl / ms / 1e9, 1), io_gbs=round((ci + co) * 2 * D * H * W * B / ms / 1e6, 1))
        print(r); res.append(r)
json.dump(res, open("gpurun_out/bench_conv.json", "w"), indent=1)
