"""Micro-benchmark of the tensor-core conv3d kernels on the real layer shapes (CUDA events, L2 flushed between runs)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from semstereo_b200 import ops_tc as tc

dev = "cuda:0"
# name, kind, Cin, Cout, input D, H, W
LAYERS = [("concat_stem", tc.S1, 64, 32, 24, 256, 256), ("classif.0", tc.S1, 32, 32, 24, 256, 256), ("classif.2", tc.S1, 32, 1, 24, 256, 256),
          ("hourglass.conv1", tc.S2, 32, 64, 24, 256, 256), ("hourglass.conv2", tc.S1, 64, 64, 12, 128, 128),
          ("hourglass.conv3", tc.S2, 64, 128, 12, 128, 128), ("hourglass.conv4", tc.S1, 128, 128, 6, 64, 64),
          ("hourglass.conv5", tc.T2, 128, 64, 6, 64, 64), ("hourglass.conv6", tc.T2, 64, 32, 12, 128, 128),
          ("hourglass.redir1", tc.K1, 32, 32, 24, 256, 256), ("hourglass.redir2", tc.K1, 64, 64, 12, 128, 128),
          ("classif_att_.0", tc.S1, 32, 32, 16, 128, 128), ("hourglass_att.conv2", tc.S1, 64, 64, 8, 64, 64),
          ("concat_stem[folded]", tc.S1F, 64, 32, 24, 256, 256), ("classif.0[folded]", tc.S1F, 32, 32, 24, 256, 256),
          ("hourglass.conv2[folded]", tc.S1F, 64, 64, 12, 128, 128), ("classif_att_.0[folded]", tc.S1F, 32, 32, 16, 128, 128)]
if len(sys.argv) > 1:
    LAYERS = [l for l in LAYERS if any(a in l[0] for a in sys.argv[1:])]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
res = []
for B in (1, 8):
    for name, kind, ci, co, D, H, W in LAYERS:
        if kind == tc.S2:
            x = torch.randn(B, 8, ci // 8, D // 2, H // 2, W // 2, 8, device=dev).to(torch.bfloat16)
            vox = D * H * W // 8
        else:
            x = torch.randn(B, ci // 8, D, H, W, 8, device=dev).to(torch.bfloat16)
            vox = D * H * W * (8 if kind == tc.T2 else 1)
        taps = 1 if kind == tc.K1 else 27
        wshape = (ci, co, 3, 3, 3) if kind == tc.T2 else (co, ci, 3 if taps == 27 else 1, 3 if taps == 27 else 1, 3 if taps == 27 else 1)
        w = tc.pack_weight(torch.randn(*wshape) / (taps * ci) ** 0.5, kind).to(dev)
        sc, sh = torch.ones(co, device=dev), torch.zeros(co, device=dev)
        f32 = co % 8 != 0
        run = lambda: tc.conv3d_tc(kind, x, w, co, sc, sh, relu=True, out_mode=tc.F32 if f32 else tc.BLOCKED)
        for _ in range(3):
            run()
        ts = []
        for _ in range(5):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); run(); b.record()
            torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        ms = sorted(ts)[len(ts) // 2]
        fl = 2 * taps * ci * co * vox * B / (8 if kind == tc.T2 else 1)
        r = dict(layer=name, B=B, ms=round(ms, 4), tflops=round(fl / ms / 1e9, 1))
        print(r); res.append(r)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/bench_conv.json", "w"), indent=1)
