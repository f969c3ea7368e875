"""BASELINE config #2: standalone gwc / gwc_norm / concat cost-volume build sweep at 1/4-res feature size (256x256) on one B200.
Reports algorithmic GB/s = 4*(2*C*H*W + Cout*D*H*W)*B / t against the measured HBM peak, next to a plain torch-op loop over
disparities on the same GPU (the reference's formulation: one slice-multiply-mean per disparity)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from semstereo_b200 import ops

dev = "cuda:0"
PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def torch_loop_gwc(l, r, M, G, norm):
    B, C, H, W = l.shape
    lg, rg = l.view(B, G, C // G, H, W), r.view(B, G, C // G, H, W)
    if norm:
        lg = lg / (lg.norm(2, 2, True) + 1e-5)
        rg = rg / (rg.norm(2, 2, True) + 1e-5)
    vol = l.new_zeros(B, G, 2 * M, H, W)
    for k, d in enumerate(range(-M, M)):
        if d < 0:
            vol[:, :, k, :, :d] = (lg[..., :d] * rg[..., -d:]).mean(2)
        elif d > 0:
            vol[:, :, k, :, d:] = (lg[..., d:] * rg[..., :-d]).mean(2)
        else:
            vol[:, :, k] = (lg * rg).mean(2)
    return vol


res = []
H = W = 256
for B in (1, 8):
    for C, G, M in [(64, 8, 16), (128, 8, 16), (128, 32, 16), (128, 32, 32), (256, 32, 16), (128, 16, 24), (128, 32, 48)]:
        l, r = torch.randn(B, C, H, W, device=dev), torch.randn(B, C, H, W, device=dev)
        nbytes = 4 * (2 * C * H * W + G * 2 * M * H * W) * B
        for norm in (False, True):
            ms = timeit(lambda: ops.gwc_volume(l, r, M, G, True, norm))
            ms_t = timeit(lambda: torch_loop_gwc(l, r, M, G, norm), 3)
            rec = dict(op="gwc_norm" if norm else "gwc", B=B, C=C, G=G, M=M, ms=round(ms, 4), gbs=round(nbytes / ms / 1e6, 1),
                       frac=round(nbytes / ms / 1e6 / PEAK, 3), torch_loop_ms=round(ms_t, 3), speedup=round(ms_t / ms, 1))
            print(rec); res.append(rec)
    for C, M in [(32, 16), (32, 32), (64, 16)]:
        l, r = torch.randn(B, C, H, W, device=dev), torch.randn(B, C, H, W, device=dev)
        nbytes = 4 * (2 * C * H * W + 2 * C * 2 * M * H * W) * B
        ms = timeit(lambda: ops.concat_volume(l, r, M, True))
        rec = dict(op="concat", B=B, C=C, M=M, ms=round(ms, 4), gbs=round(nbytes / ms / 1e6, 1), frac=round(nbytes / ms / 1e6 / PEAK, 3))
        print(rec); res.append(rec)
# the model's own call (SemStereo.py:273): C=256, G=32, M=8 at 128x128
for B in (1, 8):
    l, r = torch.randn(B, 256, 128, 128, device=dev), torch.randn(B, 256, 128, 128, device=dev)
    nbytes = 4 * (2 * 256 + 32 * 16) * 128 * 128 * B
    ms = timeit(lambda: ops.gwc_volume(l, r, 8, 32, True, True))
    rec = dict(op="gwc_norm(model call)", B=B, C=256, G=32, M=8, ms=round(ms, 4), gbs=round(nbytes / ms / 1e6, 1), frac=round(nbytes / ms / 1e6 / PEAK, 3))
    print(rec); res.append(rec)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/bench_volumes.json", "w"), indent=1)
