"""BASELINE config #2: standalone gwc / gwc_norm / concat cost-volume build sweep at 1/4-res feature size (256x256) on one B200,
signed (D = 2M, models/submodule.py) and unsigned (D = M, models/submodule_.py).
Reports algorithmic GB/s = 4*(2*C*H*W + Cout*D*H*W)*B / t against the measured HBM copy peak, next to two torch-op baselines on
the same GPU:
  `torch_ref_form`      : the REFERENCE's formulation, restated with torch ops -- one slice / (normalise both slices) / multiply /
                          mean per disparity (groupwise_correlation_norm normalises inside the loop, submodule.py:213-238);
  `torch_norm_once_loop`: a cheaper torch loop that normalises once (not what the reference does; round 1 reported only this one).
L2 is flushed between timed launches; SM clocks and throttle reasons are sampled during the run and written into the JSON."""
import json
import os
import subprocess
import sys
import threading

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from semstereo_b200 import ops  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dev = "cuda:0"
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
PEAK, PEAK_SRC = (json.load(open(pk))["hbm_gbs"], "measured") if os.path.exists(pk) else (6650.0, "fallback")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


class Clocks:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __enter__(self):
        self.lines = []
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", "0", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.p.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.p = None
        return self

    def __exit__(self, *a):
        if self.p:
            self.p.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) >= 6:
                try:
                    sm.append(float(f[0])); mx.append(float(f[1]))
                except ValueError:
                    continue
                for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


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


def disps(M, signed):
    return range(-M, M) if signed else range(0, M)


def torch_ref_form(l, r, M, G, norm, signed):
    """The reference loop (submodule.py:198-238 / submodule_.py:188-221): per disparity, slice both maps, (normalise), multiply, mean."""
    B, C, H, W = l.shape
    cg = C // G

    def corr(a, b):
        a, b = a.reshape(B, G, cg, H, a.shape[-1]), b.reshape(B, G, cg, H, b.shape[-1])
        if norm:
            a = a / (torch.norm(a, 2, 2, True) + 1e-05)
            b = b / (torch.norm(b, 2, 2, True) + 1e-05)
        return (a * b).mean(dim=2)

    vol = l.new_zeros(B, G, len(disps(M, signed)), H, W)
    for k, d in enumerate(disps(M, signed)):
        if d < 0:
            vol[:, :, k, :, :d] = corr(l[..., :d], r[..., -d:])
        elif d > 0:
            vol[:, :, k, :, d:] = corr(l[..., d:], r[..., :-d])
        else:
            vol[:, :, k] = corr(l, r)
    return vol.contiguous()


def torch_norm_once_loop(l, r, M, G, norm, signed):
    B, C, H, W = l.shape
    lg, rg = l.view(B, G, C // G, H, W), r.view(B, G, C // G, H, W)
    if norm:
        lg = lg / (lg.norm(2, 2, True) + 1e-5)
        rg = rg / (rg.norm(2, 2, True) + 1e-5)
    vol = l.new_zeros(B, G, len(disps(M, signed)), H, W)
    for k, d in enumerate(disps(M, signed)):
        if d < 0:
            vol[:, :, k, :, :d] = (lg[..., :d] * rg[..., -d:]).mean(2)
        elif d > 0:
            vol[:, :, k, :, d:] = (lg[..., d:] * rg[..., :-d]).mean(2)
        else:
            vol[:, :, k] = (lg * rg).mean(2)
    return vol


res = []
H = W = 256
with Clocks() as clk:
    for signed in (True, False):
        for B in (1, 8):
            for C, G, M in [(64, 8, 16), (128, 8, 16), (128, 32, 16), (128, 32, 32), (256, 32, 16), (128, 16, 24), (128, 32, 48)]:
                l, r = torch.randn(B, C, H, W, device=dev), torch.randn(B, C, H, W, device=dev)
                D = 2 * M if signed else M
                nbytes = 4 * (2 * C * H * W + G * D * H * W) * B
                for norm in (False, True):
                    ms = timeit(lambda: ops.gwc_volume(l, r, M, G, signed, norm))
                    ms_ref = timeit(lambda: torch_ref_form(l, r, M, G, norm, signed), 3)
                    ms_once = timeit(lambda: torch_norm_once_loop(l, r, M, G, norm, signed), 3)
                    rec = dict(op="gwc_norm" if norm else "gwc", signed=signed, B=B, C=C, G=G, M=M, D=D, ms=round(ms, 4), gbs=round(nbytes / ms / 1e6, 1),
                               frac=round(nbytes / ms / 1e6 / PEAK, 3), torch_ref_form_ms=round(ms_ref, 3), speedup_vs_ref_form=round(ms_ref / ms, 1),
                               torch_norm_once_loop_ms=round(ms_once, 3), speedup_vs_norm_once=round(ms_once / ms, 1))
                    print(rec); res.append(rec)
            for C, M in [(32, 16), (32, 32), (64, 16)]:
                l, r = torch.randn(B, C, H, W, device=dev), torch.randn(B, C, H, W, device=dev)
                D = 2 * M if signed else M
                nbytes = 4 * (2 * C * H * W + 2 * C * D * H * W) * B
                ms = timeit(lambda: ops.concat_volume(l, r, M, signed))
                rec = dict(op="concat", signed=signed, B=B, C=C, M=M, D=D, ms=round(ms, 4), gbs=round(nbytes / ms / 1e6, 1), frac=round(nbytes / ms / 1e6 / PEAK, 3))
                print(rec); res.append(rec)
    # the model's own call (SemStereo.py:273): C=256, G=32, M=8 at 128x128
    for B in (1, 8):
        l, r = torch.randn(B, 256, 128, 128, device=dev), torch.randn(B, 256, 128, 128, device=dev)
        nbytes = 4 * (2 * 256 + 32 * 16) * 128 * 128 * B
        ms = timeit(lambda: ops.gwc_volume(l, r, 8, 32, True, True))
        ms_ref = timeit(lambda: torch_ref_form(l, r, 8, 32, True, True), 3)
        rec = dict(op="gwc_norm(model call)", signed=True, B=B, C=256, G=32, M=8, D=16, ms=round(ms, 4), gbs=round(nbytes / ms / 1e6, 1),
                   frac=round(nbytes / ms / 1e6 / PEAK, 3), torch_ref_form_ms=round(ms_ref, 3), speedup_vs_ref_form=round(ms_ref / ms, 1))
        print(rec); res.append(rec)
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"peak_gbs": PEAK, "peak_source": PEAK_SRC + " (copy)", "clocks": clk.summary(), "l2": "256 MiB buffer written between timed launches",
           "baselines": {"torch_ref_form": "reference formulation (per-disparity slice / normalise / multiply / mean) as torch ops on the same GPU",
                         "torch_norm_once_loop": "torch loop that normalises once: NOT the reference op, cheaper"},
           "results": res}, open("gpurun_out/bench_volumes.json", "w"), indent=1)
