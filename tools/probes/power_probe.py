"""Diagnostic: is the steady-state step time set by the 1000 W power cap?  Times the CUDA-graph replay of the default hot path at batch 8
(a) back to back (what bench.py reports) and (b) with an idle gap before every step (the GPU cools down / the power average drops)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from semstereo_b200.graph import GraphedCall
from semstereo_b200.hotpath import DisparityHotPath
from semstereo_b200.params import make_inputs, make_params

B = 8
m = DisparityHotPath(64, False, True)
m.load_state_dict(make_params(seed=1, peaked=20.0), strict=True)
m = m.to("cuda:0")
inp = {k: v.to("cuda:0") for k, v in make_inputs(100, B, 1024, 1024).items() if k not in ("cf_l", "cf_r")}
ORDER = ("f8_l", "f8_r", "f4_l", "f4_r", "cf_l", "cf_r", "spx_pred", "pred_label")
g = GraphedCall(lambda st: m(*[st.get(k) for k in ORDER])["pred_up"], inp)
for _ in range(5):
    g.replay()
torch.cuda.synchronize()


def timed(n, gap):
    ts = []
    for _ in range(n):
        if gap:
            torch.cuda.synchronize(); time.sleep(gap)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0], ts[-1]


a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(100):
    g.replay()
b.record(); b.synchronize()
print("back to back, 100 steps: %.3f ms/step" % (a.elapsed_time(b) / 100))
print("single steps, no gap  (median, min, max): %.3f %.3f %.3f" % timed(30, 0))
print("single steps, 50 ms idle before each    : %.3f %.3f %.3f" % timed(30, 0.05))
print("single steps, 500 ms idle before each   : %.3f %.3f %.3f" % timed(10, 0.5))
os.system("nvidia-smi --query-gpu=power.draw,clocks.sm,clocks_throttle_reasons.active --format=csv,noheader")
