"""Minimal usage example (B200): everything after the backbone on synthetic pyramids, eager and as a CUDA-graph replay.

    python tools/example_inference.py [--batch 1] [--size 1024]

With a real reference checkpoint: `head.load_state_dict(torch.load(path)["model"])` (reference key names, 'module.' prefix accepted),
and `semstereo_b200.patch.patch_model(model)` keeps the reference model object with its own backbone."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from semstereo_b200.decoder import StereoHead
from semstereo_b200.graph import GraphedCall
from semstereo_b200.params import BACKBONE_CHANS, make_decoder_params, make_params

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--size", type=int, default=1024)
a = ap.parse_args()
dev = torch.device("cuda:0")
sd = dict(make_params(seed=1, peaked=20.0))
sd.update(make_decoder_params(seed=2))
head = StereoHead(maxdisp=64)
head.load_state_dict(sd, strict=True)
head = head.to(dev)
g = torch.Generator(device=dev).manual_seed(0)
feats = {f"{s}{i}": torch.randn(a.batch, c, a.size // st, a.size // st, device=dev, generator=g)
         for s in "lr" for i, (c, st) in enumerate(zip(BACKBONE_CHANS, (2, 4, 8, 16, 32)))}
call = lambda f: head([f[f"l{i}"] for i in range(5)], [f[f"r{i}"] for i in range(5)])      # noqa: E731
out = call(feats)
disp, label = head.as_model_outputs(out)
print("disparity", tuple(disp[0].shape), "label logits", tuple(label.shape), "range", float(disp[0].min()), float(disp[0].max()))
graph = GraphedCall(call, feats)
for name, fn in (("eager", lambda: call(feats)), ("cuda graph", graph.replay)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    print(f"{name}: {(time.perf_counter() - t0) * 100:.2f} ms per batch of {a.batch}")
