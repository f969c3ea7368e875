#!/usr/bin/env python
"""bench.py — throughput of the SemStereo disparity hot path on B200 (stereo pairs / s).

    python bench.py --gpus 1 --steps 10 --warmup 3                       # this framework (CUDA kernels via the C-ABI)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                                  # the reference algorithm's CPU path (oracle port)

A "step" is one pass of the hot path (SemStereo.forward:273-324: gwc volume -> attention hourglass -> top-k ->
sparse concat volume -> hourglass2 -> regression_topk -> SSR upsample) over one batch of synthetic stereo-pair
features of a 1024x1024 US3D-shaped pair (maxdisp 64), random-init weights.  Weak scaling: every rank processes
`--batch` pairs per step; the outputs are gathered to rank 0 over NCCL every step (side stream, overlapping the next step).
One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

ORDER = ("f8_l", "f8_r", "f4_l", "f4_r", "cf_l", "cf_r", "spx_pred", "pred_label")


PRECISION_DTYPE = {"split": "bf16x3 (fp32-accurate) attention branch + bf16 aggregation", "bf16": "bf16", "fp32": "f32",
                   "mixed": "f32 attention branch + bf16 aggregation"}
PRECISION_NOTE = {
    "split": "attention branch (hourglass_att, classif_att_: decides the top-k samples) on tcgen05 with bf16x3 split operands "
             "(x*w = x_hi*w_hi + x_lo*w_hi + x_hi*w_lo, fp32 accumulation: fp32-accurate, ind_k == fp32 oracle on > 99.9 % of pixels); "
             "aggregation branch (concat_stem, hourglass, classif), concat_feature: bf16 operands / fp32 accumulation; fp32 elsewhere",
    "bf16": "bf16 operands / fp32 accumulation on the tensor cores for every 3-D conv, the window attention, concat_feature and (stage head) "
            "the 2-D decoder; fp32 elsewhere",
    "fp32": "fp32 everywhere (FFMA 3-D convs)",
    "mixed": "attention branch (hourglass_att, classif_att_) fp32 FFMA, aggregation branch bf16 tensor cores"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


# ---- algorithmic work per kernel label, per stereo pair at (H, W, maxdisp) ----------
def conv_layers(H, W, maxdisp, signed=True):
    """label -> (flops, kind) for every 3-D conv launch of the path."""
    d8 = (2 if signed else 1) * (maxdisp // 8)
    out = {}

    def hg(prefix, c, D, h, w):
        v = D * h * w
        out[prefix + ".conv1"] = 2 * 27 * c * 2 * c * v / 8
        out[prefix + ".conv2"] = 2 * 27 * 2 * c * 2 * c * v / 8
        out[prefix + ".conv3"] = 2 * 27 * 2 * c * 4 * c * v / 64
        out[prefix + ".conv4"] = 2 * 27 * 4 * c * 4 * c * v / 64
        out[prefix + ".conv5"] = 2 * 27 * 4 * c * 2 * c * v / 64
        out[prefix + ".conv6"] = 2 * 27 * 2 * c * c * v / 8
        out[prefix + ".redir1"] = 2 * c * c * v
        out[prefix + ".redir2"] = 2 * 2 * c * 2 * c * v / 8

    v8, v4 = d8 * (H // 8) * (W // 8), 24 * (H // 4) * (W // 4)
    hg("hourglass_att", 32, d8, H // 8, W // 8)
    out["classif_att_.0"] = 2 * 27 * 32 * 32 * v8
    out["classif_att_.2"] = 2 * 27 * 32 * v8
    out["concat_stem"] = 2 * 27 * 64 * 32 * v4
    hg("hourglass", 32, 24, H // 4, W // 4)
    out["classif.0"] = 2 * 27 * 32 * 32 * v4
    out["classif.2"] = 2 * 27 * 32 * v4
    return out


def decoder_layers(H, W):
    """label -> FLOPs per LAUNCH and image of the 2-D decoder convs (models/SemStereo.py:59-86, 196-216)."""
    px = lambda s: (H // s) * (W // s)      # noqa: E731
    out = {}
    for pre, table in (("feature_up.", (("deconv32_16", 512, 384, 16), ("deconv16_8", 768, 256, 8), ("deconv8_4", 512, 128, 4),
                                        ("deconv4_2", 256, 64, 2))),
                       ("", (("spx32_16", 256, 384, 16), ("spx16_8", 768, 256, 8), ("spx8_4", 512, 128, 4), ("spx4_2", 256, 64, 2)))):
        for name, ci, co, s in table:
            out[pre + name + ".conv1"] = 2 * px(s) * ci * co * 4            # k4 s2 transposed: 4 taps per output pixel
            out[pre + name + ".conv2"] = 2 * px(s) * (2 * co) * (2 * co) * 9
    for i, (ci, co) in enumerate(((128, 64), (256, 128), (512, 256), (768, 384), (512, 256))):
        out[f"chal_{i}"] = 2 * px(2 << i) * ci * co
    out["head_l"] = out["head_r"] = 2 * px(2) * (128 * 32 * 9 + 32 * 6)
    out["spx2"] = 2 * px(1) * 128 * 6 * 4
    return out


def hbm_bytes(H, W, maxdisp, signed=True):
    """label -> algorithmic bytes (fp32 in + out) of the memory-bound launches, per pair."""
    p8, p4, p1 = (H // 8) * (W // 8), (H // 4) * (W // 4), H * W
    d8 = (2 if signed else 1) * (maxdisp // 8)
    nb = 2 * d8
    return {
        "ss_gwc_volume": 4 * (2 * 256 * p8 + 32 * d8 * p8),
        "ss_patch_gate": 4 * (2 * 32 * d8 * p8 + 32 * p8),
        "ss_att_stats": 4 * (d8 * p8 + nb * p4 + 2 * p4),
        "ss_sample_strength": 4 * (2 * 128 * p4 + 2 * p4 + 5 * p4),
        "ss_topk_select": 4 * (nb * p4 + 5 * p4 + 2 * 24 * p4 + p4),
        "ss_sparse_concat_volume": 4 * (2 * 32 * p4 + 2 * 24 * p4 + 64 * 24 * p4),
        "ss_patch_gate_blocked": 4 * (32 * d8 * p8 + 32 * p8) + 2 * 32 * d8 * p8,          # fp32 volume + gate in, bf16 out
        "ss_sparse_concat_volume_blocked": 4 * (2 * 32 * p4 + 2 * 24 * p4) + 2 * 64 * 24 * p4,   # bf16 volume out
        "ss_regression_topk": 4 * (2 * 24 * p4 + p4),
        "ss_ssr_upsample": 4 * (p4 + 12 * p1 + p1),
        "ss_ssr_upsample2": 4 * (2 * p4 + 12 * p1 + 2 * p1),      # two low-res maps in, spx + label once, two full-res maps out
    }


def ncu_traffic(kernel, precision, batch):
    """DRAM bytes (read + write) per launch of `kernel` from the committed `ncu --set full` capture (profiles/r0N_traffic.json:
    bytes per stereo pair, measured at the batch stated there), scaled to this run's batch; None when no capture exists."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(p):
            continue
        tab = json.load(open(p))
        # the aggregation branch (incl. the dominant concat_stem kernel) is the same bf16 kernel in the "split" and "mixed" modes
        ent = tab.get(precision, {}).get(kernel) or (tab.get("bf16", {}).get(kernel) if precision in ("split", "mixed") else None)
        if ent is not None:
            return int(ent["dram_bytes_per_pair"] * batch)
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def drop_cf(inp, external_cf):
    """Default workload: concat_feature(f4_*) (SemStereo.py:314-315) runs INSIDE the path, so cf_l / cf_r are not inputs."""
    return inp if external_cf else {k: v for k, v in inp.items() if k not in ("cf_l", "cf_r")}


def cpu_reference(H, W, maxdisp, steps, warmup, signed=True, att_only=False, external_cf=False, stage="path"):
    """The reference algorithm's CPU path (oracle port of SemStereo.forward:273-324; efficient closed forms, so it is
    FASTER than the reference's own Python-loop volume builder — a conservative baseline).  Each step = one pair."""
    from oracle import hotpath as oh
    from semstereo_b200.params import make_inputs, make_params
    torch.set_num_threads(os.cpu_count() or 1)
    p = make_params(seed=1, peaked=20.0)
    inp = drop_cf(make_inputs(3, 1, H, W), external_cf)
    if stage in ("head", "full"):
        from oracle import decoder as od
        from semstereo_b200.params import make_backbone_features, make_decoder_params
        p = dict(p)
        p.update(make_decoder_params(seed=2))
        fl, fr = make_backbone_features(3, 1, H, W)
    if stage == "full":
        from oracle import backbone as ob
        from semstereo_b200.params import make_backbone_params, make_images
        hf = ob.build(make_backbone_params(seed=4))
        left, right = make_images(3, 1, H, W)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        if stage == "full":
            with torch.no_grad():
                fl = list(hf(left, output_hidden_states=True).hidden_states)
                fr = list(hf(right, output_hidden_states=True).hidden_states)
        if stage in ("head", "full"):
            d = od.forward(p, fl, fr, right_label=False)
            inp = {k: d[k] for k in ("f8_l", "f8_r", "f4_l", "f4_r", "spx_pred", "pred_label")}
        oh.forward(p, inp, maxdisp, signed=signed, att_weights_only=att_only)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return times


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="stereo pairs per GPU per step (BASELINE config #3: 8)")
    ap.add_argument("--height", type=int, default=1024)
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--maxdisp", type=int, default=64)
    ap.add_argument("--cpu-steps", type=int, default=2, help="pairs timed for the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default="split", choices=["split", "bf16", "fp32", "mixed"],
                    help="split (default): attention branch with fp32-accurate bf16x3 products on the tensor cores (sample selection = the "
                         "fp32 oracle's), aggregation branch bf16 (BASELINE config #3 'bf16 aggregation'); bf16: everything bf16; "
                         "mixed: attention branch on the fp32 FFMA pipe; fp32: everything FFMA")
    ap.add_argument("--variant", default="us3d", choices=["us3d", "whu"],
                    help="us3d: SemStereo, signed, 1024x1024, maxdisp 64 (configs #1/#3); whu: SemStereo_WHU + submodule_.py, unsigned, "
                         "384x768, maxdisp 128 (config #4)")
    ap.add_argument("--att-only", action="store_true", help="attention_weights_only forward (the forward half of config #5)")
    ap.add_argument("--stage", default="path", choices=["path", "head", "full"],
                    help="path: the disparity hot path (forward:273-324, BASELINE north_star; default).  head: everything after the "
                         "backbone (forward:249-346): the 2-D decoder (SURVEY 8(f) rank 1) + the path; inputs are the backbone pyramids.  "
                         "full: the whole SemStereo.forward from the two images (MobileViTv2 backbone, SURVEY 8(f) rank 2, + decoder + path)")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches in region 1 instead of CUDA-graph replays")
    ap.add_argument("--no-full-model", action="store_true",
                    help="skip the `full_model` sub-measurement (the whole model from the images, --stage full) of the default line")
    ap.add_argument("--external-cf", action="store_true",
                    help="hand concat_feature(f4_*) in as inputs (round-1 boundary) instead of computing it inside the path")
    ap.add_argument("--train", action="store_true",
                    help="BASELINE config #5: attention_weights_only TRAINING step (forward + losses + backward + gradient all-reduce + Adam), "
                         "--batch pairs per GPU (default 2 => 16 on 8 GPUs); delegates to tools/train_step.py")
    a = ap.parse_args()
    if a.train:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import train_step
        return train_step.main(["--gpus", str(a.gpus), "--steps", str(a.steps), "--warmup", str(a.warmup), "--height", str(a.height),
                                "--width", str(a.width), "--maxdisp", str(a.maxdisp), "--batch", str(2 if a.batch == 8 else a.batch)])
    signed = a.variant == "us3d"
    if a.variant == "whu":
        if a.height == 1024 and a.width == 1024:
            a.height, a.width = 384, 768
        if a.maxdisp == 64:
            a.maxdisp = 128
    H, W, md = a.height, a.width, a.maxdisp
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    head = a.stage in ("head", "full")
    full = a.stage == "full"
    if head:
        if a.att_only:
            raise SystemExit("--stage head runs the full forward")
        a.precision = "bf16"                   # the decoder has no fp32-accurate mode yet
    workload = (f"{'SemStereo' if signed else 'SemStereo_WHU'} "
                f"{'whole forward from the images: MobileViTv2 backbone + decoder + disparity path (forward:246-346)' if a.stage == 'full' else ('decoder + disparity path = everything after the backbone (forward:249-346)' if head else 'disparity hot path (forward:273-324)')}, {H}x{W} "
                f"{'US3D' if signed else 'WHU'}-shaped pairs, maxdisp {md}, {'signed' if signed else 'unsigned'}"
                f"{', attention_weights_only' if a.att_only else ''}"
                f"{'' if a.external_cf or a.att_only else ', concat_feature (:314-315) computed inside the path'}")

    if a.impl == "reference":
        if rank != 0:
            return
        times = cpu_reference(H, W, md, a.steps, min(a.warmup, 1), signed, a.att_only, a.external_cf, a.stage)
        ms = 1e3 * sum(times) / len(times)
        v = 1e3 / ms
        print(json.dumps({
            "impl": "reference", "metric": "stereo pairs/sec", "value": v, "unit": "pairs/s", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": min(a.warmup, 1), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": workload, "pairs_per_step": 1},
            "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{a.steps} steps x 1 pair, oracle/hotpath.py (torch CPU fp32)"},
            "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    assert torch.cuda.is_available(), "bench.py (impl b200) needs a CUDA device: there is no CPU fallback"
    from semstereo_b200 import dist as sdist, ops
    from semstereo_b200.hotpath import DisparityHotPath
    from semstereo_b200.params import make_inputs, make_params
    import torch.distributed as tdist
    rank, local, world = sdist.init_from_env("nccl")
    torch.cuda.set_device(local)
    all_cpus = os.sched_getaffinity(0)
    numa_cpus = sdist.bind_to_gpu_numa(local)       # before any pinned allocation: NUMA-local staging buffers
    dev = torch.device("cuda", local)
    B = a.batch
    okey = "pred_att_up" if a.att_only else "pred_up"
    if full:
        from semstereo_b200.backbone import SemStereoB200
        from semstereo_b200.params import make_backbone_params, make_decoder_params, make_images
        model = SemStereoB200(md, False, signed)
        sd = dict(make_params(seed=1, peaked=20.0))
        sd.update(make_decoder_params(seed=2))
        sd.update({"feature." + k: v for k, v in make_backbone_params(seed=4).items()})
        model.load_state_dict(sd, strict=True)
        left, right = make_images(100 + rank, B, H, W)
        host = {"left": left.pin_memory(), "right": right.pin_memory()}
        call = lambda st: model(st["left"], st["right"])[okey]      # noqa: E731
    elif head:
        from semstereo_b200.decoder import StereoHead
        from semstereo_b200.params import make_backbone_features, make_decoder_params
        model = StereoHead(md, False, signed)
        sd = dict(make_params(seed=1, peaked=20.0))
        sd.update(make_decoder_params(seed=2))
        model.load_state_dict(sd, strict=True)
        fl, fr = make_backbone_features(100 + rank, B, H, W)
        host = {f"l{i}": t.pin_memory() for i, t in enumerate(fl)}
        host.update({f"r{i}": t.pin_memory() for i, t in enumerate(fr)})
        call = lambda st: model([st[f"l{i}"] for i in range(5)], [st[f"r{i}"] for i in range(5)])[okey]      # noqa: E731
    else:
        model = DisparityHotPath(md, a.att_only, signed, precision=a.precision)
        model.load_state_dict(make_params(seed=1, peaked=20.0), strict=True)
        host = {k: v.pin_memory() for k, v in drop_cf(make_inputs(100 + rank, B, H, W), a.external_cf and not a.att_only).items()}
        call = lambda st: model(*[st.get(k) for k in ORDER])[okey]      # noqa: E731
    model = model.to(dev)
    devin = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    out_host = torch.empty((B, H, W), dtype=torch.float32).pin_memory()
    # Output gather (the only collective): to rank 0, as nn.DataParallel gathers to device 0 (main_us3d.py:100), issued on a SIDE
    # stream from a double-buffered copy of the result, so that it overlaps the next step's kernels instead of serialising with
    # them (round 1: an eager all-gather on the compute stream cost 6 % of a step at N = 8).
    gathered = torch.empty((world * B, H, W), device=dev) if (world > 1 and rank == 0) else None
    gather_stream = torch.cuda.Stream(dev) if world > 1 else None
    gbuf = [torch.empty((B, H, W), device=dev) for _ in range(2)] if world > 1 else None
    gdone = [torch.cuda.Event() for _ in range(2)] if world > 1 else None
    gstep_no = [0]

    def gather_async(o):
        """Copy the step's result aside (33 MB device-to-device) and gather it to rank 0 on the side stream."""
        i = gstep_no[0] % 2
        gstep_no[0] += 1
        cur = torch.cuda.current_stream(dev)
        cur.wait_event(gdone[i])                       # the gather that last read this buffer (two steps ago) has finished
        gbuf[i].copy_(o)
        gather_stream.wait_stream(cur)
        with torch.cuda.stream(gather_stream):
            tdist.gather(gbuf[i], list(gathered.split(B)) if rank == 0 else None, dst=0)
            gdone[i].record(gather_stream)
        return o

    def step(inputs):
        o = call(inputs)
        if world > 1:
            gather_async(o)
        return o

    def barrier():
        if world > 1:
            torch.cuda.synchronize()                   # incl. the side-stream gathers of this rank
            tdist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        step(devin)
    barrier()
    # ---- instrumented pass: the K steps launched eagerly with a CUDA-event pair around every kernel (per-kernel times,
    #      roofline).  The events and the Python launch path cost ~0.3 ms per step, which matters at small batches. ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    rec = ops.LaunchRecorder(timing=True)
    ops.record_launches(rec)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step(devin)
    e1.record()
    barrier()
    ops.record_launches(None)
    ms_eager = e0.elapsed_time(e1)
    # ---- timed region 1 (`value`): inputs resident in HBM, EXACTLY K steps; each step = one CUDA-graph replay of the forward
    #      (semstereo_b200.graph.GraphedCall: the same kernels, bit-identical results) + the eager NCCL gather.  --no-graph
    #      times the eager launches instead. ----
    ms_total = ms_eager
    graph_used, graph_err = False, ""
    if not a.no_graph:
        from semstereo_b200.graph import GraphedCall
        gc = None
        try:
            gc = GraphedCall(call, devin)
        except Exception as e:             # measurement falls back to the eager launches (all ranks together); the product is untouched
            graph_err = f"{type(e).__name__}: {e}"[:200]
            torch.cuda.synchronize()
        ok = torch.tensor([0 if gc is None else 1], device=dev)
        if world > 1:
            tdist.all_reduce(ok, op=tdist.ReduceOp.MIN)
        graph_used = bool(ok.item())
    if graph_used:

        def gstep():
            o = gc.replay()
            if world > 1:
                gather_async(o)

        for _ in range(a.warmup):
            gstep()
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(a.steps):
            gstep()
        g1.record()
        barrier()
        ms_total = g0.elapsed_time(g1)
        del gc
    # ---- timed region 2: end to end from pinned host buffers, result read back ----
    h2d = sum(v.numel() * 4 for v in host.values())
    d2h = out_host.numel() * 4
    barrier()
    # public API: HostPipeline (double-buffered: the H2D of step i+1 overlaps the compute of step i; every step copies all
    # eight input tensors from pinned host memory and reads pred_up back to pinned host memory)
    from semstereo_b200.pipeline import HostPipeline

    def gather(o):
        if world > 1:
            gather_async(o)
        return o

    pipe = HostPipeline(model, depth=2, post=gather, keys=tuple(host), call=call)
    for _ in pipe.run(host for _ in range(2)):      # warm the staging buffers
        pass
    barrier()
    t0 = time.perf_counter()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for res_host in pipe.run(host for _ in range(a.steps)):
        pass
    f1.record()
    barrier()
    ms_e2e = max(f0.elapsed_time(f1), 0.0)
    ms_e2e_wall = (time.perf_counter() - t0) * 1e3
    ms_e2e = max(ms_e2e, ms_e2e_wall)               # device events and host wall clock agree up to launch latency; keep the larger
    # ---- timed region 3 (informative, not `e2e`): the same pipeline with the FEATURE maps shipped as bf16 from pinned host memory
    #      (half the PCIe bytes for 2/3 of the input; widened on the device).  The fp32 boundary above is PCIe-bound. ----
    ms_e2e_h, h2d_h = None, None
    if not head:
        host_h = {k: (v.to(torch.bfloat16).pin_memory() if k.startswith("f") else v) for k, v in host.items()}
        h2d_h = sum(v.numel() * v.element_size() for v in host_h.values())
        for _ in pipe.run(host_h for _ in range(2)):
            pass
        barrier()
        t0 = time.perf_counter()
        for res_host in pipe.run(host_h for _ in range(a.steps)):
            pass
        barrier()
        ms_e2e_h = (time.perf_counter() - t0) * 1e3
        del host_h
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms_total, ms_e2e, ms_eager, ms_e2e_h or 0.0], device=dev)
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        ms_total, ms_e2e, ms_eager, ms_e2e_h = t.tolist()
        ms_e2e_h = ms_e2e_h or None
    if rank != 0:
        if world > 1:
            tdist.destroy_process_group()
        return

    pk = peaks()
    durs = rec.durations_ms()
    per_step = {k: sum(v) / a.steps for k, v in durs.items()}
    total_k = sum(per_step.values())
    flops, nbytes = conv_layers(H, W, md, signed), hbm_bytes(H, W, md, signed)
    dflops = decoder_layers(H, W) if head else {}
    kernels = []
    for name, ms in sorted(per_step.items(), key=lambda kv: -kv[1]):
        launches = len(durs[name]) / a.steps
        ent = {"name": name, "ms_per_step": round(ms, 4), "share": round(ms / total_k, 4), "launches_per_step": launches}
        if name in flops or name in dflops:
            ach = (flops[name] if name in flops else dflops[name] * launches) * B / (ms * 1e-3) / 1e12
            ent.update(bound="tensor", achieved=round(ach, 3), unit="TFLOP/s", frac=round(ach / pk["tf_sust"], 5))
        elif name in nbytes:
            ach = nbytes[name] * B * launches / (ms * 1e-3) / 1e9
            ent.update(bound="hbm", achieved=round(ach, 1), unit="GB/s", frac=round(ach / pk["hbm"], 4))
        kernels.append(ent)
    dom = next(k for k in kernels if "bound" in k)
    roof = {"kernel": dom["name"], "bound": dom["bound"], "achieved": dom["achieved"],
            "peak": pk["tf_sust"] if dom["bound"] == "tensor" else pk["hbm"], "unit": dom["unit"], "frac": dom["frac"],
            "traffic": ncu_traffic(dom["name"], a.precision, B),
            **({"frac_of_burst_peak": round(dom["achieved"] / pk["tf_burst"], 4), "burst_peak": pk["tf_burst"]} if dom["bound"] == "tensor" else {}),
            "peak_source": pk["src"] + (" (sustained bf16 cuBLAS)" if dom["bound"] == "tensor" else " (copy)"),
            "note": ("tcgen05 bf16 implicit GEMM, fp32 accumulation in TMEM; achieved = Table-A FLOPs of the layer x batch / CUDA-event "
                     "time of the launch inside the timed region; `peak` is the SUSTAINED cuBLAS rate (8192^3 back to back for 4 s, "
                     "power-capped) as the contract asks for a kernel timed inside a long step -- a fraction above 1 means the kernel "
                     "sustains more than cuBLAS does under the same cap; frac_of_burst_peak is against cuBLAS' best-of-10" if a.precision != "fp32"
                     else "fp32-accurate FFMA mode; fraction is against the bf16 tensor peak")}
    value = world * B * a.steps / (ms_total * 1e-3)
    e2e_v = world * B * a.steps / (ms_e2e * 1e-3)
    res = {"metric": "stereo pairs/sec", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
           "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": PRECISION_DTYPE[a.precision],
           "data": "synthetic", "config": {"workload": workload, "pairs_per_gpu_per_step": B, "global_pairs_per_step": world * B,
                                           "l2": "per-step inputs (1.3 GB at batch 8) exceed the 126 MB L2", "parallelism": f"dp{world}",
                                           "precision_mode": PRECISION_NOTE[a.precision]},
           "e2e": {"value": e2e_v, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / a.steps},
           "gpu_launches": rec.count, "clocks": clocks, "roofline": roof, "kernels": kernels}
    if ms_e2e_h:
        res["e2e_bf16_features"] = {"value": world * B * a.steps / (ms_e2e_h * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": h2d_h,
                                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e_h / a.steps,
                                    "note": "same pipeline, f8_*/f4_* shipped as bf16 (HostPipeline widens on the device); informative only -- "
                                            "`e2e` above ships every input as fp32"}
    # whole-step tensor-core fraction: algorithmic FLOPs of every 3-D conv (Table A of SURVEY 8a) + concat_feature, over the step
    step_flops = 0 if head else B * (sum(v for k, v in flops.items() if not (a.att_only and not k.startswith(("hourglass_att", "classif_att_"))))
                      + (0 if a.att_only or a.external_cf else 2 * 2 * 9 * (128 * 64 + 64 * 32) * (H // 4) * (W // 4)))
    ach = step_flops / (ms_total / a.steps * 1e-3) / 1e12
    if not head:
        res["tensor_step"] = {"algorithmic_tflop_per_step": round(step_flops / 1e12, 4), "achieved": round(ach, 1), "unit": "TFLOP/s",
                              "frac_of_sustained_peak": round(ach / pk["tf_sust"], 4), "frac_of_burst_peak": round(ach / pk["tf_burst"], 4),
                              "note": "algorithmic FLOPs only: the bf16x3 attention branch executes 3x its share on the tensor cores"}
    res["host"] = {"numa_bound_cpus": len(numa_cpus) if numa_cpus else None}
    res["launch_mode"] = {"value_region": "cuda graph replay per step (+ NCCL gather to rank 0 on a side stream)" if graph_used else
                          ("eager" + (f" (graph capture failed: {graph_err})" if graph_err else "")),
                          "eager_instrumented_ms_per_step": ms_eager / a.steps,
                          "eager_instrumented_value": world * B * a.steps / (ms_eager * 1e-3),
                          "note": "kernels[] / roofline come from the instrumented eager pass of the same K steps; gpu_launches counts "
                                  "its launches (a graph replay launches the same kernels)"}
    if world == 1 and not a.no_cpu_baseline:
        os.sched_setaffinity(0, all_cpus)           # the CPU baseline gets every host core again
        times = cpu_reference(H, W, md, a.cpu_steps, 1, signed, a.att_only, a.external_cf, a.stage)
        res["cpu_baseline"] = {"value": len(times) / sum(times), "unit": "pairs/s", "cores": torch.get_num_threads(), "kind": "port",
                               "sample": f"{len(times)} pairs at {H}x{W} through {'oracle/backbone.py (HF MobileViTV2) + ' if full else ''}{'oracle/decoder.py + ' if head else ''}oracle/hotpath.py (torch CPU fp32), 1 warm-up"}
    # The default line also carries the whole model from the IMAGES (backbone + decoder + this path: `--stage full`), measured by a
    # short run of this same script in a child process: with the host boundary at the images a step ships 201 MB instead of 1.2 GB,
    # so its end-to-end rate is not PCIe-bound ("report --stage full beside --stage path").
    if (world == 1 and a.stage == "path" and a.precision == "split" and not a.att_only and not a.external_cf and not a.no_full_model
            and a.variant == "us3d" and a.batch == 8 and (H, W) == (1024, 1024)):
        import subprocess
        try:
            torch.cuda.empty_cache()
            cp = subprocess.run([sys.executable, os.path.abspath(__file__), "--stage", "full", "--steps", "5", "--warmup", "3", "--no-cpu-baseline",
                                 "--batch", str(a.batch)], capture_output=True, text=True, timeout=240)
            line = [ln for ln in cp.stdout.splitlines() if ln.startswith("{")][-1]
            fm = json.loads(line)
            res["full_model"] = {"value": fm["value"], "unit": fm["unit"], "ms_per_step": fm["ms_per_step"], "e2e": fm["e2e"],
                                 "dtype": fm["dtype"], "clocks": fm["clocks"], "workload": fm["config"]["workload"],
                                 "note": "python bench.py --stage full --steps 5 --warmup 3 (child process, same GPU, after the path run)"}
        except Exception as e:                          # informative only: never let it break the contract line
            res["full_model"] = {"error": f"{type(e).__name__}: {e}"[:200]}
    print(json.dumps(res))
    if world > 1:
        tdist.destroy_process_group()


if __name__ == "__main__":
    main()
