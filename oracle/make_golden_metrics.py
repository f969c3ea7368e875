"""oracle/make_golden_metrics.py -- TEST INFRASTRUCTURE ONLY.  Generates tests/golden/metrics.npz by running the UNMODIFIED
reference metric / loss functions (utils/metrics.py, utils/experiment.py, models/loss.py, imported by path from /root/reference)
on seeded tensors.  Run in the authoring container: python oracle/make_golden_metrics.py
(numpy >= 1.24 removed np.int, which utils/metrics.py:151 uses: it is aliased to int for the import, nothing else is touched)."""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("SEMSTEREO_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def inputs(seed=0, B=3, H=64, W=96):
    g = torch.Generator().manual_seed(seed)
    gt = (torch.rand(B, H, W, generator=g) * 160 - 80)
    gt[2] = gt[2].abs() + 70           # image 2: mostly outside [-64, 64) -> the "mask too small" skip rule fires
    est = gt + torch.randn(B, H, W, generator=g) * torch.rand(B, 1, 1, generator=g) * 6
    label = torch.randint(0, 6, (B, H, W), generator=g).float()
    logits = torch.randn(B, 6, H, W, generator=g) * 2
    hit = torch.rand(B, H, W, generator=g) < 0.6          # make the prediction mostly right
    logits.scatter_add_(1, label.long().unsqueeze(1), hit.float().unsqueeze(1) * 6)
    return gt, est, label, logits


def main():
    if not hasattr(np, "int"):
        np.int = int
    pkg = types.ModuleType("utils")
    pkg.__path__ = [os.path.join(REF, "utils")]
    sys.modules["utils"] = pkg
    load("utils.experiment", os.path.join(REF, "utils", "experiment.py"))
    m = load("utils.metrics", os.path.join(REF, "utils", "metrics.py"))
    loss = load("ref_loss", os.path.join(REF, "models", "loss.py"))
    gt, est, label, logits = inputs()
    mask = (gt < 64) & (gt >= -64)
    out = {"EPE": m.EPE_metric(est, gt, mask), "D1": m.D1_metric(est, gt, mask), "Thres1": m.Thres_metric(est, gt, mask, 1.0),
           "Thres2": m.Thres_metric(est, gt, mask, 2.0)}
    seg = m.SegmentationMetric(5)
    seg.addBatch(logits, label)
    out.update(confusion=seg.confusionMatrix, PA=seg.pixelAccuracy(), MPA=seg.meanPixelAccuracy(), mIoU=seg.meanIntersectionOverUnion(),
               CPA=seg.classPixelAccuracy(), IoU=seg.IoU())
    out["disp_loss"] = loss.model_loss_test([est], [gt], [mask])
    out["label_loss"] = loss.model_label_loss(logits, label, 6, False)
    out["label_loss_att"] = loss.model_label_loss(logits, label, 6, True)
    # meters: three updates incl. a NaN
    exp = sys.modules["utils.experiment"]
    a1, a2 = exp.AverageMeterDict(), exp.AverageMeterDict2()
    for v in (1.0, float("nan"), 4.0):
        a1.update({"x": v, "l": [v, 2.0]})
        a2.update({"c": [v]})
    out["meter_x"], out["meter_l"], out["meter2_c"] = a1.mean()["x"], np.array(a1.mean()["l"]), a2.mean()["c"]
    np.savez(os.path.join(ROOT, "tests", "golden", "metrics.npz"), **{k: np.asarray(v, dtype=np.float64) for k, v in out.items()})
    print({k: np.asarray(v).round(5).tolist() for k, v in out.items() if np.asarray(v).size < 8})


if __name__ == "__main__":
    main()
