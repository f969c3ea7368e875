"""Evaluation harness around the path: what `test_us3d.py` does on the host side -- file-list
datasets, image normalisation, checkpoint loading, disparity / segmentation metrics and their running averages -- so that a
reference checkpoint can be scored end to end through the B200 path (`patch_model`, `StereoHead`) or any callable
`model(left, right) -> ([disp], label)`.

Host-side Python like the reference's own (`test_us3d.py:41-128`, `datasets/us3d_.py`, `datasets/whu_dataset.py`,
`datasets/data_io.py:6-13`, `utils/metrics.py:37-59, 98-213`, `utils/experiment.py:122-219`, `models/loss.py:26-31, 106-119`);
nothing here is on the GPU hot path.  The reference's quirks are kept where they change a reported number (per-image metric
averaging with the "mask too small" skip rule, NaN handling of the meters, the 5-class confusion matrix built from 6-class
predictions) and are called out in the docstrings.  Pinned against the reference's functions by `tests/golden/metrics.npz`
(`oracle/make_golden_metrics.py`).
"""
from __future__ import annotations

import argparse
import copy
import os
from typing import Callable, Dict, Iterable, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

IMAGENET_MEAN = (0.485, 0.456, 0.406)      # datasets/data_io.py:7-8
IMAGENET_STD = (0.229, 0.224, 0.225)


# ------------------------------------------------------------------------------------------------------------------
# datasets (test-time branch of the reference datasets: no augmentation, whole images)
# ------------------------------------------------------------------------------------------------------------------
def read_all_lines(filename: str) -> List[str]:
    with open(filename) as f:
        return [line.rstrip() for line in f.readlines()]


def normalize_image(img_u8: np.ndarray) -> torch.Tensor:
    """transforms.ToTensor() + Normalize(ImageNet mean/std) (datasets/data_io.py:6-13): (H,W,3) uint8 -> (3,H,W) float32."""
    x = torch.from_numpy(np.array(img_u8, copy=True)).permute(2, 0, 1).float().div(255.0)
    mean = torch.tensor(IMAGENET_MEAN).view(3, 1, 1)
    std = torch.tensor(IMAGENET_STD).view(3, 1, 1)
    return (x - mean) / std


class StereoListDataset(torch.utils.data.Dataset):
    """Us3dDataset / WhuDataset at test time (datasets/us3d_.py:41-66, 193-213; datasets/whu_dataset.py:16-40).
    List file: one sample per line, `left right disparity [label]` relative to `datapath`.  US3D disparities are float TIFFs read
    as they are, WHU disparities are 16-bit PNGs divided by 256 (whu_dataset.py:34-37).  Returns the reference's sample dict
    (without the unused image-gradient entries gx / gy)."""

    def __init__(self, datapath: str, list_filename: str, dataset: str = "us3d"):
        if dataset not in ("us3d", "whu"):
            raise ValueError("dataset must be 'us3d' or 'whu'")
        self.datapath, self.dataset = datapath, dataset
        rows = [ln.split() for ln in read_all_lines(list_filename) if ln.strip()]
        need = 4 if dataset == "us3d" else 3
        if any(len(r) < need for r in rows):
            raise ValueError(f"{list_filename}: every line needs {need} paths")
        self.rows = rows

    def __len__(self):
        return len(self.rows)

    def _open(self, rel):
        from PIL import Image
        return Image.open(os.path.join(self.datapath, rel))

    def __getitem__(self, i):
        r = self.rows[i]
        left = np.asarray(self._open(r[0]).convert("RGB"))
        right = np.asarray(self._open(r[1]).convert("RGB"))
        disp = np.ascontiguousarray(self._open(r[2]), dtype=np.float32)
        if self.dataset == "whu":
            disp = disp / 256.0
        out = {"left": normalize_image(left), "right": normalize_image(right), "disparity": torch.from_numpy(disp),
               "top_pad": 0, "right_pad": 0, "left_filename": r[0]}
        if self.dataset == "us3d":
            out["label"] = torch.from_numpy(np.ascontiguousarray(self._open(r[3]), dtype=np.float32))
        return out


# ------------------------------------------------------------------------------------------------------------------
# disparity metrics (utils/metrics.py:9-59)
# ------------------------------------------------------------------------------------------------------------------
def _per_image(metric: Callable, d_est, d_gt, mask, *args):
    """compute_metric_for_each_image (utils/metrics.py:16-35): the metric of every image on its own, then the mean over the
    images -- skipping images whose valid-mask fraction is below 10 % of their positive-disparity fraction; 0 if all skipped."""
    if not (d_est.dim() == d_gt.dim() == mask.dim() == 3 and d_est.shape == d_gt.shape == mask.shape):
        raise ValueError("metrics take (B,H,W) estimate, ground truth and mask")
    vals = []
    for i in range(d_gt.shape[0]):
        if mask[i].float().mean() / (d_gt[i] > 0).float().mean() < 0.1:
            continue
        vals.append(metric(d_est[i], d_gt[i], mask[i], *args))
    if not vals:
        return torch.tensor(0, dtype=torch.float32, device=d_gt.device)
    return torch.stack(vals).mean()


@torch.no_grad()
def epe_metric(d_est, d_gt, mask):
    return _per_image(lambda e, g, m: (e[m] - g[m]).abs().mean(), d_est, d_gt, mask)


@torch.no_grad()
def d1_metric(d_est, d_gt, mask):
    def f(e, g, m):
        err = (g[m] - e[m]).abs()
        return ((err > 3) & (err / g[m].abs() > 0.05)).float().mean()
    return _per_image(f, d_est, d_gt, mask)


@torch.no_grad()
def thres_metric(d_est, d_gt, mask, thres: float):
    if not isinstance(thres, (int, float)):
        raise TypeError("thres must be a number")
    return _per_image(lambda e, g, m: ((g[m] - e[m]).abs() > thres).float().mean(), d_est, d_gt, mask)


# ------------------------------------------------------------------------------------------------------------------
# segmentation metrics (utils/metrics.py:98-213)
# ------------------------------------------------------------------------------------------------------------------
class SegmentationMetric:
    """Confusion-matrix metrics as the reference computes them.  test_us3d.py:92 builds it with num_classes - 1 = 5 classes
    while the head predicts 6: the matrix is filled from bincount(gt * 5 + argmax) for gt, pred < 5 only, so ground-truth
    label 5 ("ignore") drops out, and a predicted class 5 on gt class g is counted as (g + 1, 0) -- kept, it is what the
    reference reports."""

    def __init__(self, num_class: int):
        self.num_class = num_class
        self.confusion = np.zeros((num_class, num_class))

    def add_batch(self, pred_logits: torch.Tensor, label: torch.Tensor):
        pred = pred_logits.detach().argmax(dim=1).cpu().numpy().astype(np.uint8)
        gt = label.detach().cpu().numpy()[:, : pred.shape[-2], : pred.shape[-1]].astype(np.int64)
        index = (gt * self.num_class + pred).astype("int32").flatten()
        counts = np.bincount(index)
        n = self.num_class * self.num_class
        m = np.zeros(n)
        m[: min(n, len(counts))] = counts[:n]
        self.confusion += m.reshape(self.num_class, self.num_class)

    def pixel_accuracy(self):
        return np.diag(self.confusion).sum() / self.confusion.sum()

    def class_pixel_accuracy(self):
        return np.diag(self.confusion) / self.confusion.sum(axis=1)

    def mean_pixel_accuracy(self):
        return np.nanmean(self.class_pixel_accuracy())

    def iou(self):
        inter = np.diag(self.confusion)
        return inter / (self.confusion.sum(axis=1) + self.confusion.sum(axis=0) - inter)

    def mean_iou(self):
        return np.nanmean(self.iou())

    def reset(self):
        self.confusion[:] = 0


# ------------------------------------------------------------------------------------------------------------------
# test-time losses (models/loss.py:26-31, 33-64, 106-119)
# ------------------------------------------------------------------------------------------------------------------
def model_loss_test(disp_ests, disp_gts, masks):
    return sum(F.l1_loss(e[m], g[m]) for e, g, m in zip(disp_ests[:1], disp_gts, masks))


def dice_loss_multiclass(logits, target, num_classes: int, drop_last: bool = True, eps: float = 1e-6):
    p = F.softmax(logits, dim=1).float()
    t = F.one_hot(target.to(torch.int64), num_classes).permute(0, 3, 1, 2).float()
    if drop_last:
        p, t = p[:, :-1], t[:, :-1]
    p, t = p.flatten(0, 1), t.flatten(0, 1)
    inter = 2 * (p * t).sum(dim=(-1, -2, -3))
    sets = p.sum(dim=(-1, -2, -3)) + t.sum(dim=(-1, -2, -3))
    sets = torch.where(sets == 0, inter, sets)
    return 1 - ((inter + eps) / (sets + eps)).mean()


def model_label_loss(logits, label, num_classes: int, attention_weights_only: bool, ignore: int = 5):
    ce = F.cross_entropy(logits, label.long(), ignore_index=ignore)
    return (ce + dice_loss_multiclass(logits, label, num_classes)) * (1.6 if attention_weights_only else 2.4)


# ------------------------------------------------------------------------------------------------------------------
# running averages (utils/experiment.py:136-219)
# ------------------------------------------------------------------------------------------------------------------
class AverageMeterDict:
    """Sum / number of updates; a NaN contributes 0 to the sum but still counts as an update (utils/experiment.py:141-173)."""

    def __init__(self):
        self.data, self.count = None, 0

    def update(self, x: Dict):
        self.count += 1
        if self.data is None:
            self.data = copy.deepcopy(x)
            return
        for k, v in x.items():
            if isinstance(v, (list, tuple)):
                for i, e in enumerate(v):
                    self.data[k][i] += 0 if np.isnan(e) else e
            else:
                self.data[k] += 0 if np.isnan(v) else v

    def mean(self):
        div = lambda v: v / float(self.count)      # noqa: E731
        return {k: ([div(e) for e in v] if isinstance(v, (list, tuple)) else div(v)) for k, v in (self.data or {}).items()}


class AverageMeterDict2:
    """Per-key mean over the non-NaN updates only (per-class accuracies; utils/experiment.py:175-219)."""

    def __init__(self):
        self.sum, self.n = {}, {}

    def update(self, x: Dict):
        for k, v in x.items():
            for e in (v if isinstance(v, (list, tuple)) else [v]):
                if not np.isnan(e):
                    self.sum[k] = self.sum.get(k, 0.0) + float(e)
                    self.n[k] = self.n.get(k, 0) + 1

    def mean(self):
        return {k: self.sum[k] / self.n[k] for k in self.sum if self.n[k]}


# ------------------------------------------------------------------------------------------------------------------
# checkpoints (test_us3d.py:61-64, main_us3d.py:116-123)
# ------------------------------------------------------------------------------------------------------------------
def load_checkpoint(target: torch.nn.Module, path_or_state, strict: bool = False):
    """Loads a reference checkpoint ({'model': state_dict, ...} saved from nn.DataParallel, keys prefixed 'module.') into
    `target` (the reference model, DisparityHotPath, Decoder2D or StereoHead).  Only keys `target` owns are taken, like the
    reference's filtered load (main_us3d.py:120-123).  Returns (loaded, skipped) key lists."""
    sd = torch.load(path_or_state, map_location="cpu") if isinstance(path_or_state, (str, os.PathLike)) else path_or_state
    if isinstance(sd, dict) and "model" in sd and isinstance(sd["model"], dict):
        sd = sd["model"]
    sd = {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}
    own = target.state_dict()
    take = {k: v for k, v in sd.items() if k in own and tuple(own[k].shape) == tuple(v.shape)}
    missing = [k for k in own if k not in take and not k.endswith("num_batches_tracked")]
    if strict and missing:
        raise KeyError(f"checkpoint lacks {len(missing)} keys of the target, e.g. {missing[:3]}")
    target.load_state_dict(take, strict=False)
    return sorted(take), sorted(k for k in sd if k not in take)


# ------------------------------------------------------------------------------------------------------------------
# the evaluation loop (test_us3d.py:66-128)
# ------------------------------------------------------------------------------------------------------------------
@torch.no_grad()
def evaluate(model: Callable, samples: Iterable[Dict], maxdisp: int, num_classes: int = 6, attention_weights_only: bool = False,
             signed: bool = True, device: Optional[str] = None, log: Optional[Callable[[str], None]] = None):
    """Scores `model(left, right) -> ([disp, ...], label_logits)` (or just a list of disparities when the model has no
    segmentation head) on an iterable of batched sample dicts (a DataLoader over StereoListDataset).  Returns
    (averaged scalars, averaged per-class scalars) with the keys test_us3d.py prints."""
    avg, avg2 = AverageMeterDict(), AverageMeterDict2()
    for it, s in enumerate(samples):
        mv = (lambda t: t.to(device)) if device else (lambda t: t)
        left, right, gt = mv(s["left"]), mv(s["right"]), mv(s["disparity"])
        label = mv(s["label"]) if "label" in s else None
        mask = (gt < maxdisp) & (gt >= (-maxdisp if signed else 0))
        out = model(left, right)
        disp_ests, logits = (out if isinstance(out, tuple) else (out, None))
        disp_loss = model_loss_test(disp_ests, [gt], [mask])
        sc = {"disp_loss": float(disp_loss)}
        sc["EPE"] = [float(epe_metric(d, gt, mask)) for d in disp_ests]
        sc["D1"] = [float(d1_metric(d, gt, mask)) for d in disp_ests]
        sc["Thres1"] = [float(thres_metric(d, gt, mask, 1.0)) for d in disp_ests]
        sc["Thres2"] = [float(thres_metric(d, gt, mask, 2.0)) for d in disp_ests]
        sc2 = {}
        if logits is not None and label is not None:
            metric = SegmentationMetric(num_classes - 1)
            metric.add_batch(logits, label)
            label_loss = float(model_label_loss(logits, label, num_classes, attention_weights_only))
            sc.update(label_loss=label_loss, loss=sc["disp_loss"] + label_loss, PA=[float(metric.pixel_accuracy())],
                      MPA=[float(metric.mean_pixel_accuracy())], mIoU=[float(metric.mean_iou())])
            cpa, iou = metric.class_pixel_accuracy(), metric.iou()
            for c in range(num_classes - 1):
                sc2[f"CPA{c}"] = [float(cpa[c])]
                sc2[f"IoU{c}"] = [float(iou[c])]
        else:
            sc["loss"] = sc["disp_loss"]
        avg.update(sc)
        avg2.update(sc2)
        if log:
            log(f"Iter {it}, test loss = {sc['loss']:.3f}, EPE = {sc['EPE'][0]:.3f}")
    return avg.mean(), avg2.mean()


def main(argv=None):
    """CLI in the shape of test_us3d.py: evaluates a reference checkpoint through the B200 path.  The backbone (timm
    MobileViTv2) is the caller's: pass --reference <path to the SemStereo tree> so that the reference model class can be built
    (its 3-D modules and operators then ARE this package, via models.submodule -> semstereo_b200.submodule) and patch_model
    routes everything after the backbone through the fused kernels."""
    ap = argparse.ArgumentParser(description="SemStereo evaluation through the B200 path")
    ap.add_argument("--model", default="SemStereo", choices=["SemStereo", "SemStereo_WHU"])
    ap.add_argument("--maxdisp", type=int, default=64)
    ap.add_argument("--num_classes", type=int, default=6)
    ap.add_argument("--attention_weights_only", action="store_true")
    ap.add_argument("--dataset", default="us3d", choices=["us3d", "whu"])
    ap.add_argument("--datapath", required=True)
    ap.add_argument("--testlist", required=True)
    ap.add_argument("--test_batch_size", type=int, default=1)
    ap.add_argument("--loadckpt", required=True)
    ap.add_argument("--reference", required=True, help="path of the reference repository (for the model class and its timm backbone)")
    a = ap.parse_args(argv)
    import importlib
    import sys
    import types
    pkg = types.ModuleType("models")
    pkg.__path__ = [os.path.join(a.reference, "models")]
    sys.modules["models"] = pkg
    signed = a.model == "SemStereo"
    sys.modules["models.submodule"] = importlib.import_module("semstereo_b200.submodule" if signed else "semstereo_b200.submodule_")
    sys.modules["models.submodule_other"] = importlib.import_module("semstereo_b200.submodule_other")
    cls = getattr(importlib.import_module("models." + a.model), a.model)
    model = cls(a.maxdisp, a.attention_weights_only, True, True, a.num_classes)
    loaded, skipped = load_checkpoint(model, a.loadckpt)
    print(f"loaded {len(loaded)} tensors, skipped {len(skipped)}")
    model = model.cuda().eval()
    from .patch import patch_model
    patch_model(model, signed=signed)
    ds = StereoListDataset(a.datapath, a.testlist, a.dataset)
    loader = torch.utils.data.DataLoader(ds, a.test_batch_size, shuffle=False, num_workers=4, drop_last=False)
    scalars, per_class = evaluate(model, loader, a.maxdisp, a.num_classes, a.attention_weights_only, signed, device="cuda", log=print)
    print("avg_test_scalars", scalars, "avg_test_scalars2", per_class)


if __name__ == "__main__":
    main()
