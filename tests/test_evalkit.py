"""Evaluation harness: metrics / losses / meters against values recorded from the reference's own functions
(tests/golden/metrics.npz, oracle/make_golden_metrics.py), the file-list datasets on small synthetic files, the checkpoint loader
with the DataParallel prefix, and the evaluate() loop on a stand-in model.  CPU only."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
from semstereo_b200 import evalkit as ek


@pytest.fixture
def gold(golden_dir):
    return dict(np.load(os.path.join(golden_dir, "metrics.npz")))


def make():
    from make_golden_metrics import inputs
    return inputs()


def test_disparity_metrics_match_the_reference(gold):
    gt, est, label, logits = make()
    mask = (gt < 64) & (gt >= -64)
    assert abs(float(ek.epe_metric(est, gt, mask)) - gold["EPE"]) <= 1e-6
    assert abs(float(ek.d1_metric(est, gt, mask)) - gold["D1"]) <= 1e-7
    assert abs(float(ek.thres_metric(est, gt, mask, 1.0)) - gold["Thres1"]) <= 1e-7
    assert abs(float(ek.thres_metric(est, gt, mask, 2.0)) - gold["Thres2"]) <= 1e-7
    assert abs(float(ek.model_loss_test([est], [gt], [mask])) - gold["disp_loss"]) <= 1e-6
    none = torch.zeros_like(mask)
    assert float(ek.epe_metric(est, gt, none)) == 0.0          # every image skipped -> 0 (utils/metrics.py:31-33)
    with pytest.raises(ValueError):
        ek.epe_metric(est[0], gt[0], mask[0])


def test_segmentation_metrics_match_the_reference(gold):
    gt, est, label, logits = make()
    m = ek.SegmentationMetric(5)
    m.add_batch(logits, label)
    assert np.array_equal(m.confusion, gold["confusion"])
    assert abs(m.pixel_accuracy() - gold["PA"]) <= 1e-12 and abs(m.mean_pixel_accuracy() - gold["MPA"]) <= 1e-12
    assert abs(m.mean_iou() - gold["mIoU"]) <= 1e-12
    assert np.allclose(m.class_pixel_accuracy(), gold["CPA"], atol=1e-12) and np.allclose(m.iou(), gold["IoU"], atol=1e-12)
    assert abs(float(ek.model_label_loss(logits, label, 6, False)) - gold["label_loss"]) <= 2e-6
    assert abs(float(ek.model_label_loss(logits, label, 6, True)) - gold["label_loss_att"]) <= 2e-6


def test_meters_match_the_reference(gold):
    a1, a2 = ek.AverageMeterDict(), ek.AverageMeterDict2()
    for v in (1.0, float("nan"), 4.0):
        a1.update({"x": v, "l": [v, 2.0]})
        a2.update({"c": [v]})
    assert abs(a1.mean()["x"] - gold["meter_x"]) <= 1e-12 and np.allclose(a1.mean()["l"], gold["meter_l"])
    assert abs(a2.mean()["c"] - gold["meter2_c"]) <= 1e-12


def _write_samples(root, n, dataset):
    from PIL import Image
    rng = np.random.default_rng(0)
    lines = []
    for i in range(n):
        left = rng.integers(0, 256, (32, 48, 3), dtype=np.uint8)
        right = np.roll(left, -3, axis=1)
        Image.fromarray(left).save(root / f"l{i}.png")
        Image.fromarray(right).save(root / f"r{i}.png")
        if dataset == "us3d":
            Image.fromarray(np.full((32, 48), 3.0 + i, dtype=np.float32)).save(root / f"d{i}.tif")
            Image.fromarray(rng.integers(0, 6, (32, 48), dtype=np.uint8)).save(root / f"c{i}.tif")
            lines.append(f"l{i}.png r{i}.png d{i}.tif c{i}.tif")
        else:
            Image.fromarray(np.full((32, 48), (3 + i) * 256, dtype=np.uint16)).save(root / f"d{i}.png")
            lines.append(f"l{i}.png r{i}.png d{i}.png")
    (root / "list.txt").write_text("\n".join(lines) + "\n")
    return left


@pytest.mark.parametrize("dataset", ["us3d", "whu"])
def test_list_dataset_and_evaluate_loop(tmp_path, dataset):
    last_left = _write_samples(tmp_path, 3, dataset)
    ds = ek.StereoListDataset(str(tmp_path), str(tmp_path / "list.txt"), dataset)
    assert len(ds) == 3
    s = ds[2]
    assert tuple(s["left"].shape) == (3, 32, 48) and s["left"].dtype == torch.float32
    want = (torch.from_numpy(last_left).permute(2, 0, 1).float() / 255 - torch.tensor(ek.IMAGENET_MEAN).view(3, 1, 1)) / torch.tensor(ek.IMAGENET_STD).view(3, 1, 1)
    assert torch.allclose(s["left"], want, atol=1e-6)
    assert float(s["disparity"][0, 0]) == 5.0                      # 3 + i; WHU stored as uint16 * 256
    assert ("label" in s) == (dataset == "us3d")
    loader = torch.utils.data.DataLoader(ds, 2, shuffle=False)

    def model(left, right):                                       # a perfect disparity off by 0.5 px, random-ish labels
        b = left.shape[0]
        disp = torch.stack([torch.full((32, 48), 3.5 + i) for i in range(model.seen, model.seen + b)])
        model.seen += b
        logits = torch.zeros(b, 6, 32, 48)
        logits[:, 1] = 1.0
        return ([disp], logits) if dataset == "us3d" else [disp]
    model.seen = 0
    scalars, per_class = ek.evaluate(model, loader, 64, signed=(dataset == "us3d"))
    assert abs(scalars["EPE"][0] - 0.5) <= 1e-6 and scalars["D1"][0] == 0.0 and scalars["Thres1"][0] == 0.0
    if dataset == "us3d":
        assert 0.0 < scalars["PA"][0] < 1.0 and "IoU1" in per_class and scalars["loss"] > scalars["disp_loss"]
    with pytest.raises(ValueError):
        ek.StereoListDataset(str(tmp_path), str(tmp_path / "list.txt"), "kitti")


def test_checkpoint_loader_takes_a_dataparallel_checkpoint(tmp_path):
    from semstereo_b200.hotpath import DisparityHotPath
    from semstereo_b200.params import make_params
    p = make_params(seed=3)
    ck = {"model": {"module." + k: v for k, v in p.items()}, "epoch": 7}
    ck["model"]["module.feature.conv_stem.weight"] = torch.zeros(32, 3, 3, 3)      # a key outside the path: skipped
    torch.save(ck, tmp_path / "checkpoint_000007.ckpt")
    m = DisparityHotPath(64)
    loaded, skipped = ek.load_checkpoint(m, str(tmp_path / "checkpoint_000007.ckpt"), strict=True)
    assert len(loaded) == len(p) and skipped == ["feature.conv_stem.weight"]
    assert torch.equal(m.state_dict()["classif.2.weight"], p["classif.2.weight"])
    with pytest.raises(KeyError):
        ek.load_checkpoint(DisparityHotPath(64), {"model": {"gamma": torch.zeros(1)}}, strict=True)
