"""Training closure (BASELINE config #5: SemStereo attention_weights_only forward + backward, batch
16 over 8 GPUs): the reference losses and the data-parallel gradient reduction (the kernels' autograd is train_ops.py).

A user trains THE REFERENCE MODEL OBJECT through the level-1 drop-in: with `semstereo_b200.submodule` /
`submodule_other` installed, every module and function `SemStereo.py` star-imports is differentiable on the CUDA kernels
(train_ops.py, torch_ops.py), and `loss.backward()` (main_us3d.py:220) just works.  The reference tree is not on the GPU box, so
`SemStereoTrainGlue` restates the model's `__init__` / `forward` (models/SemStereo.py:184-346, training-mode returns :329-337) the
way the reference writes them; it lives in tools/train_glue.py (a stand-in for the reference model file, not product code).
This module holds what the training step needs besides the model: the losses and the gradient all-reduce.

Losses: models/loss.py:19-31 (smooth-L1 pyramid), :106-119 (cross entropy + dice), :121-135 (LRSC: left labels warped by the
predicted disparity supervise the right segmentation head), combined as main_us3d.py:204-208.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------------------------
# losses
# ------------------------------------------------------------------------------------------------------------------
def model_loss_train(disp_ests, disp_gts, masks):
    """models/loss.py:19-24: weighted smooth-L1 over the prediction pyramid (weights 1.0, 0.6, 0.5, 0.3)."""
    return sum(w * F.smooth_l1_loss(e[m], g[m]) for e, g, w, m in zip(disp_ests, disp_gts, (1.0, 0.6, 0.5, 0.3), masks))


def model_label_loss(logits, label, num_classes, attention_weights_only, ignore=5):
    """models/loss.py:106-119 (same restatement as evalkit.model_label_loss, differentiable)."""
    from .evalkit import model_label_loss as f
    return f(logits, label, num_classes, attention_weights_only, ignore)


def lrsc_loss(label_est_r, disp_ests, label):
    """LRSC_loss (models/loss.py:121-135): label_r[y, x] = label[y, clamp(x - disp[y, x])] (integer gather), cross entropy on the
    right head, ignore index -1."""
    b, h, w = label.shape
    xs = torch.arange(w, device=label.device).view(1, 1, w).expand(b, h, w)
    src = torch.clamp(xs - disp_ests[0], min=0, max=w - 1).long()
    warped = torch.gather(label, 2, src)
    return F.cross_entropy(label_est_r, warped.long(), ignore_index=-1)


def total_loss(outputs, disp_gt, disp_gt_4, label, maxdisp, num_classes=6, attention_weights_only=True):
    """main_us3d.py:196-208."""
    disp_ests, label_est, label_est_r = outputs
    mask = (disp_gt < maxdisp) & (disp_gt >= -maxdisp)
    mask_4 = (disp_gt_4 < maxdisp) & (disp_gt_4 >= -maxdisp)
    disp = model_loss_train(disp_ests, [disp_gt, disp_gt_4, disp_gt, disp_gt_4], [mask, mask_4, mask, mask_4])
    lab = model_label_loss(label_est, label, num_classes, attention_weights_only)
    lrsc = lrsc_loss(label_est_r, disp_ests, label)
    return disp + lab + lrsc, {"disp_loss": disp.detach(), "label_loss": lab.detach(), "lrsc_loss": lrsc.detach()}


def allreduce_gradients(model: nn.Module, world: int):
    """Data-parallel gradient averaging, one flattened NCCL all-reduce (what DistributedDataParallel / DataParallel's backward
    reduction amounts to for this model: ~35 M parameters, 140 MB fp32)."""
    import torch.distributed as dist
    grads = [p.grad for p in model.parameters() if p.grad is not None]
    if world <= 1 or not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat)
    flat.div_(world)
    o = 0
    for g in grads:
        g.copy_(flat[o:o + g.numel()].view_as(g))
        o += g.numel()
