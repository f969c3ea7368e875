"""torch.library registration of the cost-volume / regression operators ("registers each op with
torch.library"): `torch.ops.semstereo_b200.<op>` with shape-only fake implementations, so the operators can sit inside
`torch.export` / `torch.compile`d callers of the reference model without graph breaks.  The real implementations are the C-ABI
kernels (CUDA only — a CPU tensor raises, there is no fallback).  Importing this module performs the registration."""
from __future__ import annotations

import torch

from . import ops

NS = "semstereo_b200"


@torch.library.custom_op(f"{NS}::gwc_volume", mutates_args=())
def gwc_volume(left: torch.Tensor, right: torch.Tensor, maxdisp: int, num_groups: int, signed: bool, norm: bool) -> torch.Tensor:
    """build_gwc_volume / build_gwc_volume_norm (models/submodule.py:198-238, submodule_.py:188-221)."""
    return ops.gwc_volume(left.contiguous(), right.contiguous(), maxdisp, num_groups, signed, norm)


@gwc_volume.register_fake
def _(left, right, maxdisp, num_groups, signed, norm):
    B, C, H, W = left.shape
    return left.new_empty((B, num_groups, 2 * maxdisp if signed else maxdisp, H, W))


@torch.library.custom_op(f"{NS}::concat_volume", mutates_args=())
def concat_volume(left: torch.Tensor, right: torch.Tensor, maxdisp: int, signed: bool) -> torch.Tensor:
    """build_concat_volume (models/submodule.py:173-187, submodule_.py:166-178)."""
    return ops.concat_volume(left.contiguous(), right.contiguous(), maxdisp, signed)


@concat_volume.register_fake
def _(left, right, maxdisp, signed):
    B, C, H, W = left.shape
    return left.new_empty((B, 2 * C, 2 * maxdisp if signed else maxdisp, H, W))


@torch.library.custom_op(f"{NS}::regression_topk", mutates_args=())
def regression_topk(cost: torch.Tensor, disparity_samples: torch.Tensor, k: int) -> torch.Tensor:
    """regression_topk (models/submodule.py:434-442)."""
    return ops.regression_topk(cost.contiguous(), disparity_samples.contiguous(), k)


@regression_topk.register_fake
def _(cost, disparity_samples, k):
    B, D, H, W = cost.shape
    return cost.new_empty((B, 1, H, W))


@torch.library.custom_op(f"{NS}::context_upsample", mutates_args=())
def context_upsample(depth_low: torch.Tensor, up_weights: torch.Tensor) -> torch.Tensor:
    """context_upsample (models/submodule_.py:311-323)."""
    return ops.context_upsample(depth_low.contiguous(), up_weights.contiguous())


@context_upsample.register_fake
def _(depth_low, up_weights):
    B, _, h, w = depth_low.shape
    return depth_low.new_empty((B, 4 * h, 4 * w))


@torch.library.custom_op(f"{NS}::disparity_regression", mutates_args=())
def disparity_regression(prob: torch.Tensor, maxdisp: int, signed: bool) -> torch.Tensor:
    """disparity_regression (models/submodule.py:164-170, submodule_.py:159-163)."""
    return ops.disparity_regression(prob.contiguous(), float(-maxdisp if signed else 0))


@disparity_regression.register_fake
def _(prob, maxdisp, signed):
    B, D, H, W = prob.shape
    return prob.new_empty((B, H, W))


# ---- autograd: vector-Jacobian products from csrc/backward.cu (BASELINE config #5, first step) -------------------------
def _gwc_setup(ctx, inputs, output):
    left, right, maxdisp, num_groups, signed, norm = inputs
    ctx.save_for_backward(left, right)
    ctx.args = (maxdisp, num_groups, signed, norm)


def _gwc_backward(ctx, grad):
    left, right = ctx.saved_tensors
    gl, gr = ops.gwc_volume_backward(left.contiguous(), right.contiguous(), grad.contiguous(), *ctx.args)
    return gl, gr, None, None, None, None


torch.library.register_autograd(f"{NS}::gwc_volume", _gwc_backward, setup_context=_gwc_setup)


def _concat_setup(ctx, inputs, output):
    ctx.args = (inputs[2], inputs[3])


def _concat_backward(ctx, grad):
    gl, gr = ops.concat_volume_backward(grad.contiguous(), *ctx.args)
    return gl, gr, None, None


torch.library.register_autograd(f"{NS}::concat_volume", _concat_backward, setup_context=_concat_setup)


def _topk_setup(ctx, inputs, output):
    cost, samples, k = inputs
    ctx.save_for_backward(cost, samples)
    ctx.k = k


def _topk_backward(ctx, grad):
    cost, samples = ctx.saved_tensors
    gc, gs = ops.regression_topk_backward(cost.contiguous(), samples.contiguous(), grad.contiguous(), ctx.k)
    return gc, gs, None


torch.library.register_autograd(f"{NS}::regression_topk", _topk_backward, setup_context=_topk_setup)


def _ctx_up_setup(ctx, inputs, output):
    ctx.save_for_backward(*inputs)


def _ctx_up_backward(ctx, grad):
    depth_low, up_weights = ctx.saved_tensors
    return ops.context_upsample_backward(depth_low.contiguous(), up_weights.contiguous(), grad.contiguous())


torch.library.register_autograd(f"{NS}::context_upsample", _ctx_up_backward, setup_context=_ctx_up_setup)


def _dreg_setup(ctx, inputs, output):
    prob, maxdisp, signed = inputs
    ctx.D, ctx.dmin = prob.shape[1], float(-maxdisp if signed else 0)


def _dreg_backward(ctx, grad):
    return ops.disparity_regression_backward(grad.contiguous(), ctx.D, ctx.dmin), None, None


torch.library.register_autograd(f"{NS}::disparity_regression", _dreg_backward, setup_context=_dreg_setup)


# ---- the remaining free functions / stateless modules of the surface, with autograd ----------------------------------------
@torch.library.custom_op(f"{NS}::propagation", mutates_args=())
def propagation(x: torch.Tensor) -> torch.Tensor:
    """Propagation / Propagation_prob (models/submodule.py:290-307, 361-377): (B,1,[D,]H,W) -> (B,5,[D,]H,W)."""
    return ops.propagation(x.contiguous())


@propagation.register_fake
def _(x):
    return x.new_empty((x.shape[0], 5) + tuple(x.shape[2:]))


torch.library.register_autograd(f"{NS}::propagation", lambda ctx, g: ops.propagation_backward(g.contiguous()))


@torch.library.custom_op(f"{NS}::disparity_variance", mutates_args=())
def disparity_variance(prob: torch.Tensor, maxdisp: int, disparity: torch.Tensor, signed: bool) -> torch.Tensor:
    """disparity_variance (models/submodule.py:257-263)."""
    return ops.disparity_variance(prob.contiguous(), disparity.contiguous(), float(-maxdisp if signed else 0))


@disparity_variance.register_fake
def _(prob, maxdisp, disparity, signed):
    B, D, H, W = prob.shape
    return prob.new_empty((B, 1, H, W))


def _dvar_setup(ctx, inputs, output):
    prob, maxdisp, disparity, signed = inputs
    ctx.save_for_backward(prob, disparity)
    ctx.dmin = float(-maxdisp if signed else 0)


def _dvar_backward(ctx, grad):
    prob, disparity = ctx.saved_tensors
    gp, gm = ops.disparity_variance_backward(prob.contiguous(), disparity.contiguous(), grad.contiguous(), ctx.dmin)
    return gp, None, gm.view_as(disparity), None


torch.library.register_autograd(f"{NS}::disparity_variance", _dvar_backward, setup_context=_dvar_setup)


@torch.library.custom_op(f"{NS}::spatial_transformer_grid", mutates_args=())
def spatial_transformer_grid(x: torch.Tensor, y: torch.Tensor, disp_range_samples: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """SpatialTransformer_grid (models/submodule.py:265-288): returns (y_warped, x_repeated), both (B,C,K,H,W)."""
    yw, xr = ops.spatial_transformer_grid(x.contiguous(), y.contiguous(), disp_range_samples.contiguous())
    return yw, xr


@spatial_transformer_grid.register_fake
def _(x, y, disp_range_samples):
    B, C, H, W = y.shape
    K = disp_range_samples.shape[1]
    return y.new_empty((B, C, K, H, W)), y.new_empty((B, C, K, H, W))


def _stn_setup(ctx, inputs, output):
    x, y, disp = inputs
    ctx.save_for_backward(y, disp)


def _stn_backward(ctx, g_yw, g_xr):
    y, disp = ctx.saved_tensors
    gx, gy, gd = ops.spatial_transformer_grid_backward(y.contiguous(), disp.contiguous(), g_yw.contiguous(), g_xr.contiguous())
    return gx, gy, gd


torch.library.register_autograd(f"{NS}::spatial_transformer_grid", _stn_backward, setup_context=_stn_setup)
