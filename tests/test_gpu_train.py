"""Training closure of the 3-D conv stack (BASELINE config #5): Conv3d (forward, dX, dW) and BatchNorm3d with
batch statistics (forward, backward, running-stat update) on the CUDA kernels against torch autograd in true fp32 (TF32 off);
tolerance 1e-4 relative; the surface modules convbn_3d / BasicConv(is_3d) in training mode."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

if torch.cuda.is_available():
    from semstereo_b200 import train_ops


@pytest.fixture(autouse=True)
def true_fp32_torch_convs():
    tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32


def relerr(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


@pytest.mark.parametrize("Cin,Cout,k,stride,B,D,H,W", [(32, 64, 3, 2, 2, 8, 16, 16), (64, 64, 3, 1, 1, 4, 16, 24), (128, 128, 3, 1, 2, 4, 8, 8),
                                                       (32, 32, 1, 1, 2, 6, 10, 12), (64, 128, 3, 2, 1, 4, 8, 16), (32, 32, 3, 1, 1, 6, 20, 36)])
def test_conv3d_forward_and_gradients(Cin, Cout, k, stride, B, D, H, W):
    g = torch.Generator().manual_seed(Cin + Cout + k + stride)
    x = torch.randn(B, Cin, D, H, W, generator=g).to(DEV).requires_grad_(True)
    w = (torch.randn(Cout, Cin, k, k, k, generator=g) / (Cin * k ** 3) ** 0.5).to(DEV).requires_grad_(True)
    y = train_ops.conv3d(x, w, stride)
    ref = F.conv3d(x, w, None, stride, k // 2)
    assert relerr(y, ref) <= 1e-5
    gy = torch.randn(ref.shape, generator=g).to(DEV)
    gx, gw = torch.autograd.grad(y, (x, w), gy)
    rx, rw = torch.autograd.grad(ref, (x, w), gy)
    assert relerr(gx, rx) <= 1e-4 and relerr(gw, rw) <= 1e-4, (relerr(gx, rx), relerr(gw, rw))


@pytest.mark.parametrize("shape", [(2, 32, 4, 16, 16), (3, 6, 1, 40, 24), (2, 64, 8, 8, 8)])
def test_batchnorm_with_batch_statistics(shape):
    g = torch.Generator().manual_seed(shape[1])
    C = shape[1]
    x = (torch.randn(shape, generator=g) * 2 + 0.7).to(DEV).requires_grad_(True)
    bn = (nn.BatchNorm3d(C) if True else None).to(DEV).train()
    with torch.no_grad():
        bn.weight.copy_(torch.rand(C, generator=g) + 0.5)
        bn.bias.copy_(0.2 * torch.randn(C, generator=g))
    ref_bn = nn.BatchNorm3d(C).to(DEV).train()
    ref_bn.load_state_dict(bn.state_dict())
    y = train_ops.batch_norm_train(x, bn)
    ref = ref_bn(x)
    assert relerr(y, ref) <= 1e-5
    assert relerr(bn.running_mean, ref_bn.running_mean) <= 1e-5 and relerr(bn.running_var, ref_bn.running_var) <= 1e-5
    assert int(bn.num_batches_tracked) == 1
    gy = torch.randn(shape, generator=g).to(DEV)
    gx, gw, gb = torch.autograd.grad(y, (x, bn.weight, bn.bias), gy)
    rx, rw, rb = torch.autograd.grad(ref, (x, ref_bn.weight, ref_bn.bias), gy)
    assert relerr(gx, rx) <= 1e-4 and relerr(gw, rw) <= 1e-4 and relerr(gb, rb) <= 1e-4


def test_surface_modules_train():
    """convbn_3d / BasicConv(is_3d) of the surface in training mode == the torch modules they replace, incl. parameter gradients."""
    import semstereo_b200.submodule as sub
    import semstereo_b200.submodule_other as other
    torch.manual_seed(0)
    m = other.convbn_3d(32, 64, 3, 2, 1).to(DEV).train()
    t = nn.Sequential(nn.Conv3d(32, 64, 3, 2, 1, bias=False), nn.BatchNorm3d(64)).to(DEV).train()
    t.load_state_dict(m.state_dict())
    x = torch.randn(2, 32, 8, 16, 16, device=DEV, requires_grad=True)
    y, r = m(x), t(x)
    assert relerr(y, r) <= 1e-5
    (y.relu().square().mean()).backward()
    gm = [p.grad.clone() for p in m.parameters()] + [x.grad.clone()]
    x.grad = None
    (r.relu().square().mean()).backward()
    gt = [p.grad for p in t.parameters()] + [x.grad]
    for a, b in zip(gm, gt):
        assert relerr(a, b) <= 1e-4
    bc = sub.BasicConv(64, 32, is_3d=True, kernel_size=3, stride=1, padding=1).to(DEV).train()
    x2 = torch.randn(1, 64, 4, 16, 16, device=DEV, requires_grad=True)
    ref = F.relu(F.batch_norm(F.conv3d(x2, bc.conv.weight, None, 1, 1), None, None, bc.bn.weight, bc.bn.bias, True, 0.1, 1e-5))
    assert relerr(bc(x2), ref) <= 1e-5
    # eval mode afterwards uses the running statistics the training steps updated
    m.eval()
    with torch.no_grad():
        e = m(x.detach())
    t.eval()
    assert relerr(e, t(x.detach())) <= 1e-4


# (4, 8, 10): W padded to the window (nothing masked); (4, 6, 10) and (6, 6, 6): H and W both padded, the masked branch
@pytest.mark.parametrize("D,H,W,block", [(4, 8, 8, (4, 4, 4)), (6, 8, 12, (6, 4, 4)), (4, 8, 10, (4, 4, 4)), (4, 6, 10, (4, 4, 4)), (6, 6, 6, (6, 4, 4))])
def test_attention_block_train(D, H, W, block):
    """attention_block in training mode (k = 1 convs + softmax core kernels, forward and backward) vs torch autograd through the oracle."""
    from oracle import ops as oo
    import semstereo_b200.submodule_other as other
    torch.manual_seed(1)
    m = other.attention_block(128, 16, block).to(DEV).train()
    x = torch.randn(2, 128, D, H, W, device=DEV, requires_grad=True)
    y = m(x)
    p = {"a.qkv_3d.weight": m.qkv_3d.weight, "a.qkv_3d.bias": m.qkv_3d.bias, "a.final1x1.weight": m.final1x1.weight, "a.final1x1.bias": m.final1x1.bias}
    ref = oo.window_attention3d(x, p, "a", 16, block)
    assert relerr(y, ref) <= 1e-5
    gy = torch.randn_like(ref)
    params = [x, m.qkv_3d.weight, m.qkv_3d.bias, m.final1x1.weight, m.final1x1.bias]
    got = torch.autograd.grad(y, params, gy)
    want = torch.autograd.grad(ref, params, gy)
    for a, b in zip(got, want):
        assert relerr(a, b) <= 1e-4, relerr(a, b)


def test_ssr_upsample_train():
    """SSR_upsample in training mode (batch statistics in its four BatchNorm2d layers) vs the torch formulation of the reference."""
    import semstereo_b200.submodule as sub
    torch.manual_seed(2)
    m = sub.SSR_upsample(6).to(DEV).train()
    with torch.no_grad():
        for p_ in m.parameters():
            p_.add_(0.1 * torch.randn_like(p_))
    t = sub.SSR_upsample(6).to(DEV).train()
    t.load_state_dict(m.state_dict())
    d = (torch.randn(2, 1, 16, 24, device=DEV) * 4).requires_grad_(True)
    spx = torch.randn(2, 6, 64, 96, device=DEV, requires_grad=True)
    lab = torch.randn(2, 6, 64, 96, device=DEV, requires_grad=True)

    def torch_forward(mod, depth_low, weights, pred_label):       # submodule.py:421-431 with torch modules
        bn = lambda x, b: F.batch_norm(x, b.running_mean, b.running_var, b.weight, b.bias, True, b.momentum, b.eps)   # noqa: E731
        pl = F.softmax(pred_label, dim=1)
        dup = F.interpolate(depth_low, scale_factor=4, mode="bilinear", align_corners=False)
        x = bn(F.conv2d(bn(dup, mod.conv[0]), mod.conv[1].weight, mod.conv[1].bias, padding=1), mod.conv[2])
        g = torch.sigmoid(bn(F.conv2d(pl * weights, mod.conv1[0].weight, mod.conv1[0].bias), mod.conv1[1]))
        g = torch.sigmoid(bn(F.conv2d(g * weights, mod.conv2[0].weight, mod.conv2[0].bias), mod.conv2[1]))
        return (dup + F.conv2d(x * g, mod.conv3.weight, mod.conv3.bias)).squeeze(1)

    y, ref = m(d, spx, lab), torch_forward(t, d, spx, lab)
    assert relerr(y, ref) <= 2e-5
    assert relerr(m.conv[2].running_var, t.conv[2].running_var) <= 1e-5
    gy = torch.randn_like(ref)
    got = torch.autograd.grad(y, [d, spx, lab] + list(m.parameters()), gy)
    want = torch.autograd.grad(ref, [d, spx, lab] + list(t.parameters()), gy)
    for a, b in zip(got, want):
        if b.abs().max().item() < 1e-3:      # biases in front of a training-mode BatchNorm: the exact gradient is 0, both sides are rounding noise
            assert (a - b).abs().max().item() <= 1e-3
        else:
            assert relerr(a, b) <= 2e-4, relerr(a, b)


def test_training_step_of_the_model_glue():
    """One attention_weights_only training step of the reference-style glue (tools/train_glue.py) at 128x128: finite loss, gradients
    on every parameter that the forward uses, and the gradients of the kernel-backed modules equal torch autograd through torch
    re-implementations of those modules (same glue, torch modules swapped in)."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))
    from train_glue import SemStereoTrainGlue, synthetic_batch
    from semstereo_b200 import train as T
    torch.manual_seed(3)
    model = SemStereoTrainGlue(64).to(DEV).train()
    fl, fr, disp, disp4, label = synthetic_batch(5, 2, 128, 128, DEV)
    loss, parts = T.total_loss(model(fl, fr), disp, disp4, label, 64)
    loss.backward()
    assert torch.isfinite(loss) and all(torch.isfinite(v) for v in parts.values())
    missing = [n for n, p in model.named_parameters() if p.grad is None]
    assert not missing, missing[:5]
    g_native = {n: p.grad.clone() for n, p in model.named_parameters()}

    # the same model with torch modules in place of the kernel-backed ones (what the reference itself runs)
    import semstereo_b200.submodule_other as other
    ref = SemStereoTrainGlue(64).to(DEV).train()
    ref.load_state_dict(model.state_dict())       # NOTE: running stats were updated by the step above; reload the parameters only
    with torch.no_grad():
        for (n, p), (_, q) in zip(ref.named_parameters(), model.named_parameters()):
            p.copy_(q)

    def torchify(m):
        for name, child in list(m.named_children()):
            if type(child).__name__ == "_ConvBN3d":
                conv, bn = child[0], child[1]
                setattr(m, name, nn.Sequential(conv, bn))
            elif type(child).__name__ == "attention_block" and type(child).__module__.startswith("semstereo_b200"):
                from oracle import ops as oo

                class TorchAttn(nn.Module):
                    def __init__(self, src):
                        super().__init__()
                        self.qkv_3d, self.final1x1, self.block = src.qkv_3d, src.final1x1, src.block

                    def forward(self, x):
                        p = {"a.qkv_3d.weight": self.qkv_3d.weight, "a.qkv_3d.bias": self.qkv_3d.bias,
                             "a.final1x1.weight": self.final1x1.weight, "a.final1x1.bias": self.final1x1.bias}
                        return oo.window_attention3d(x, p, "a", 16, self.block)
                setattr(m, name, TorchAttn(child))
            else:
                torchify(child)
    torchify(ref)
    loss_r, _ = T.total_loss(ref(fl, fr), disp, disp4, label, 64)
    loss_r.backward()
    assert abs(float(loss) - float(loss_r)) <= 1e-4 * abs(float(loss_r))
    checked = 0
    for n, p in ref.named_parameters():
        if n.startswith(("hourglass_att", "classif_att_", "corr_feature_att_8", "patch", "feature_up.deconv8_4")) and p.grad is not None:
            a, b = g_native[n], p.grad
            if b.abs().max().item() > 1e-6:       # fp32 rounding differs between the two routes and the forward has discontinuities
                # (sort / top-k): a 1e-6 perturbation of the forward moves a few samples, measured 0.7 % - 3.5 % of max|grad| depending
                # on which samples a run lands on (two runs of the same code differ by as much, see below)
                assert relerr(a, b) <= 1e-1, (n, relerr(a, b))
                checked += 1
    assert checked >= 20

    # the tensor-core route (bf16x3 forward / dX) against the fp32 FFMA kernels on the same model and batch.  The glue's forward is
    # not bitwise reproducible between two runs of the SAME route (its torch / cuDNN parts pick different algorithms on a first call:
    # the loss toggles between 56.750996 and 56.750999), and one moved top-k sample shifts a weight gradient by up to 3.5 % of its
    # maximum -- so this comparison can only be as tight as the one above; when both runs land on the same samples the two routes
    # agree to 1e-5 (the tight, discontinuity-free gradient checks are test_conv3d_forward_and_gradients, per layer, <= 1e-4).
    def native_grads(tc_on):
        train_ops.set_tensor_core_route(tc_on)
        try:
            torch.manual_seed(3)
            mm = SemStereoTrainGlue(64).to(DEV).train()
            ls, _ = T.total_loss(mm(fl, fr), disp, disp4, label, 64)
            ls.backward()
            return {n: p.grad.clone() for n, p in mm.named_parameters() if p.grad is not None}
        finally:
            train_ops.set_tensor_core_route(True)
    g_tc, g_f32 = native_grads(True), native_grads(False)
    for n, b in g_f32.items():
        if n.startswith(("hourglass_att", "classif_att_", "patch")) and b.abs().max().item() > 1e-6:
            assert relerr(g_tc[n], b) <= 1e-1, (n, relerr(g_tc[n], b))


@pytest.mark.parametrize("Cin,Cout,k", [(1, 6, 3), (6, 6, 1), (6, 1, 1)])
def test_small_channel_conv2d_of_ssr_upsample(Cin, Cout, k):
    """The SSR_upsample convs on the small-channel kernels (forward, dX = the same kernel with the flipped / transposed weight, dW)."""
    g = torch.Generator().manual_seed(10 * Cin + Cout + k)
    x = torch.randn(2, Cin, 37, 52, generator=g).to(DEV).requires_grad_(True)
    w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).to(DEV).requires_grad_(True)
    b = torch.randn(Cout, generator=g).to(DEV).requires_grad_(True)
    y = train_ops.conv2d(x, w, b)
    ref = F.conv2d(x, w, b, 1, k // 2)
    assert relerr(y, ref) <= 1e-5
    gy = torch.randn(ref.shape, generator=g).to(DEV)
    gx, gw, gb = torch.autograd.grad(y, (x, w, b), gy)
    rx, rw, rb = torch.autograd.grad(ref, (x, w, b), gy)
    assert relerr(gx, rx) <= 1e-5 and relerr(gw, rw) <= 1e-4 and relerr(gb, rb) <= 1e-5, (relerr(gx, rx), relerr(gw, rw))
