"""Backward (VJP) kernels of the volume / regression operators vs torch autograd through the CPU oracle (oracle/ops.py is plain
differentiable torch): the first step of BASELINE config #5.  fp32, tolerance relative to the gradient's scale."""
import pytest
import torch

from oracle import ops as oo

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

if torch.cuda.is_available():
    import semstereo_b200.torch_ops  # noqa: F401  (registers torch.ops.semstereo_b200.* with autograd)
    T = torch.ops.semstereo_b200


def rnd(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def close(got, ref, tol=2e-5):
    assert got.shape == ref.shape
    err = (got.cpu() - ref).abs().max().item()
    assert err <= tol * max(1.0, ref.abs().max().item()), (err, ref.abs().max().item())


@pytest.mark.parametrize("B,C,H,W,M,G", [(1, 16, 3, 24, 4, 4), (2, 32, 2, 37, 6, 8), (1, 256, 2, 128, 8, 32), (1, 24, 3, 9, 12, 3)])
@pytest.mark.parametrize("signed", [True, False])
@pytest.mark.parametrize("norm", [False, True])
def test_gwc_volume_backward(B, C, H, W, M, G, signed, norm):
    l, r = rnd(B, C, H, W, seed=1).requires_grad_(), rnd(B, C, H, W, seed=2).requires_grad_()
    gv = rnd(B, G, 2 * M if signed else M, H, W, seed=3)
    oo.gwc_volume(l, r, M, G, signed, norm).backward(gv)
    lc, rc = l.detach().to(DEV).requires_grad_(), r.detach().to(DEV).requires_grad_()
    T.gwc_volume(lc, rc, M, G, signed, norm).backward(gv.to(DEV))
    close(lc.grad, l.grad)
    close(rc.grad, r.grad)


@pytest.mark.parametrize("signed", [True, False])
def test_concat_volume_backward(signed):
    B, C, H, W, M = 2, 8, 3, 21, 5
    l, r = rnd(B, C, H, W, seed=1).requires_grad_(), rnd(B, C, H, W, seed=2).requires_grad_()
    gv = rnd(B, 2 * C, 2 * M if signed else M, H, W, seed=3)
    oo.concat_volume(l, r, M, signed).backward(gv)
    lc, rc = l.detach().to(DEV).requires_grad_(), r.detach().to(DEV).requires_grad_()
    T.concat_volume(lc, rc, M, signed).backward(gv.to(DEV))
    close(lc.grad, l.grad)
    close(rc.grad, r.grad)


@pytest.mark.parametrize("D,k", [(24, 2), (24, 5), (48, 3)])
def test_regression_topk_backward(D, k):
    c, s = rnd(2, D, 5, 13, seed=4).requires_grad_(), (4 * rnd(2, D, 5, 13, seed=5)).requires_grad_()
    g = rnd(2, 1, 5, 13, seed=6)
    oo.regression_topk(c, s, k).backward(g)
    cc, sc = c.detach().to(DEV).requires_grad_(), s.detach().to(DEV).requires_grad_()
    T.regression_topk(cc, sc, k).backward(g.to(DEV))
    close(cc.grad, c.grad)
    close(sc.grad, s.grad)


def test_context_upsample_and_disparity_regression_backward():
    d, w = rnd(2, 1, 6, 9, seed=7).requires_grad_(), rnd(2, 9, 24, 36, seed=8).requires_grad_()
    g = rnd(2, 24, 36, seed=9)
    oo.context_upsample(d, w).backward(g)
    dc, wc = d.detach().to(DEV).requires_grad_(), w.detach().to(DEV).requires_grad_()
    T.context_upsample(dc, wc).backward(g.to(DEV))
    close(dc.grad, d.grad)
    close(wc.grad, w.grad)
    for signed in (True, False):
        p = torch.softmax(rnd(2, 16 if signed else 8, 4, 7, seed=10), 1).requires_grad_()
        go = rnd(2, 4, 7, seed=11)
        oo.disparity_regression(p, 8, signed).backward(go)
        pc = p.detach().to(DEV).requires_grad_()
        T.disparity_regression(pc, 8, signed).backward(go.to(DEV))
        close(pc.grad, p.grad)


def test_propagation_and_variance_backward():
    for shape in ((2, 1, 6, 9), (1, 1, 4, 5, 7)):
        x = rnd(*shape, seed=12).requires_grad_()
        g = rnd(shape[0], 5, *shape[2:], seed=13)
        (oo.propagation(x) if len(shape) == 4 else oo.propagation_prob(x)).backward(g)
        xc = x.detach().to(DEV).requires_grad_()
        T.propagation(xc).backward(g.to(DEV))
        close(xc.grad, x.grad)
    for signed in (True, False):
        D = 16 if signed else 8
        p = torch.softmax(rnd(2, D, 4, 7, seed=14), 1).requires_grad_()
        mu = (3 * rnd(2, 1, 4, 7, seed=15)).requires_grad_()
        go = rnd(2, 1, 4, 7, seed=16)
        oo.disparity_variance(p, 8, mu, signed).backward(go)
        pc, mc = p.detach().to(DEV).requires_grad_(), mu.detach().to(DEV).requires_grad_()
        T.disparity_variance(pc, 8, mc, signed).backward(go.to(DEV))
        close(pc.grad, p.grad)
        close(mc.grad, mu.grad)


@pytest.mark.parametrize("integer_disp", [False, True])
def test_spatial_transformer_grid_backward(integer_disp):
    B, C, K, H, W = 2, 6, 5, 7, 19
    x, y = rnd(B, C, H, W, seed=17).requires_grad_(), rnd(B, C, H, W, seed=18).requires_grad_()
    d = 4 * rnd(B, K, H, W, seed=19)
    if integer_disp:
        d = d.round()
    d = (d + 0.25 * (not integer_disp)).requires_grad_()      # keep real-valued samples away from integers (kink of the bilinear)
    g1, g2 = rnd(B, C, K, H, W, seed=20), rnd(B, C, K, H, W, seed=21)
    yw, xr = oo.spatial_transformer_grid(x, y, d)
    (yw * g1 + xr * g2).sum().backward()
    xc, yc, dc = (t.detach().to(DEV).requires_grad_() for t in (x, y, d))
    yw2, xr2 = T.spatial_transformer_grid(xc, yc, dc)
    (yw2 * g1.to(DEV) + xr2 * g2.to(DEV)).sum().backward()
    close(xc.grad, x.grad)
    close(yc.grad, y.grad, 5e-5)                              # atomics: summation order differs
    if not integer_disp:                                      # at integer positions the derivative w.r.t. the sample is one-sided
        close(dc.grad, d.grad, 5e-5)


def test_surface_functions_are_differentiable():
    """The reference-named surface (semstereo_b200.submodule) records autograd when an input requires grad: a toy loss through
    build_gwc_volume_norm -> softmax -> disparity_regression + SpatialTransformer_grid back-propagates to the features exactly
    like the oracle composition does."""
    import semstereo_b200.submodule as sm
    l, r = rnd(1, 32, 4, 24, seed=30), rnd(1, 32, 4, 24, seed=31)

    def loss(gwc, reg, stn, a, b):
        vol = gwc(a, b, 4, 8)                                  # (1,8,8,4,24)
        p = torch.softmax(vol.mean(1), 1)                      # (1,8,4,24)
        mu = reg(p, 4)                                         # (1,4,24)
        yw, xr = stn(a, b, torch.stack((mu, mu + 1.5), 1))     # (1,32,2,4,24)
        return (yw * xr).mean() + mu.square().mean()

    a, b = l.clone().requires_grad_(), r.clone().requires_grad_()
    loss(lambda x, y, m, g: oo.gwc_volume(x, y, m, g, True, True), lambda p, m: oo.disparity_regression(p, m, True),
         oo.spatial_transformer_grid, a, b).backward()
    ac, bc = l.to(DEV).requires_grad_(), r.to(DEV).requires_grad_()
    out = loss(sm.build_gwc_volume_norm, sm.disparity_regression, sm.SpatialTransformer_grid, ac, bc)
    out.backward()
    close(ac.grad, a.grad, 1e-4)
    close(bc.grad, b.grad, 1e-4)
    with torch.no_grad():                                      # inference calls bypass the dispatcher and still agree
        assert torch.equal(sm.build_gwc_volume_norm(ac, bc, 4, 8), sm.build_gwc_volume_norm(ac.detach(), bc.detach(), 4, 8))
