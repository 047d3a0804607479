"""Kernel-level parity (through the C-ABI, via the autograd Functions) against the CPU oracle's primitives.
fp32 mode: north-star tolerance 1e-3 relative per tensor (observed ~1e-6).  bf16 mode: 3e-2 (bf16 parity unpinned:
the reference has no reduced-precision path)."""
import math

import pytest
import torch
import torch.nn.functional as F

import stylegan_oracle as O

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 1e-3, torch.bfloat16: 3e-2}
DEV = "cuda"


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def nhwc(x, dt):
    return x.permute(0, 2, 3, 1).contiguous().to(DEV).to(dt)


def nchw(y):
    return y.detach().float().permute(0, 3, 1, 2).cpu()


@pytest.fixture(autouse=True)
def _fresh():
    from models import _ops
    _ops.clear_weight_cache()
    _ops.set_precision("fp32")
    yield
    _ops.set_precision("fp32")


def _ref_conv(kind, x, W, scale):
    ws = W * scale
    if kind == "same":
        return F.conv2d(x, ws, padding=W.shape[-1] // 2)
    if kind == "down":
        return F.avg_pool2d(F.conv2d(x, ws, padding=1), 2)
    if kind == "up":
        return F.conv2d(O.nearest_up2(x), ws, padding=1)
    wt = F.pad(ws.permute(1, 0, 2, 3), [1, 1, 1, 1])
    wt = wt[:, :, 1:, 1:] + wt[:, :, :-1, 1:] + wt[:, :, 1:, :-1] + wt[:, :, :-1, :-1]
    return F.conv_transpose2d(x, wt, stride=2, padding=1)


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("kind,k,ci,co,hw", [("same", 3, 16, 24, 12), ("same", 1, 8, 12, 8), ("down", 3, 16, 32, 16),
                                              ("up", 3, 32, 16, 8), ("up_fused", 3, 8, 4, 64), ("same", 3, 9, 8, 4),
                                              ("same", 3, 64, 64, 32), ("down", 3, 4, 8, 128),
                                              ("same", 3, 40, 1, 8), ("same", 3, 1, 40, 8)])
def test_conv_forward_backward(kind, k, ci, co, hw, dt):
    from models import _ops
    torch.manual_seed(0)
    N = 3
    x = torch.randn(N, ci, hw, hw)
    W = torch.randn(co, ci, k, k)
    b = torch.randn(co)
    scale = O.he_scale(ci * k * k, math.sqrt(2))
    if dt == torch.bfloat16:                 # compare like with like: the oracle sees bf16-rounded operands
        x = x.bfloat16().float()
        W = (W * scale).bfloat16().float() / scale
    xr = x.clone().requires_grad_(True)
    Wr = W.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    y_ref = F.leaky_relu(_ref_conv(kind, xr, Wr, scale) + br.view(1, -1, 1, 1), 0.2)
    gy = torch.randn_like(y_ref)
    gx_ref, gW_ref, gb_ref = torch.autograd.grad(y_ref, [xr, Wr, br], gy)

    xd = nhwc(x, dt).requires_grad_(True)
    Wd = torch.nn.Parameter(W.to(DEV))
    bd = torch.nn.Parameter(b.to(DEV))
    y = _ops.ConvBiasAct.apply(xd, Wd, bd, kind, scale, 1.0, True)
    assert rel(nchw(y), y_ref) < TOL[dt]
    gx, gW, gb = torch.autograd.grad(y, [xd, Wd, bd], nhwc(gy, dt))
    # bf16: LeakyReLU masks of near-zero pre-activations flip under rounding -> sqrt(eps)-sized gradient noise
    gt = TOL[dt] if dt == torch.float32 else 6e-2
    assert rel(nchw(gx), gx_ref) < gt
    assert rel(gW, gW_ref) < gt
    assert rel(gb, gb_ref) < gt


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_conv_double_backward_r1_like(dt):
    """Two conv+lrelu layers, a down conv and a blur; penalty = |d out / d x|^2 ; gradients wrt all weights."""
    from models import _ops
    torch.manual_seed(1)
    N, c0, c1, c2, hw = 2, 8, 16, 8, 16
    x = torch.randn(N, c0, hw, hw)
    W1, W2 = torch.randn(c1, c0, 3, 3), torch.randn(c2, c1, 3, 3)
    b1, b2 = 0.1 * torch.randn(c1), 0.1 * torch.randn(c2)
    s1, s2 = O.he_scale(c0 * 9, math.sqrt(2)), O.he_scale(c1 * 9, math.sqrt(2))

    def ref(x, W1, b1, W2, b2):
        h = F.leaky_relu(_ref_conv("same", x, W1, s1) + b1.view(1, -1, 1, 1), 0.2)
        h = O.blur3(h)
        return F.leaky_relu(_ref_conv("down", h, W2, s2) + b2.view(1, -1, 1, 1), 0.2)

    ps = [t.clone().requires_grad_(True) for t in (W1, b1, W2, b2)]
    xr = x.clone().requires_grad_(True)
    out = ref(xr, *ps)
    (g,) = torch.autograd.grad(out.sum(), xr, create_graph=True)
    pen = (g * g).sum() + out.mean()
    ref_grads = torch.autograd.grad(pen, ps)

    xd = nhwc(x, dt).requires_grad_(True)
    pd = [torch.nn.Parameter(t.to(DEV)) for t in (W1, b1, W2, b2)]
    h = _ops.ConvBiasAct.apply(xd, pd[0], pd[1], "same", s1, 1.0, True)
    h = _ops.Blur.apply(h, None, 1.0)
    o = _ops.ConvBiasAct.apply(h, pd[2], pd[3], "down", s2, 1.0, True)
    (gd,) = torch.autograd.grad(o.float().sum(), xd, create_graph=True)
    pen_d = _ops.SqSum.apply(_ops.Cast.apply(gd, torch.float32), 1.0).reshape(()) + o.float().mean()
    assert abs(float(pen_d) - float(pen)) < TOL[dt] * abs(float(pen)) * 3
    grads = torch.autograd.grad(pen_d, pd)
    for a, b in zip(grads, ref_grads):
        assert rel(a, b) < (3e-3 if dt == torch.float32 else 0.15)


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_blur_and_axpby(dt):
    from models import _ops
    torch.manual_seed(2)
    x = torch.randn(2, 12, 9, 7)
    b = torch.randn(12)
    xd = nhwc(x, dt).requires_grad_(True)
    bd = b.to(DEV).requires_grad_(True)
    y = _ops.Blur.apply(xd, bd, 1.0)
    xr = (x.bfloat16().float() if dt == torch.bfloat16 else x).requires_grad_(True)
    br = b.clone().requires_grad_(True)
    y_ref = O.blur3(xr) + br.view(1, -1, 1, 1)
    assert rel(nchw(y), y_ref) < TOL[dt]
    gy = torch.randn_like(y_ref)
    gx, gb = torch.autograd.grad(y, [xd, bd], nhwc(gy, dt))
    gx_ref, gb_ref = torch.autograd.grad(y_ref, [xr, br], gy)
    assert rel(nchw(gx), gx_ref) < TOL[dt] and rel(gb, gb_ref) < TOL[dt]
    z = _ops.Axpby.apply(xd, y, 0.3, 0.7)
    assert rel(nchw(z), 0.3 * xr + 0.7 * y_ref) < TOL[dt]


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(2, 16, 9, 7), (1, 64, 20, 5), (3, 8, 4, 4), (1, 256, 17, 3), (2, 64, 25, 20)])
def test_blur_vector_path_odd_sizes(dt, shape):
    """Strip/batch logic of blur3_kernel: heights that are not a multiple of the 8-row strip, 1-3 pixel wide borders,
    the 16-byte vector path in both dtypes; and the plain per-channel sum (bias gradient) on the same odd sizes."""
    from models import _ops
    torch.manual_seed(11)
    x = torch.randn(*shape)
    if dt == torch.bfloat16:
        x = x.bfloat16().float()
    xd = nhwc(x, dt)
    y = _ops.Blur.apply(xd, None, 1.0)
    assert rel(nchw(y), O.blur3(x)) < TOL[dt]
    s = _ops.ChannelSum.apply(xd, None, True, 0.5)
    assert rel(s.reshape(-1), 0.5 * x.sum(dim=(0, 2, 3))) < 1e-4


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("N,C,hw", [(1, 24, 5), (3, 8, 7), (2, 64, 6), (1, 512, 4)])
def test_rgb_kernels_ragged_sizes(dt, N, C, hw):
    """to_rgb / from_rgb adjoints with pixel counts that do not fill the last lane group, channel-vector counts that
    are not a power of two (idle lanes), and C wider than one warp of vectors (second-pass loop)."""
    from models import _ops
    torch.manual_seed(12)
    x_hi, x_lo = torch.randn(N, C, hw * 2, hw * 2), torch.randn(N, C, hw, hw)
    if dt == torch.bfloat16:
        x_hi, x_lo = x_hi.bfloat16().float(), x_lo.bfloat16().float()
    Wh, bh, Wl, bl = torch.randn(3, C, 1, 1), torch.randn(3), torch.randn(3, C, 1, 1), torch.randn(3)
    leaves = [t.clone().requires_grad_(True) for t in (x_hi, Wh, bh, x_lo, Wl, bl)]
    a = 0.6
    out_ref = a * O.eq_conv(leaves[0], leaves[1], leaves[2], 1.0) + \
        (1 - a) * O.eq_conv(O.nearest_up2(leaves[3]), leaves[4], leaves[5], 1.0)
    g = torch.randn_like(out_ref)
    ref_grads = torch.autograd.grad(out_ref, leaves, g)
    dl = [nhwc(x_hi, dt).requires_grad_(True), torch.nn.Parameter(Wh.to(DEV)), torch.nn.Parameter(bh.to(DEV)),
          nhwc(x_lo, dt).requires_grad_(True), torch.nn.Parameter(Wl.to(DEV)), torch.nn.Parameter(bl.to(DEV))]
    out = _ops.ToRgb.apply(dl[0], dl[1], dl[2], O.he_scale(C, 1.0), a, dl[3], dl[4], dl[5], O.he_scale(C, 1.0), 1 - a)
    assert rel(out, out_ref) < TOL[dt]
    grads = torch.autograd.grad(out, dl, g.to(DEV))
    assert rel(nchw(grads[0]), ref_grads[0]) < TOL[dt] and rel(nchw(grads[3]), ref_grads[3]) < TOL[dt]
    for i in (1, 2, 4, 5):
        assert rel(grads[i], ref_grads[i]) < TOL[dt]
    # from_rgb and its image gradient
    img = torch.randn(N, 3, hw, hw)
    W, b = torch.randn(C, 3, 1, 1), torch.randn(C)
    lv = [t.clone().requires_grad_(True) for t in (img, W, b)]
    y_ref = O.eq_conv(lv[0], lv[1], lv[2], math.sqrt(2))
    gy = torch.randn_like(y_ref)
    rg = torch.autograd.grad(y_ref, lv, gy)
    imgd = img.to(DEV).requires_grad_(True)
    Wd, bd = torch.nn.Parameter(W.to(DEV)), torch.nn.Parameter(b.to(DEV))
    y = _ops.RgbIn.apply(imgd, Wd, bd, O.he_scale(3, math.sqrt(2)), 1.0, False, dt)
    assert rel(nchw(y), y_ref) < TOL[dt]
    gr = torch.autograd.grad(y, [imgd, Wd, bd], nhwc(gy, dt))
    for u, v in zip(gr, rg):
        assert rel(u, v) < TOL[dt]


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("c,hw", [(16, 8), (64, 32), (4, 4)])
def test_layer_epilogue(dt, c, hw):
    from models import _ops
    torch.manual_seed(3)
    N = 3
    y = torch.randn(N, c, hw, hw) * 2 + 0.5
    if dt == torch.bfloat16:
        y = y.bfloat16().float()
    bias, nw = torch.randn(c), torch.randn(c)
    noise = torch.randn(N, 1, hw, hw)
    sw, sb = torch.randn(2 * c, 32), 0.1 * torch.randn(2 * c)
    dl = torch.randn(N, 32)
    leaves = [t.clone().requires_grad_(True) for t in (y, bias, nw, sw, sb, dl)]
    out_ref = O.layer_epilogue(leaves[0] + leaves[1].view(1, -1, 1, 1), leaves[2], noise, leaves[3], leaves[4],
                               leaves[5])
    g = torch.randn_like(out_ref)
    ref_grads = torch.autograd.grad(out_ref, leaves, g)

    yd = nhwc(y, dt).requires_grad_(True)
    bd, nwd, swd, sbd, dld = [t.to(DEV).requires_grad_(True) for t in (bias, nw, sw, sb, dl)]
    style = _ops.Linear.apply(dld, swd, sbd, O.he_scale(32, 1.0), 1.0, False)
    out = _ops.Epilogue.apply(yd, bd, 1.0, noise.to(DEV), nwd, style, 1e-5, True)
    assert rel(nchw(out), out_ref) < TOL[dt]
    grads = torch.autograd.grad(out, [yd, bd, nwd, swd, sbd, dld], nhwc(g, dt))
    assert rel(nchw(grads[0]), ref_grads[0]) < TOL[dt] * 2
    for a, b in zip(grads[1:], ref_grads[1:]):
        assert rel(a, b) < TOL[dt] * 2


def test_linear_pixelnorm_mapping():
    from models import _ops
    torch.manual_seed(4)
    B, I, Oo = 5, 48, 40
    x, W, b = torch.randn(B, I), torch.randn(Oo, I), torch.randn(Oo)
    leaves = [t.clone().requires_grad_(True) for t in (x, W, b)]
    y_ref = F.leaky_relu(O.eq_linear(O.pixel_norm(leaves[0]), leaves[1], leaves[2], math.sqrt(2), 0.01), 0.2)
    g = torch.randn_like(y_ref)
    ref_grads = torch.autograd.grad(y_ref, leaves, g)
    xd, Wd, bd = [t.to(DEV).requires_grad_(True) for t in (x, W, b)]
    y = _ops.Linear.apply(_ops.PixelNorm.apply(xd, 1e-8), Wd, bd, O.he_scale(I, math.sqrt(2), 0.01), 0.01, True)
    assert rel(y, y_ref) < 1e-4
    for a, r in zip(torch.autograd.grad(y, [xd, Wd, bd], g.to(DEV)), ref_grads):
        assert rel(a, r) < 1e-4
    # strided rows (dlatent slices) and double backward of the dgrad
    big = torch.randn(B, 3, I, device=DEV)
    y2 = _ops.Linear.apply(big[:, 1], Wd, None, 0.5, 1.0, False)
    assert rel(y2, (big[:, 1].cpu() @ W.t()) * 0.5) < 1e-5
    xr = x.clone().requires_grad_(True)
    Wr = W.clone().requires_grad_(True)
    o = F.leaky_relu(xr @ (Wr * 0.3).t(), 0.2)
    (gx,) = torch.autograd.grad(o.sum(), xr, create_graph=True)
    (gw_ref,) = torch.autograd.grad((gx * gx).sum(), Wr)
    xd2 = x.to(DEV).requires_grad_(True)
    od = _ops.Linear.apply(xd2, Wd, None, 0.3, 1.0, True)
    (gxd,) = torch.autograd.grad(od.sum(), xd2, create_graph=True)
    (gw,) = torch.autograd.grad(_ops.SqSum.apply(gxd, 1.0), Wd)
    assert rel(gw, gw_ref) < 1e-4


@pytest.mark.parametrize("B,G", [(8, 4), (4, 4), (2, 4), (12, 4)])
def test_minibatch_stddev_all_orders(B, G):
    from models import _ops
    torch.manual_seed(5)
    C, hw = 6, 4
    x = torch.randn(B, C, hw, hw)
    xr = x.clone().requires_grad_(True)
    y_ref = O.minibatch_stddev(xr, G)
    w = torch.randn_like(y_ref)
    (gx_ref,) = torch.autograd.grad((y_ref * w).sum(), xr, create_graph=True)
    (ggx_ref,) = torch.autograd.grad((gx_ref * gx_ref).sum(), xr)
    xd = x.permute(0, 2, 3, 1).contiguous().to(DEV).requires_grad_(True)
    y = _ops.MbStd.apply(xd, G)
    assert rel(y.permute(0, 3, 1, 2), y_ref) < 1e-5
    wd = w.permute(0, 2, 3, 1).contiguous().to(DEV)
    (gx,) = torch.autograd.grad((y * wd).sum(), xd, create_graph=True)
    assert rel(gx.permute(0, 3, 1, 2), gx_ref) < 1e-4
    if min(G, B) > 2:      # with groups of 2 the stddev is |x1-x2|/2: its second derivative is pure round-off
        (ggx,) = torch.autograd.grad(_ops.SqSum.apply(gx, 1.0), xd)
        assert rel(ggx.permute(0, 3, 1, 2), ggx_ref) < 1e-3


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_rgb_in_out_and_pooling(dt):
    from models import _ops
    torch.manual_seed(6)
    N, C, hw = 2, 16, 8
    img = torch.randn(N, 3, hw, hw)
    W, b = torch.randn(C, 3, 1, 1), torch.randn(C)
    s = O.he_scale(3, math.sqrt(2))
    leaves = [t.clone().requires_grad_(True) for t in (img, W, b)]
    y_ref = O.eq_conv(O.avg_pool2(leaves[0]), leaves[1], leaves[2], math.sqrt(2))
    g = torch.randn_like(y_ref)
    ref_grads = torch.autograd.grad(y_ref, leaves, g)
    imgd = img.to(DEV).requires_grad_(True)
    Wd, bd = torch.nn.Parameter(W.to(DEV)), torch.nn.Parameter(b.to(DEV))
    y = _ops.RgbIn.apply(_ops.PoolNCHW.apply(imgd, 2, 0.25), Wd, bd, s, 1.0, False, dt)
    assert rel(nchw(y), y_ref) < TOL[dt]
    grads = torch.autograd.grad(y, [imgd, Wd, bd], nhwc(g, dt))
    for a, r in zip(grads, ref_grads):
        assert rel(a, r) < TOL[dt]
    # to_rgb with fade-in
    x_hi, x_lo = torch.randn(N, C, hw, hw), torch.randn(N, 2 * C, hw // 2, hw // 2)
    if dt == torch.bfloat16:
        x_hi, x_lo = x_hi.bfloat16().float(), x_lo.bfloat16().float()
    Wh, bh, Wl, bl = torch.randn(3, C, 1, 1), torch.randn(3), torch.randn(3, 2 * C, 1, 1), torch.randn(3)
    leaves = [t.clone().requires_grad_(True) for t in (x_hi, Wh, bh, x_lo, Wl, bl)]
    a = 0.3
    out_ref = a * O.eq_conv(leaves[0], leaves[1], leaves[2], 1.0) + \
        (1 - a) * O.eq_conv(O.nearest_up2(leaves[3]), leaves[4], leaves[5], 1.0)
    g = torch.randn_like(out_ref)
    ref_grads = torch.autograd.grad(out_ref, leaves, g)
    dl = [nhwc(x_hi, dt).requires_grad_(True), torch.nn.Parameter(Wh.to(DEV)), torch.nn.Parameter(bh.to(DEV)),
          nhwc(x_lo, dt).requires_grad_(True), torch.nn.Parameter(Wl.to(DEV)), torch.nn.Parameter(bl.to(DEV))]
    out = _ops.ToRgb.apply(dl[0], dl[1], dl[2], O.he_scale(C, 1.0), a, dl[3], dl[4], dl[5], O.he_scale(2 * C, 1.0),
                           1 - a)
    assert rel(out, out_ref) < TOL[dt]
    grads = torch.autograd.grad(out, dl, g.to(DEV))
    assert rel(nchw(grads[0]), ref_grads[0]) < TOL[dt] and rel(nchw(grads[3]), ref_grads[3]) < TOL[dt]
    for i in (1, 2, 4, 5):
        assert rel(grads[i], ref_grads[i]) < TOL[dt]


def test_dlatents_losses_adam_randn():
    from models import _ops
    from models._optim import FusedAdam
    torch.manual_seed(7)
    B, L, D = 4, 6, 32
    w1, w2, avg = torch.randn(B, D), torch.randn(B, D), torch.randn(D)
    l1, l2 = w1.clone().requires_grad_(True), w2.clone().requires_grad_(True)
    dl = torch.where(torch.arange(L).view(1, -1, 1) < 3, l1.unsqueeze(1).expand(-1, L, -1),
                     l2.unsqueeze(1).expand(-1, L, -1))
    ref = O.truncate(dl, avg, 0.7, 4)
    g = torch.randn_like(ref)
    r1, r2 = torch.autograd.grad(ref, [l1, l2], g)
    d1, d2 = w1.to(DEV).requires_grad_(True), w2.to(DEV).requires_grad_(True)
    out = _ops.Dlatents.apply(d1, d2, avg.to(DEV), L, 3, 0.7, 4)
    assert rel(out, ref) < 1e-6
    g1, g2 = torch.autograd.grad(out, [d1, d2], g.to(DEV))
    assert rel(g1, r1) < 1e-5 and rel(g2, r2) < 1e-5
    # losses
    x = torch.randn(9, 1) * 3
    xr = x.clone().requires_grad_(True)
    ref = F.softplus(-xr).mean()
    (gr,) = torch.autograd.grad(ref, xr)
    xd = x.to(DEV).requires_grad_(True)
    o = _ops.SoftplusMean.apply(xd, -1.0)
    (gd,) = torch.autograd.grad(o, xd)
    assert rel(o, ref) < 1e-5 and rel(gd, gr) < 1e-5
    # Adam + clip + EMA against the oracle's restatement
    ps = {"a": torch.randn(1000), "b": torch.randn(17, 5), "c": torch.randn(3)}
    opt_ref = O.AdamState(list(ps), ps)
    pr = {k: v.clone() for k, v in ps.items()}
    pd = [torch.nn.Parameter(v.to(DEV)) for v in ps.values()]
    shadow = [torch.nn.Parameter(v.to(DEV).clone()) for v in ps.values()]
    sh_ref = {k: v.clone() for k, v in ps.items()}
    opt = FusedAdam(pd, lr=0.003, betas=(0.0, 0.99), eps=1e-8)
    for it in range(3):
        gs = {k: torch.randn_like(v) * (5.0 if it == 1 else 0.1) for k, v in ps.items()}
        if it == 2:
            gs["c"] = None
        used = [k for k in gs if gs[k] is not None]
        gc = {k: gs[k].clone() for k in used}
        O.clip_grad_norm(gc, used, 10.0)
        opt_ref.update(pr, gc)
        for k in pr:
            sh_ref[k] = 0.999 * sh_ref[k] + 0.001 * pr[k]
        for p, k in zip(pd, ps):
            p.grad = None if gs[k] is None else gs[k].to(DEV)
        opt.step(max_norm=10.0, ema_map=dict(zip(pd, shadow)), ema_decay=0.999)
    for p, s, k in zip(pd, shadow, ps):
        assert rel(p, pr[k]) < 1e-5, k
        assert rel(s, sh_ref[k]) < 1e-5, k
        assert rel(opt.state[p]["exp_avg_sq"], opt_ref.v[k]) < 1e-5
    # Philox normal generator: moments
    z = _ops.randn_device((1 << 20,), DEV, 1234, 0)
    assert abs(float(z.mean())) < 5e-3 and abs(float(z.std()) - 1) < 5e-3
    z2 = _ops.randn_device((1 << 20,), DEV, 1234, 0)
    assert torch.equal(z, z2)
