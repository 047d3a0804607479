"""CTA-pair (cta_group::2) tcgen05 convolution with the fused LayerEpilogue front half (csrc/sg_conv_pair_tc.cu):
kernel-level parity against the CPU oracle arithmetic (F.conv2d on bf16-rounded operands, fp32), the statistics against
sums over the kernel's own output, and the fused Function against the unfused composition."""
import math

import pytest
import torch
import torch.nn.functional as F

import stylegan_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


SHAPES = [(2, 16, 16, 64, 64), (2, 16, 24, 64, 128), (3, 32, 32, 128, 64), (2, 16, 16, 512, 512), (2, 64, 64, 64, 64),
          (4, 32, 16, 256, 256), (2, 16, 32, 128, 128), (6, 16, 8, 64, 64), (2, 32, 32, 256, 128), (1, 128, 128, 64, 64)]


@pytest.mark.parametrize("N,H,W,ci,co", SHAPES)
@pytest.mark.parametrize("fused", [False, True])
def test_pair_conv_matches_oracle(N, H, W, ci, co, fused):
    from models import _ops, _tc
    from models._lib import lib, ptr, stream
    _ops.clear_weight_cache()
    torch.manual_seed(0)
    assert lib.sg_conv3x3_pair_supported(N, H, W, ci, co) == 1
    x = torch.randn(N, ci, H, W).bfloat16().float()
    Wt = torch.randn(co, ci, 3, 3)
    scale = O.he_scale(ci * 9, math.sqrt(2))
    Wq = (Wt * scale).bfloat16().float()
    b = torch.randn(co)
    nz = torch.randn(N, 1, H, W)
    nw = torch.randn(co)
    ref = F.conv2d(x, Wq, padding=1) + 0.5 * b.view(1, -1, 1, 1)
    if fused:
        ref = ref + nw.view(1, -1, 1, 1) * nz
    ref = F.leaky_relu(ref, 0.2)
    xd = x.permute(0, 2, 3, 1).contiguous().to(DEV).bfloat16()
    Wd = torch.nn.Parameter(Wt.to(DEV))
    E = _tc.tc_weight(Wd, "fwd", scale, "same")
    y = torch.empty((N, H, W, co), dtype=torch.bfloat16, device=DEV)
    part = None
    if fused:
        slots = lib.sg_conv3x3_pair_slots(N, H, W, ci, co)
        assert slots >= 2
        part = torch.full((N, slots, 2, co), float("nan"), device=DEV)
    nzd, nwd = nz.to(DEV), nw.to(DEV)
    lib.sg_conv3x3_pair(ptr(xd), ptr(E), ptr(y), ptr(b.to(DEV)), 0.5, 1, ptr(nzd if fused else None),
                        ptr(nwd if fused else None), ptr(part), N, H, W, ci, co, stream())
    torch.cuda.synchronize()
    got = y.float().permute(0, 3, 1, 2)
    assert rel(got, ref) < 6e-3          # bf16 output rounding only (fp32 accumulation)
    if fused:
        style = torch.randn(N, 2 * co, device=DEV)
        mr = torch.empty(N, co, 2, device=DEV)
        ab = torch.empty(N, co, 3, device=DEV)
        lib.sg_epi_finalize_parts(ptr(part), ptr(style), ptr(mr), ptr(ab), N, H, W, ci, co, 1e-5, stream())
        torch.cuda.synchronize()
        z = y.double()                                         # statistics are those of the stored (rounded) tensor
        mean = z.mean(dim=(1, 2))
        var = z.var(dim=(1, 2), unbiased=False)
        rstd = 1.0 / torch.sqrt(var + 1e-5)
        assert rel(mr[..., 0], mean) < 1e-5
        assert rel(mr[..., 1], rstd) < 1e-4
        assert rel(ab[..., 1], rstd * (1.0 + style[:, :co].double())) < 1e-4
        assert torch.equal(ab[..., 2], style[:, co:])
        # deterministic: a second launch gives bit-identical statistics and output
        part2 = torch.full_like(part, float("nan"))
        y2 = torch.empty_like(y)
        lib.sg_conv3x3_pair(ptr(xd), ptr(E), ptr(y2), ptr(b.to(DEV)), 0.5, 1, ptr(nzd), ptr(nwd), ptr(part2), N, H, W,
                            ci, co, stream())
        mr2 = torch.empty_like(mr)
        lib.sg_epi_finalize_parts(ptr(part2), ptr(style), ptr(mr2), ptr(ab), N, H, W, ci, co, 1e-5, stream())
        torch.cuda.synchronize()
        assert torch.equal(y, y2) and torch.equal(mr, mr2)


@pytest.mark.parametrize("N,H,W,ci,co", [(2, 16, 16, 64, 64), (3, 32, 32, 128, 64), (2, 16, 16, 512, 512)])
def test_pair_dgrad_and_no_activation(N, H, W, ci, co):
    """the same kernel as the data gradient (transposed, 180-degree rotated weights), no bias, no activation"""
    from models import _ops, _tc
    from models._lib import lib
    _ops.clear_weight_cache()
    torch.manual_seed(1)
    assert lib.sg_conv3x3_pair_supported(N, H, W, co, ci) == 1
    g = torch.randn(N, co, H, W).bfloat16().float()
    Wt = torch.randn(co, ci, 3, 3)
    scale = 0.11
    ref = F.conv_transpose2d(g, (Wt * scale).bfloat16().float(), padding=1)
    gd = g.permute(0, 2, 3, 1).contiguous().to(DEV).bfloat16()
    y = torch.empty((N, H, W, ci), dtype=torch.bfloat16, device=DEV)
    assert _tc.try_conv(gd, torch.nn.Parameter(Wt.to(DEV)), "same", "dgrad", scale, None, 1.0, False, y)
    torch.cuda.synchronize()
    assert rel(y.float().permute(0, 3, 1, 2), ref) < 6e-3


def test_pair_equals_single_cta_kernel_at_the_north_star_shape():
    """64->64 @256x256 (batch 4): the pair kernel against the single-CTA tcgen05 kernel, element for element."""
    from models import _ops, _tc
    _ops.clear_weight_cache()
    torch.manual_seed(2)
    N, R, C = 4, 256, 64
    x = torch.randn(N, R, R, C, device=DEV).bfloat16()
    W = torch.nn.Parameter(torch.randn(C, C, 3, 3, device=DEV))
    b = torch.randn(C, device=DEV)
    outs = []
    for pair in (True, False):
        _ops._state["use_pair"] = pair
        try:
            y = torch.empty_like(x)
            assert _tc.try_conv(x, W, "same", "fwd", 0.05, b, 1.0, True, y)
            outs.append(y)
        finally:
            _ops._state["use_pair"] = True
    torch.cuda.synchronize()
    assert rel(outs[0].float(), outs[1].float()) < 1e-6      # same products, fp32 accumulation order may differ


@pytest.mark.parametrize("N,R,C", [(4, 16, 128), (2, 64, 64), (4, 32, 256)])
def test_fused_conv_epilogue_function_equals_unfused(N, R, C):
    """_ops.ConvEpilogue (conv + noise + lrelu + stats in one kernel) against conv -> _ops.Epilogue, forward and all
    gradients."""
    from models import _ops
    _ops.clear_weight_cache()
    torch.manual_seed(3)
    x0 = torch.randn(N, R, R, C, device=DEV).bfloat16()
    W = torch.nn.Parameter(torch.randn(C, C, 3, 3, device=DEV))
    bias = torch.nn.Parameter(0.1 * torch.randn(C, device=DEV))
    nw = torch.nn.Parameter(torch.randn(C, device=DEV))
    noise = torch.randn(N, 1, R, R, device=DEV)
    style0 = torch.randn(N, 2 * C, device=DEV)
    gout = torch.randn(N, R, R, C, device=DEV).bfloat16()
    scale = O.he_scale(C * 9, math.sqrt(2))
    res = []
    for fused in (True, False):
        x = x0.clone().requires_grad_(True)
        style = style0.clone().requires_grad_(True)
        if fused:
            out = _ops.ConvEpilogue.apply(x, W, bias, scale, 1.0, noise, nw, style, 1e-5)
        else:
            y = _ops.conv_layer(x, W, None, "same", scale, 1.0, False)
            out = _ops.Epilogue.apply(y, bias, 1.0, noise, nw, style, 1e-5, True)
        grads = torch.autograd.grad(out, [x, W, bias, nw, style], gout)
        res.append([out] + list(grads))
    torch.cuda.synchronize()
    names = ["out", "gx", "gW", "gbias", "gnoise_w", "gstyle"]
    # the unfused path rounds the convolution result to bf16 BEFORE bias/noise/activation, the fused one does not:
    # differences are one bf16 rounding of the pre-activation (and the LeakyReLU mask flips that follow from it).  The
    # bias / noise-weight gradients sit in front of InstanceNorm: they are small residuals of cancelling sums (exactly
    # zero for a linear activation), so the same absolute noise is a larger relative one (observed 3-4e-2).
    tol = dict(out=8e-3, gx=3e-2, gW=3e-2, gbias=1e-1, gnoise_w=1e-1, gstyle=3e-2)
    for n, a, b in zip(names, res[0], res[1]):
        assert rel(a.float(), b.float()) < tol[n], n


@pytest.mark.parametrize("N,R,C", [(4, 8, 512), (2, 64, 256), (3, 128, 128), (2, 256, 64), (5, 16, 512)])
def test_fused_blur_epilogue_function_equals_unfused(N, R, C):
    """_ops.BlurEpilogue (blur + bias + noise + lrelu + stats in one pass) against Blur -> Epilogue, forward and grads."""
    from models import _ops
    torch.manual_seed(4)
    y0 = torch.randn(N, R, R, C, device=DEV).bfloat16()
    assert _ops.BlurEpilogue.supported(y0)
    bias = torch.nn.Parameter(0.1 * torch.randn(C, device=DEV))
    nw = torch.nn.Parameter(torch.randn(C, device=DEV))
    noise = torch.randn(N, 1, R, R, device=DEV)
    style0 = torch.randn(N, 2 * C, device=DEV)
    gout = torch.randn(N, R, R, C, device=DEV).bfloat16()
    res = []
    for fused in (True, False):
        y = y0.clone().requires_grad_(True)
        style = style0.clone().requires_grad_(True)
        if fused:
            out = _ops.BlurEpilogue.apply(y, bias, 1.0, noise, nw, style, 1e-5)
        else:
            out = _ops.Epilogue.apply(_ops.Blur.apply(y, bias, 1.0), None, 1.0, noise, nw, style, 1e-5, True)
        grads = torch.autograd.grad(out, [y, bias, nw, style], gout)
        res.append([out] + list(grads))
    torch.cuda.synchronize()
    # the unfused path rounds the blurred tensor to bf16 before the noise/activation, the fused one does not
    tol = dict(out=8e-3, gy=3e-2, gbias=1e-1, gnoise_w=1e-1, gstyle=3e-2)      # see the conv test above
    for n, a, b in zip(["out", "gy", "gbias", "gnoise_w", "gstyle"], res[0], res[1]):
        assert rel(a.float(), b.float()) < tol[n], n
    # oracle arithmetic for the forward: blur of the bf16 input in fp32, then the epilogue
    x = y0.float().permute(0, 3, 1, 2).cpu()
    k = torch.tensor([1.0, 2.0, 1.0])
    k2 = (k[:, None] * k[None, :] / 16.0)[None, None].expand(C, 1, 3, 3)
    v = F.conv2d(x, k2, padding=1, groups=C) + bias.detach().cpu().view(1, -1, 1, 1) + \
        nw.detach().cpu().view(1, -1, 1, 1) * noise.cpu()
    zz = F.leaky_relu(v, 0.2)
    zn = O.instance_norm(zz)
    s = style0.cpu().view(N, 2, C, 1, 1)
    ref = zn * (s[:, 0] + 1.0) + s[:, 1]
    assert rel(res[0][0].float().permute(0, 3, 1, 2), ref) < 1.2e-2


@pytest.mark.parametrize("B,L,D", [(4, 6, 512), (32, 14, 512), (40, 4, 512), (8, 4, 128)])
def test_grouped_style_affines_equal_separate_linears(B, L, D):
    """_ops.StyleAffines (one launch for all StyleMod affines, one for their dgrad, one for their wgrad) against the
    per-layer _ops.Linear and against plain torch."""
    from models import _ops
    torch.manual_seed(5)
    chans = [512, 512, 256, 128, 64, 32, 16][: max(1, L // 2)]
    idx = tuple(range(0, 2 * len(chans)))[:L]
    Os = [2 * chans[i // 2] for i in idx]
    dl0 = torch.randn(B, L, D, device=DEV)
    Ws = [torch.nn.Parameter(torch.randn(O, D, device=DEV)) for O in Os]
    bs = [torch.nn.Parameter(torch.randn(O, device=DEV)) for O in Os]
    cfg = tuple((D ** -0.5, 1.0) for _ in Os)
    gouts = [torch.randn(B, O, device=DEV) for O in Os]
    params = [p for pair in zip(Ws, bs) for p in pair]
    res = []
    for grouped in (True, False):
        dl = dl0.clone().requires_grad_(True)
        if grouped:
            outs = _ops.StyleAffines.apply(dl, idx, cfg, *params)
        else:
            outs = [_ops.Linear.apply(dl[:, l].contiguous(), W, b, c[0], c[1], False) for l, W, b, c in zip(idx, Ws, bs, cfg)]
        if B > 32 and grouped:
            with _ops.skip_param_grads():
                grads = torch.autograd.grad(outs, [dl], gouts)
        else:
            grads = torch.autograd.grad(outs, [dl] + (params if B <= 32 else []), gouts)
        res.append((list(outs), list(grads)))
    torch.cuda.synchronize()
    for a, b in zip(res[0][0], res[1][0]):
        assert rel(a, b) < 1e-5
    for a, b in zip(res[0][1], res[1][1]):
        assert rel(a, b) < 1e-5
    for l, W, b, o in zip(idx, Ws, bs, res[0][0]):
        ref = dl0[:, l].double() @ (W.double() * D ** -0.5).t() + b.double()
        assert rel(o, ref) < 1e-5
