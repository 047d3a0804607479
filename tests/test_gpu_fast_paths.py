"""The shape-specialised kernels behind sg_blur3 / sg_rgb_in / sg_rgb_out / sg_channel_sum (csrc/sg_blur_v2.cu,
csrc/sg_rgb_k3.cu) against the general kernels of the same entry points (switch: sg_fast_paths) and against plain torch."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _s():
    return torch.cuda.current_stream().cuda_stream


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


class general_kernels:
    def __enter__(self):
        from models._lib import lib
        self.old = lib.sg_fast_paths(0)

    def __exit__(self, *a):
        from models._lib import lib
        lib.sg_fast_paths(self.old)


def both(fn):
    """fn() -> tensor; evaluated with the fast paths on and off."""
    from models._lib import lib
    assert lib.sg_fast_paths(-1) == 1
    fast = fn()
    with general_kernels():
        slow = fn()
    torch.cuda.synchronize()
    return fast, slow


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("shape", [(2, 8, 4, 8), (1, 16, 16, 64), (3, 8, 12, 16), (2, 64, 64, 128), (1, 8, 8, 512),
                                   (2, 256, 256, 64)])
def test_blur_fast_equals_general_and_torch(dt, shape):
    from models._lib import lib, ptr, code
    torch.manual_seed(1)
    N, H, W, C = shape
    x = torch.randn(N, H, W, C, device=DEV).to(dt)
    bias = torch.randn(C, device=DEV)
    k = torch.tensor([1.0, 2.0, 1.0], device=DEV)
    k2 = (k[:, None] * k[None, :] / 16.0)[None, None].expand(C, 1, 3, 3)
    for b in (None, bias):
        def run():
            out = torch.empty_like(x)
            lib.sg_blur3(ptr(x), ptr(out), ptr(b), 0.5, code(dt), N, H, W, C, _s())
            return out
        fast, slow = both(run)
        tol = 1e-6 if dt == torch.float32 else 4e-3
        assert rel(fast.float(), slow.float()) < tol
        ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), k2, padding=1, groups=C).permute(0, 2, 3, 1)
        if b is not None:
            ref = ref + 0.5 * b
        assert rel(fast.float(), ref) < (1e-5 if dt == torch.float32 else 4e-3)


@pytest.mark.parametrize("act", [0, 1])
@pytest.mark.parametrize("N,hw,C", [(2, 8, 16), (3, 5, 64), (1, 64, 128), (4, 16, 512), (2, 256, 64)])
def test_rgb_in_fast_equals_general_and_torch(act, N, hw, C):
    from models._lib import lib, ptr
    torch.manual_seed(2)
    img = torch.randn(N, 3, hw, hw, device=DEV)
    W, b = torch.randn(C, 3, device=DEV), torch.randn(C, device=DEV)

    def run():
        out = torch.empty(N, hw, hw, C, device=DEV, dtype=torch.bfloat16)
        lib.sg_rgb_in(ptr(img), ptr(W), ptr(b), 0.3, 1.0, act, ptr(out), 1, N, hw * hw, 3, C, _s())
        return out
    fast, slow = both(run)
    assert rel(fast.float(), slow.float()) < 4e-3
    ref = torch.einsum("nkhw,ck->nhwc", img, W) * 0.3 + b
    if act:
        ref = torch.nn.functional.leaky_relu(ref, 0.2)
    assert rel(fast.float(), ref) < 4e-3


@pytest.mark.parametrize("up,beta", [(1, 0.0), (2, 0.0), (1, 0.7), (2, 0.4)])
@pytest.mark.parametrize("N,hw,C", [(2, 16, 64), (1, 32, 128), (3, 16, 16), (1, 64, 256), (2, 128, 64)])
def test_rgb_out_fast_equals_general_and_torch(up, beta, N, hw, C):
    from models._lib import lib, ptr
    torch.manual_seed(3)
    x = torch.randn(N, hw, hw, C, device=DEV).to(torch.bfloat16)
    W, b = torch.randn(3, C, device=DEV), torch.randn(3, device=DEV)
    init = torch.randn(N, 3, hw * up, hw * up, device=DEV)

    def run():
        out = init.clone()
        lib.sg_rgb_out(ptr(x), 1, ptr(W), ptr(b), 0.2, 1.0, 0.6, beta, ptr(out), N, hw, hw, C, 3, up, _s())
        return out
    fast, slow = both(run)
    assert rel(fast, slow) < 1e-5
    rgb = (torch.einsum("nhwc,kc->nkhw", x.float(), W) * 0.2 + b.view(1, 3, 1, 1)) * 0.6
    if up == 2:
        rgb = rgb.repeat_interleave(2, 2).repeat_interleave(2, 3)
    assert rel(fast, beta * init + rgb) < 1e-5


@pytest.mark.parametrize("N,hw,C", [(2, 8, 16), (3, 31, 64), (1, 64, 128), (2, 16, 512), (2, 256, 64)])
def test_channel_sum_fast_equals_general_and_torch(N, hw, C):
    from models._lib import lib, ptr
    torch.manual_seed(4)
    P = N * hw * hw
    x = torch.randn(N, hw, hw, C, device=DEV).to(torch.bfloat16)
    wt = torch.randn(N, 1, hw, hw, device=DEV)

    def run():
        out = torch.empty(2, C, device=DEV)
        ws = torch.empty(lib.sg_channel_sum_blocks(P) * 2 * C, device=DEV)
        lib.sg_channel_sum(ptr(x), 1, ptr(wt), 1, 1, P, hw * hw, C, ptr(ws), ptr(out), 0.5, 0, _s())
        return out
    fast, slow = both(run)
    assert rel(fast, slow) < 1e-4
    xf = x.float().reshape(P, C)
    ref = torch.stack([(wt.reshape(P, 1) * xf).sum(0), xf.sum(0)]) * 0.5
    assert rel(fast, ref) < 1e-4


@pytest.mark.parametrize("B,I,Oo", [(5, 512, 40), (32, 512, 1024), (40, 256, 130), (4, 128, 8), (32, 8192, 512),
                                     (7, 48, 40)])
@pytest.mark.parametrize("act", [False, True])
def test_linear_fast_equals_general_and_torch(B, I, Oo, act):
    """sg_linear_fwd / _dgrad / _wgrad: the weights-in-registers kernels (csrc/sg_linear_group.cu) for I in {128, 256,
    512}, the split-I path for wide inputs, the general kernels otherwise -- each against the other setting of the switch
    and against torch, forward and all gradients; strided input rows as in dlatents[:, l]."""
    import torch.nn.functional as F
    from models import _ops
    torch.manual_seed(6)
    big = torch.randn(B, 3, I, device=DEV)
    W = torch.nn.Parameter(torch.randn(Oo, I, device=DEV))
    b = torch.nn.Parameter(torch.randn(Oo, device=DEV))
    g = torch.randn(B, Oo, device=DEV)

    def run():
        x = big[:, 1].clone().requires_grad_(True) if I % 4 else big.clone().requires_grad_(True)
        xin = x if I % 4 else x[:, 1]
        y = _ops.Linear.apply(xin, W, b, 0.37, 0.5, act)
        gx, gW, gb = torch.autograd.grad(y, [x, W, b], g)
        return torch.cat([y.flatten(), gx.flatten(), gW.flatten(), gb.flatten()])
    fast, slow = both(run)
    assert rel(fast, slow) < 1e-5
    xr = big[:, 1].double().clone().requires_grad_(True)
    Wr, br = W.detach().double().requires_grad_(True), b.detach().double().requires_grad_(True)
    yr = xr @ (Wr * 0.37).t() + br * 0.5
    if act:
        yr = F.leaky_relu(yr, 0.2)
    grads = torch.autograd.grad(yr, [xr, Wr, br], g.double())
    n = B * Oo
    assert rel(fast[:n], yr.flatten()) < 1e-5
    assert rel(fast[-(Oo * I + Oo):-Oo], grads[1].flatten()) < 1e-5
    assert rel(fast[-Oo:], grads[2]) < 1e-5


# ---- weight preparation (csrc/sg_conv_generic.cu: wprep_fwd_kernel / wprep_bwd_kernel) ------------------------------
@pytest.mark.parametrize("co,ci,KK,npt", [(64, 64, 9, 9), (48, 80, 9, 16), (33, 7, 9, 9), (512, 512, 9, 9), (40, 24, 16, 16),
                                          (96, 32, 1, 1), (5, 130, 9, 16)])
@pytest.mark.parametrize("transpose", [0, 1])
def test_wprep_fwd_bwd_match_einsum(co, ci, KK, npt, transpose):
    """E[pt][a][b] = scale * sum_j coef[pt][j] W[co][ci][j] and its adjoint over pixel-split partials, for full tiles,
    ragged tiles (sizes that are not multiples of the 32 x 32 tile / 128-position block) and every tap count in use."""
    from models._lib import lib, ptr, code
    g = torch.Generator(device="cpu").manual_seed(co * 131 + ci * 7 + KK + npt + transpose)
    W = torch.randn(co, ci, KK, generator=g).to(DEV)
    coef = torch.randn(npt, KK, generator=g).to(DEV)
    scale = 0.37
    ref = torch.einsum("pj,oij->pio", coef.double(), W.double()) * scale          # [npt, ci, co]
    if transpose:
        ref = ref.transpose(1, 2).contiguous()
    for dt, tol in ((torch.float32, 1e-6), (torch.bfloat16, 4e-3)):
        E = torch.full(ref.shape, float("nan"), device=DEV, dtype=dt)
        lib.sg_wprep_fwd(ptr(W), ptr(E), code(dt), co, ci, KK, npt, ptr(coef), scale, transpose, _s())
        assert rel(E, ref) < tol
    for S in (1, 3, 6):
        part = torch.randn(S, *ref.shape, generator=g).to(DEV)                    # [S, npt, a, b]
        p = part.double().sum(0)
        if transpose:
            p = p.transpose(1, 2)                                                 # -> [npt, ci, co]
        want = torch.einsum("pj,pio->oij", coef.double(), p) * scale
        dW = torch.full((co, ci, KK), float("nan"), device=DEV)
        lib.sg_wprep_bwd(ptr(part), S, ptr(dW), co, ci, KK, npt, ptr(coef), scale, transpose, 0, _s())
        assert rel(dW, want) < 1e-5
        base = torch.randn(co, ci, KK, generator=g).to(DEV)
        acc = base.clone()
        lib.sg_wprep_bwd(ptr(part), S, ptr(acc), co, ci, KK, npt, ptr(coef), scale, transpose, 1, _s())
        assert rel(acc, want + base.double()) < 1e-5


@pytest.mark.parametrize("N,hw,C", [(2, 8, 16), (3, 31, 64), (1, 64, 128), (2, 16, 512), (2, 128, 64)])
@pytest.mark.parametrize("ones", [0, 1])
def test_channel_sum_three_weight_rows(N, hw, C, ones):
    """rgb weight gradients (J = 3 per-pixel weight rows, + the plain sum for from_rgb's bias): csrc/sg_rgb_k3.cu
    channel_sum_w3_kernel against the general kernel and torch."""
    from models._lib import lib, ptr
    torch.manual_seed(5)
    P = N * hw * hw
    rows = 3 + ones
    x = torch.randn(N, hw, hw, C, device=DEV).to(torch.bfloat16)
    wt = torch.randn(N, 3, hw, hw, device=DEV)

    def run():
        out = torch.full((rows, C), float("nan"), device=DEV)
        ws = torch.empty(lib.sg_channel_sum_blocks(P) * rows * C, device=DEV)
        lib.sg_channel_sum(ptr(x), 1, ptr(wt), 3, ones, P, hw * hw, C, ptr(ws), ptr(out), 0.5, 0, _s())
        return out
    fast, slow = both(run)
    assert rel(fast, slow) < 1e-4
    xf = x.float()
    ref = [torch.einsum("nhw,nhwc->c", wt[:, j], xf) for j in range(3)]
    if ones:
        ref.append(xf.sum((0, 1, 2)))
    assert rel(fast, torch.stack(ref) * 0.5) < 1e-4


@pytest.mark.parametrize("N,hw,C", [(2, 8, 16), (3, 31, 64), (1, 64, 128), (2, 16, 512), (2, 128, 64)])
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float32])
def test_lrelu_bwd_bias_equals_two_passes(N, hw, C, dt):
    """sg_lrelu_bwd_bias = sg_lrelu_bwd followed by the plain channel sum of its output (bitwise for gx; the sums are taken
    over the rounded gx in both), and differentiable through LreluBwdBias exactly like LreluBwd."""
    from models import _ops
    from models._lib import lib, ptr, code
    torch.manual_seed(7)
    g = torch.randn(N, hw, hw, C, device=DEV).to(dt)
    ref = torch.randn(N, hw, hw, C, device=DEV).to(dt)
    P = N * hw * hw
    assert lib.sg_lrelu_bwd_bias_supported(code(dt), P, C) == 1
    want = torch.empty_like(g)
    lib.sg_lrelu_bwd(ptr(g), ptr(ref), ptr(want), g.numel(), code(dt), _s())
    gx, gb = _ops.LreluBwdBias.apply(g, ref, 0.7)
    assert torch.equal(gx, want)
    assert rel(gb, want.float().sum((0, 1, 2)) * 0.7) < 1e-5
    gg = torch.randn_like(g).requires_grad_(True)
    out, _ = _ops.LreluBwdBias.apply(gg, ref, 1.0)
    (d,) = torch.autograd.grad(out, gg, g)
    want2 = torch.empty_like(g)
    lib.sg_lrelu_bwd(ptr(g), ptr(ref), ptr(want2), g.numel(), code(dt), _s())
    assert torch.equal(d, want2)


def test_batched_weight_refresh_equals_individual_builds():
    """refresh_derived re-issues every recorded bf16 sg_wprep_fwd of the updated parameters as ONE sg_wprep_fwd_multi launch
    (csrc/sg_conv_generic.cu): the refreshed buffers equal freshly built ones, bit for bit, for mixed shapes / tap counts /
    orientations, and the buffers are refreshed IN PLACE (a captured graph keeps reading the same addresses)."""
    from models import _ops, _tc
    from models._lib import launches
    _ops.clear_weight_cache()
    torch.manual_seed(11)
    shapes = [(64, 64), (128, 64), (48, 80), (512, 256)]
    Ws = [torch.nn.Parameter(torch.randn(co, ci, 3, 3, device=DEV)) for co, ci in shapes]
    tok = [0]
    for W in Ws:
        W._sg_token = tok
    def variants(W):
        out = [_tc.tc_weight(W, "fwd", 0.3), _tc.tc_weight(W, "dgrad", 0.3)]
        if W.shape[0] % 64 == 0 and W.shape[1] % 64 == 0:
            out += [_tc.ud_weight(W, "up", "fwd", 0.2), _tc.ud_weight(W, "down", "dgrad", 0.2)]
        return out
    before = [variants(W) for W in Ws]
    ptrs = [[e.data_ptr() for e in v] for v in before]
    with torch.no_grad():
        for W in Ws:
            W.data.add_(torch.randn_like(W))             # rewritten behind autograd's back (no _version bump), as the fused optimizer does
    tok[0] += 1
    l0 = launches()
    _ops.refresh_derived(Ws)
    assert launches() - l0 == 1                           # one launch for all 12 buffers
    after = [variants(W) for W in Ws]                     # cache hits: the refreshed buffers
    assert launches() - l0 == 1
    assert [[e.data_ptr() for e in v] for v in after] == ptrs
    refreshed = [[e.clone() for e in v] for v in after]
    _ops.clear_weight_cache()
    fresh = [variants(W) for W in Ws]
    for a, b in zip(refreshed, fresh):
        for x, y in zip(a, b):
            assert torch.equal(x, y)
