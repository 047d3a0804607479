"""tcgen05 / TMA 3x3 convolution (csrc/sg_conv_tc.cu) against the generic CUDA-core kernel and the CPU oracle."""
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


@pytest.mark.parametrize("N,H,W,ci,co", [(2, 16, 16, 64, 64), (1, 8, 32, 64, 128), (3, 32, 32, 128, 64),
                                          (2, 16, 16, 512, 512), (2, 64, 64, 64, 64), (1, 16, 16, 64, 16),
                                          (4, 32, 16, 256, 256)])
@pytest.mark.parametrize("direction", ["fwd", "dgrad"])
def test_conv3x3_tc_matches_oracle(N, H, W, ci, co, direction):
    from models import _ops, _tc
    from models._lib import lib
    _ops.clear_weight_cache()
    torch.manual_seed(0)
    cin, cout = (ci, co) if direction == "fwd" else (co, ci)
    if not lib.sg_conv3x3_tc_supported(N, H, W, cin, cout):
        pytest.skip("shape not covered by the tensor-core kernel in this direction")
    x = torch.randn(N, cin, H, W).bfloat16().float()
    Wt = torch.randn(co, ci, 3, 3)
    scale = O.he_scale(ci * 9, math.sqrt(2))
    Wq = (Wt * scale).bfloat16().float()
    b = torch.randn(cout)
    if direction == "fwd":
        ref = F.conv2d(x, Wq, padding=1)
    else:
        ref = F.conv_transpose2d(x, Wq, padding=1)
    ref = F.leaky_relu(ref + b.view(1, -1, 1, 1), 0.2)
    xd = x.permute(0, 2, 3, 1).contiguous().to(DEV).bfloat16()
    Wd = torch.nn.Parameter(Wt.to(DEV))
    y = torch.empty((N, H, W, cout), dtype=torch.bfloat16, device=DEV)
    ok = _tc.try_conv(xd, Wd, "same", direction, scale, b.to(DEV), 1.0, True, y)
    assert ok
    torch.cuda.synchronize()
    got = y.float().permute(0, 3, 1, 2)
    assert rel(got, ref) < 6e-3          # bf16 output rounding only (fp32 accumulation)
    # and the generic kernel agrees
    _ops.set_use_tc(False)
    try:
        y2 = _ops._conv_raw(xd, Wd, "same", direction, scale, b.to(DEV), 1.0, True)
    finally:
        _ops.set_use_tc(True)
    assert rel(y2.float().permute(0, 3, 1, 2), ref) < 6e-3


def test_unsupported_shapes_fall_back_to_generic_kernel():
    from models._lib import lib
    assert lib.sg_conv3x3_tc_supported(2, 4, 4, 64, 64) == 0      # below the 8x16 tile
    assert lib.sg_conv3x3_tc_supported(2, 16, 16, 32, 64) == 1    # Cin < 64: TMA zero-fills the channel padding
    assert lib.sg_conv3x3_tc_supported(2, 16, 16, 12, 64) == 0    # global stride not a multiple of 16 bytes
    assert lib.sg_conv3x3_tc_supported(2, 16, 16, 64, 513) == 0


@pytest.mark.parametrize("N,H,W,ci,co", [(2, 16, 16, 64, 64), (1, 8, 32, 64, 128), (3, 32, 32, 128, 64),
                                          (2, 16, 16, 512, 512), (2, 64, 64, 64, 64), (4, 32, 16, 256, 128),
                                          (5, 8, 16, 128, 128)])
@pytest.mark.parametrize("v2", [True, False])
def test_wgrad3x3_tc_matches_oracle(N, H, W, ci, co, v2):
    """both generations of the weight-gradient kernel (csrc/sg_wgrad2_tc.cu: 128x192x16 MMAs; csrc/sg_wgrad_tc.cu)"""
    from models import _ops, _tc
    from models._lib import lib
    torch.manual_seed(1)
    assert lib.sg_wgrad3x3_tc_supported(N, H, W, ci, co) == 1
    assert lib.sg_wgrad3x3_v2_supported(N, H, W, ci, co) == 1
    _ops._state["wgrad_v2"] = v2
    x = torch.randn(N, ci, H, W).bfloat16().float()
    dy = torch.randn(N, co, H, W).bfloat16().float()
    scale = 0.37
    Wr = torch.zeros(co, ci, 3, 3, requires_grad=True)
    (ref,) = torch.autograd.grad(F.conv2d(x, Wr * scale, padding=1), Wr, dy)
    xd = x.permute(0, 2, 3, 1).contiguous().to(DEV).bfloat16()
    dyd = dy.permute(0, 2, 3, 1).contiguous().to(DEV).bfloat16()
    dW = torch.empty(co, ci, 3, 3, device=DEV)
    try:
        assert _tc.try_wgrad(xd, dyd, "same", scale, dW)
    finally:
        _ops._state["wgrad_v2"] = True
    torch.cuda.synchronize()
    assert rel(dW, ref) < 1e-4           # exact bf16 products, fp32 accumulation


@pytest.mark.parametrize("kind,ci,co,hw", [("up", 128, 64, 16), ("up_fused", 128, 64, 64), ("down", 64, 128, 32),
                                            ("down", 64, 128, 128), ("same", 128, 128, 4), ("same", 64, 128, 8),
                                            ("up", 128, 128, 4), ("down", 128, 64, 8),
                                            # channel counts below 64 (the 512x512 / 1024x1024 layers): TMA zero-fills
                                            ("same", 32, 32, 32), ("same", 16, 16, 64), ("up_fused", 64, 32, 64),
                                            ("up_fused", 32, 16, 64), ("down", 16, 32, 128), ("down", 32, 64, 64)])
def test_up_down_layers_on_tensor_cores(kind, ci, co, hw):
    """bf16 up/down EqualizedConv2d layers (nearest-up -> tc conv, tc conv -> box pool) vs the oracle, fwd + grads."""
    from models import _ops
    import test_gpu_ops as T
    torch.manual_seed(3)
    N = 2
    x = torch.randn(N, ci, hw, hw).bfloat16().float()
    scale = O.he_scale(ci * 9, math.sqrt(2))
    W = ((torch.randn(co, ci, 3, 3) * scale).bfloat16().float() / scale)
    b = torch.randn(co)
    xr, Wr, br = x.clone().requires_grad_(True), W.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y_ref = F.leaky_relu(T._ref_conv(kind, xr, Wr, scale) + br.view(1, -1, 1, 1), 0.2)
    gy = torch.randn_like(y_ref)
    ref = torch.autograd.grad(y_ref, [xr, Wr, br], gy)
    _ops.clear_weight_cache()
    launches0 = None
    xd = x.permute(0, 2, 3, 1).contiguous().to(DEV).bfloat16().requires_grad_(True)
    Wd, bd = torch.nn.Parameter(W.to(DEV)), torch.nn.Parameter(b.to(DEV))
    y = _ops.conv_layer(xd, Wd, bd, kind, scale, 1.0, True)
    assert rel(y.float().permute(0, 3, 1, 2), y_ref) < 1e-2
    got = torch.autograd.grad(y, [xd, Wd, bd], gy.permute(0, 2, 3, 1).contiguous().to(DEV).bfloat16())
    assert rel(got[0].float().permute(0, 3, 1, 2), ref[0]) < 6e-2
    assert rel(got[1], ref[1]) < 6e-2 and rel(got[2], ref[2]) < 6e-2


@pytest.mark.parametrize("kind", ["up", "up_fused", "down"])
@pytest.mark.parametrize("direction", ["fwd", "dgrad"])
@pytest.mark.parametrize("N,lo,ci,co", [(2, 16, 64, 64), (1, 32, 128, 64), (2, 16, 256, 128), (3, 64, 64, 128),
                                        (2, 16, 256, 256), (1, 8, 512, 512)])
@pytest.mark.parametrize("pair", [2, 1, 0])
def test_fused_updown_kernel_matches_oracle(kind, direction, N, lo, ci, co, pair):
    """sg_convud_tc (tap-merged 4-phase 2x2 / 4x4-stride-2 tcgen05 kernels; halo-box pair, first pair and single-CTA forms) vs torch on
    CPU, both directions."""
    import test_gpu_ops as T
    from models import _ops, _tc
    from models._lib import lib
    _ops.clear_weight_cache()
    torch.manual_seed(4)
    scale = O.he_scale(ci * 9, math.sqrt(2))
    W = ((torch.randn(co, ci, 3, 3) * scale).bfloat16().float() / scale)
    mode = _tc.ud_mode(kind, direction)
    cin, cout = (ci, co) if direction == "fwd" else (co, ci)
    if not lib.sg_convud_tc_supported(mode, N, lo, lo, cin, cout):
        pytest.skip("shape not covered")
    in_hw = lo if mode == _tc.UP else 2 * lo
    xin = torch.randn(N, cin, in_hw, in_hw).bfloat16().float()
    if direction == "fwd":
        ref = T._ref_conv(kind, xin, W, scale)
    else:
        # adjoint of the forward map applied to xin (a tensor shaped like the forward OUTPUT)
        fwd_in_hw = lo if kind != "down" else 2 * lo
        probe = torch.zeros(N, ci, fwd_in_hw, fwd_in_hw, requires_grad=True)
        (ref,) = torch.autograd.grad(T._ref_conv(kind, probe, W, scale), probe, xin)
    xd = xin.permute(0, 2, 3, 1).contiguous().to(DEV).bfloat16()
    Wd = torch.nn.Parameter(W.to(DEV))
    y = torch.empty((N, ref.shape[2], ref.shape[3], cout), dtype=torch.bfloat16, device=DEV)
    old = lib.sg_pair_updown(pair)
    try:
        assert _tc.try_conv(xd, Wd, kind, direction, scale, None, 1.0, False, y)      # the fused tensor-core path
    finally:
        lib.sg_pair_updown(old)
    torch.cuda.synchronize()
    assert rel(y.float().permute(0, 3, 1, 2), ref) < 1e-2
