"""Host logic: the tap-table geometries/coefficients reproduce the reference's convolutions and their adjoints."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import stylegan_oracle as O
from models import _geom as G


def effective(W, t, scale):
    """E[pt][a][b] as sg_wprep_fwd builds it."""
    co, ci, k, _ = W.shape
    coef = torch.from_numpy(t["coef"]).to(W.dtype)           # [npt, KK]
    E = torch.einsum("pj,oij->pio", coef, W.reshape(co, ci, k * k)) * scale   # [npt, ci, co]
    return E.transpose(1, 2).contiguous() if t["transpose"] else E


def ref_forward(kind, x, W, scale):
    ws = W * scale
    if kind == "same":
        return F.conv2d(x, ws, padding=W.shape[-1] // 2)
    if kind == "same_flip":
        return F.conv2d(x, ws.flip(2, 3), padding=W.shape[-1] // 2)
    if kind == "down":
        return F.avg_pool2d(F.conv2d(x, ws, padding=1), 2)
    if kind == "up":
        return F.conv2d(O.nearest_up2(x), ws, padding=1)
    wt = F.pad(ws.permute(1, 0, 2, 3), [1, 1, 1, 1])
    wt = wt[:, :, 1:, 1:] + wt[:, :, :-1, 1:] + wt[:, :, 1:, :-1] + wt[:, :, :-1, :-1]
    return F.conv_transpose2d(x, wt, stride=2, padding=1)


@pytest.mark.parametrize("kind,k", [("same", 3), ("same", 1), ("same_flip", 3), ("down", 3), ("up", 3), ("up_fused", 3)])
def test_forward_and_dgrad_tables(kind, k):
    torch.manual_seed(0)
    N, Ci, Co, H, W_ = 2, 5, 7, 8, 6
    x = torch.randn(N, Ci, H, W_, dtype=torch.float64, requires_grad=True)
    W = torch.randn(Co, Ci, k, k, dtype=torch.float64)
    scale = 0.37
    y_ref = ref_forward(kind, x, W, scale)
    t = G.tables(kind, k, "fwd")
    Ho, Wo = G.out_hw(kind, "fwd", H, W_)
    assert (Ho, Wo) == tuple(y_ref.shape[2:])
    y = G.conv_by_taps(x.detach().permute(0, 2, 3, 1), effective(W, t, scale).double(), t, (Ho, Wo))
    assert torch.allclose(y.permute(0, 3, 1, 2), y_ref, atol=1e-10)
    # adjoint
    gy = torch.randn_like(y_ref)
    (gx_ref,) = torch.autograd.grad(y_ref, x, gy)
    td = G.tables(kind, k, "dgrad")
    assert G.out_hw(kind, "dgrad", Ho, Wo) == (H, W_)
    gx = G.conv_by_taps(gy.permute(0, 2, 3, 1), effective(W, td, scale).double(), td, (H, W_))
    assert torch.allclose(gx.permute(0, 3, 1, 2), gx_ref, atol=1e-10)


def test_down_equals_fused_branch():
    """Both branches of the reference's downscale conv (CustomLayers.py:158-180) are the same function."""
    torch.manual_seed(1)
    x = torch.randn(1, 3, 128, 128)
    w = torch.randn(4, 3, 3, 3)
    b = torch.randn(4)
    fused = O.down_conv(x, w, b, 1.0)
    unfused = F.avg_pool2d(F.conv2d(x, w * O.he_scale(27, 1.0), padding=1), 2) + b.view(1, -1, 1, 1)
    assert torch.allclose(fused, unfused, atol=1e-5)


def test_geom_struct_fields():
    g = G.make_geom("up_fused", 3, "fwd", 2, 64, 64, 8, 4)
    assert (g.Hq, g.Wq, g.Hout, g.Wout, g.in_stride, g.out_stride, g.nphase, g.ntaps) == (64, 64, 128, 128, 1, 2, 4, 4)
    g = G.make_geom("down", 3, "fwd", 2, 64, 64, 8, 4)
    assert (g.Hq, g.Hout, g.in_stride, g.out_stride, g.nphase, g.ntaps) == (32, 32, 2, 1, 1, 16)
    assert [g.dy[0][i] for i in range(16)] == [u - 1 for u in range(4) for _ in range(4)]


def test_up_fused_is_flipped_conv_after_nearest_up():
    """The identity the bf16 tensor-core path relies on: conv_transpose2d of the box-summed kernel ==
    nearest x2 followed by a stride-1 conv with the 180-degree rotated kernel."""
    torch.manual_seed(2)
    x = torch.randn(2, 5, 64, 64, dtype=torch.float64)
    w = torch.randn(7, 5, 3, 3, dtype=torch.float64)
    fused = ref_forward("up_fused", x, w, 1.0)
    alt = F.conv2d(O.nearest_up2(x), w.flip(2, 3), padding=1)
    assert torch.allclose(fused, alt, atol=1e-10)
