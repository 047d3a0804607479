"""Host-side geometry of the tap-table convolution family (see csrc/sg_conv_generic.cu).

Every dense convolution of the reference and each of its adjoints is described by
    out[n, qy*os+py, qx*os+px] = sum_t in[n, qy*is+dy[ph][t], qx*is+dx[ph][t]] . E[ph][t]
with effective weights  E[ph][t] = scale * sum_j coef[ph][t][j] * W[:, :, j]  built from the 3x3 (or 1x1)
parameter.  Layer kinds (reference models/CustomLayers.py):
    'same'      kxk stride-1 conv                                   :170-173
    'down'      conv3x3 -> 2x2 box  ==  4x4 stride-2 averaged kernel :156-180 (both branches are the same function)
    'up'        nearest x2 -> conv3x3                                :153-154,172-173   (output < 128)
    'up_fused'  conv_transpose2d of the box-summed 4x4 kernel        :143-152           (output >= 128)
                == 'up' with the 3x3 kernel rotated by 180 degrees
This module is pure Python/numpy so the tables are unit-tested on CPU (tests/test_geom.py)."""
import functools

import numpy as np

from ._lib import SgConvGeom

KINDS = ("same", "same_flip", "down", "up", "up_fused")


def _merge_1d(kind):
    """M[p][a] -> set of kernel rows kh folded into tap a of output phase p (nearest-up then 3-tap conv)."""
    m = {(0, 0): (0,), (0, 1): (1, 2), (1, 0): (0, 1), (1, 1): (2,)}
    if kind == "up_fused":
        m = {k: tuple(2 - kh for kh in v) for k, v in m.items()}
    return m


@functools.lru_cache(maxsize=None)
def tables(kind, k, direction):
    """-> dict(in_stride, out_stride, nphase, ntaps, dy[nphase][ntaps], dx, coef[nphase*ntaps][k*k], transpose)

    direction 'fwd' maps the layer input to its output; 'dgrad' maps the output gradient to the input gradient
    (channels swap: E is laid out [.., Cout, Cin] -> transpose=1)."""
    assert kind in KINDS and direction in ("fwd", "dgrad")
    KK = k * k
    if kind in ("same", "same_flip"):
        # 'same_flip' = stride-1 conv with the kernel rotated by 180 degrees (the fused up-conv after a nearest x2)
        taps = [(kh, kw) for kh in range(k) for kw in range(k)]
        sgn = 1 if direction == "fwd" else -1
        dy = [[sgn * (kh - k // 2) for kh, kw in taps]]
        dx = [[sgn * (kw - k // 2) for kh, kw in taps]]
        coef = np.eye(KK, dtype=np.float32)
        if kind == "same_flip":
            coef = coef[::-1].copy()
        return dict(in_stride=1, out_stride=1, nphase=1, ntaps=KK, dy=dy, dx=dx, coef=coef,
                    transpose=int(direction == "dgrad"))
    assert k == 3
    # 4x4 stride-2 form: tap (u,v), offset d = u-1; coefficient of W[kh][kw]
    def down_coef(u, v):
        c = np.zeros(9, np.float32)
        for kh in range(3):
            for kw in range(3):
                if (u - kh) in (0, 1) and (v - kw) in (0, 1):
                    c[kh * 3 + kw] = 0.25
        return c

    def up_coef(py, a, px, b, kind_):
        m = _merge_1d(kind_)
        c = np.zeros(9, np.float32)
        for kh in m[(py, a)]:
            for kw in m[(px, b)]:
                c[kh * 3 + kw] = 1.0
        return c

    phases = [(py, px) for py in range(2) for px in range(2)]
    taps22 = [(a, b) for a in range(2) for b in range(2)]
    taps44 = [(u, v) for u in range(4) for v in range(4)]
    if (kind == "down" and direction == "fwd"):
        return dict(in_stride=2, out_stride=1, nphase=1, ntaps=16, dy=[[u - 1 for u, v in taps44]],
                    dx=[[v - 1 for u, v in taps44]], coef=np.stack([down_coef(u, v) for u, v in taps44]), transpose=0)
    if kind in ("up", "up_fused") and direction == "fwd":
        dy = [[a + py - 1 for a, b in taps22] for py, px in phases]
        dx = [[b + px - 1 for a, b in taps22] for py, px in phases]
        coef = np.stack([up_coef(py, a, px, b, kind) for py, px in phases for a, b in taps22])
        return dict(in_stride=1, out_stride=2, nphase=4, ntaps=4, dy=dy, dx=dx, coef=coef, transpose=0)
    if kind == "down" and direction == "dgrad":
        # dx[2m+p] = sum_a W4[u(p,a)]^T dy[m + a + p - 1],  u = 3 - p - 2a
        dy = [[a + py - 1 for a, b in taps22] for py, px in phases]
        dx = [[b + px - 1 for a, b in taps22] for py, px in phases]
        coef = np.stack([down_coef(3 - py - 2 * a, 3 - px - 2 * b) for py, px in phases for a, b in taps22])
        return dict(in_stride=1, out_stride=2, nphase=4, ntaps=4, dy=dy, dx=dx, coef=coef, transpose=1)
    # up / up_fused dgrad: dx[j] = sum_{p,a} U[p,a]^T dy[2j + (2 - 2a - p)]  ->  tap u = 3 - 2a - p
    inv = {3 - 2 * a - p: (p, a) for p in range(2) for a in range(2)}
    coef = np.stack([up_coef(inv[u][0], inv[u][1], inv[v][0], inv[v][1], kind) for u, v in taps44])
    return dict(in_stride=2, out_stride=1, nphase=1, ntaps=16, dy=[[u - 1 for u, v in taps44]],
                dx=[[v - 1 for u, v in taps44]], coef=coef, transpose=1)


def out_hw(kind, direction, H, W):
    up = kind in ("up", "up_fused")
    if kind in ("same", "same_flip"):
        return H, W
    grow = (up and direction == "fwd") or (kind == "down" and direction == "dgrad")
    return (2 * H, 2 * W) if grow else (H // 2, W // 2)


def make_geom(kind, k, direction, N, H, W, Cin, Cout):
    """SgConvGeom for an input [N,H,W,Cin] -> output [N,Ho,Wo,Cout] (Cin/Cout are those of THIS map)."""
    t = tables(kind, k, direction)
    Ho, Wo = out_hw(kind, direction, H, W)
    g = SgConvGeom()
    g.N, g.Hin, g.Win, g.Cin, g.Hout, g.Wout, g.Cout = N, H, W, Cin, Ho, Wo, Cout
    g.in_stride, g.out_stride, g.nphase, g.ntaps = t["in_stride"], t["out_stride"], t["nphase"], t["ntaps"]
    g.Hq, g.Wq = Ho // t["out_stride"], Wo // t["out_stride"]
    for ph in range(t["nphase"]):
        for i in range(t["ntaps"]):
            g.dy[ph][i] = t["dy"][ph][i]
            g.dx[ph][i] = t["dx"][ph][i]
    return g


def conv_by_taps(x_nhwc, E, t, out_hw_):
    """Reference evaluation of the tap-table formula with torch ops (CPU tests only; never on the product path).
    x [N,H,W,Ci], E [nphase*ntaps, Ci, Co] -> [N,Ho,Wo,Co]."""
    import torch
    N, H, W, Ci = x_nhwc.shape
    Ho, Wo = out_hw_
    Co = E.shape[-1]
    os_, is_ = t["out_stride"], t["in_stride"]
    Hq, Wq = Ho // os_, Wo // os_
    out = torch.zeros(N, Ho, Wo, Co, dtype=x_nhwc.dtype)
    pad = 4
    xp = torch.nn.functional.pad(x_nhwc, (0, 0, pad, pad + 4, pad, pad + 4))
    for ph in range(t["nphase"]):
        py, px = ph >> 1, ph & 1
        acc = torch.zeros(N, Hq, Wq, Co, dtype=x_nhwc.dtype)
        for i in range(t["ntaps"]):
            dy, dx = t["dy"][ph][i], t["dx"][ph][i]
            ys = torch.arange(Hq) * is_ + dy + pad
            xs = torch.arange(Wq) * is_ + dx + pad
            patch = xp[:, ys][:, :, xs]                       # [N,Hq,Wq,Ci]
            acc = acc + patch @ E[ph * t["ntaps"] + i]
        out[:, py::os_, px::os_] = acc
    return out
