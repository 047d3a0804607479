"""Dispatch to the tcgen05 tensor-core convolution kernels (csrc/sg_conv_tc.cu) for bf16 activations.

Returns False when the layer shape is not covered; the caller then runs the generic CUDA-core kernel.  Both are
hand-written sm_100a kernels of this library -- there is no library or CPU fallback."""
import numpy as np
import torch

from ._lib import lib, ptr, stream, count_launch, BF16
from . import _ops

_coef = {}


def _tc_coef(flipped, device):
    key = (bool(flipped), device)
    c = _coef.get(key)
    if c is None:
        m = np.eye(9, dtype=np.float32)
        if flipped:
            m = m[::-1].copy()
        c = torch.from_numpy(m).to(device)
        _coef[key] = c
    return c


def tc_weight(W, direction, scale, kind="same"):
    """fwd: E[t][co][ci] = s*W[co][ci][t];  dgrad: E[t][ci][co] = s*W[co][ci][8-t] (180-degree rotated taps);
    kind 'same_flip' rotates once more."""
    def build(W, out):
        co, ci = W.shape[0], W.shape[1]
        rows, cols = (co, ci) if direction == "fwd" else (ci, co)
        flipped = (direction == "dgrad") != (kind == "same_flip")
        E = out if out is not None else torch.empty((9, rows, cols), dtype=torch.bfloat16, device=W.device)
        _ops.wprep_fwd(W, E, torch.bfloat16, co, ci, 9, 9, _tc_coef(flipped, W.device), scale,
                       1 if direction == "fwd" else 0)
        return E
    return _ops.cached_weight(W, ("tc3x3", kind, direction, float(scale)), build)


def conv_and_wgrad_supported(N, H, W, ci, co):
    return bool(lib.sg_conv3x3_tc_supported(N, H, W, ci, co)) and bool(lib.sg_conv3x3_tc_supported(N, H, W, co, ci)) \
        and bool(lib.sg_wgrad3x3_tc_supported(N, H, W, ci, co))


UP, DOWN = 0, 1


def ud_mode(kind, direction):
    """Geometry of the fused up/down kernel that evaluates (kind, direction)."""
    if kind in ("up", "up_fused"):
        return UP if direction == "fwd" else DOWN
    return DOWN if direction == "fwd" else UP


def ud_weight(W, kind, direction, scale):
    """E[16][out][in] bf16 for sg_convud_tc from the host tap tables (models/_geom.py)."""
    from . import _geom

    def build(W, out):
        co, ci = W.shape[0], W.shape[1]
        t = _geom.tables(kind, 3, direction)
        rows, cols = (co, ci) if direction == "fwd" else (ci, co)
        E = out if out is not None else torch.empty((16, rows, cols), dtype=torch.bfloat16, device=W.device)
        coef = _ops._coef_dev(kind, 3, direction, W.device)
        _ops.wprep_fwd(W, E, torch.bfloat16, co, ci, 9, 16, coef, scale, 1 - t["transpose"])
        return E
    return _ops.cached_weight(W, ("tcud", kind, direction, float(scale)), build)


def updown_supported(kind, N, Hlo, Wlo, ci, co):
    """Fused up/down layer on the tensor cores: forward, dgrad and (via the 3x3 wgrad at high resolution) wgrad."""
    m_f, m_d = ud_mode(kind, "fwd"), ud_mode(kind, "dgrad")
    return bool(lib.sg_convud_tc_supported(m_f, N, Hlo, Wlo, ci, co)) and \
        bool(lib.sg_convud_tc_supported(m_d, N, Hlo, Wlo, co, ci)) and \
        bool(lib.sg_wgrad3x3_tc_supported(N, 2 * Hlo, 2 * Wlo, ci, co))


def try_conv(x, W, kind, direction, scale, bias, bias_mul, act, y):
    if not _ops._state["use_tc"] or x.dtype != torch.bfloat16:
        return False
    if kind in ("up", "up_fused", "down") and W.shape[2] == 3 and _ops._state.get("fused_updown", True):
        co, ci = W.shape[0], W.shape[1]
        cin, cout = (ci, co) if direction == "fwd" else (co, ci)
        mode = ud_mode(kind, direction)
        N, H, Wd, _ = x.shape
        Hlo, Wlo = (H, Wd) if mode == UP else (H // 2, Wd // 2)
        if not lib.sg_convud_tc_supported(mode, N, Hlo, Wlo, cin, cout):
            return False
        E = ud_weight(W, kind, direction, scale)
        lib.sg_convud_tc(mode, ptr(x), ptr(E), ptr(y), ptr(bias), float(bias_mul), int(act), N, Hlo, Wlo, cin, cout,
                         stream())
        count_launch()
        return True
    if kind not in ("same", "same_flip"):
        return False
    co, ci, k, _ = W.shape
    if k != 3:
        return False
    N, H, Wd, _ = x.shape
    cin, cout = (ci, co) if direction == "fwd" else (co, ci)
    if not lib.sg_conv3x3_tc_supported(N, H, Wd, cin, cout):
        return False
    E = tc_weight(W, direction, scale, kind)
    if _ops._state.get("use_pair", True) and lib.sg_conv3x3_pair_supported(N, H, Wd, cin, cout):
        # CTA-pair (cta_group::2) kernel: 1.6x the MMA issue rate of the single-CTA form at Cout = 64
        lib.sg_conv3x3_pair(ptr(x), ptr(E), ptr(y), ptr(bias), float(bias_mul), int(act), None, None, None, N, H, Wd,
                            cin, cout, stream())
    else:
        lib.sg_conv3x3_tc(ptr(x), ptr(E), ptr(y), ptr(bias), float(bias_mul), int(act), N, H, Wd, cin, cout, stream())
    count_launch()
    return True


def conv_epilogue_supported(x, W):
    """conv3x3 (stride 1) + noise + LeakyReLU + InstanceNorm statistics in ONE kernel (csrc/sg_conv_pair_tc.cu)."""
    if not _ops._state["use_tc"] or not _ops._state.get("use_pair", True) or not _ops._state.get("fuse_epilogue", True):
        return False
    if x.dtype != torch.bfloat16 or W.shape[2] != 3:
        return False
    N, H, Wd, cin = x.shape
    co, ci = W.shape[0], W.shape[1]
    return ci == cin and bool(lib.sg_conv3x3_pair_supported(N, H, Wd, ci, co)) and \
        bool(lib.sg_conv3x3_tc_supported(N, H, Wd, co, ci)) and bool(lib.sg_wgrad3x3_tc_supported(N, H, Wd, ci, co))


def conv_epilogue_forward(x, W, scale, bias, bias_mul, noise, noise_w, style, eps):
    """-> (z, mr, ab): z = lrelu(conv(x) + bias + noise_w*noise) in bf16, mr = (mean, rstd) per (n,c), ab = the folded
    InstanceNorm + StyleMod coefficients for sg_affine_nc."""
    N, H, Wd, cin = x.shape
    co = W.shape[0]
    E = tc_weight(W, "fwd", scale, "same")
    z = torch.empty((N, H, Wd, co), dtype=torch.bfloat16, device=x.device)
    slots = lib.sg_conv3x3_pair_slots(N, H, Wd, cin, co)
    part = torch.empty((N, slots, 2, co), dtype=torch.float32, device=x.device)
    lib.sg_conv3x3_pair(ptr(x), ptr(E), ptr(z), ptr(bias), float(bias_mul), 1, ptr(noise),
                        ptr(noise_w if noise is not None else None), ptr(part), N, H, Wd, cin, co, stream())
    mr = torch.empty((N, co, 2), dtype=torch.float32, device=x.device)
    ab = torch.empty((N, co, 3), dtype=torch.float32, device=x.device)
    lib.sg_epi_finalize_parts(ptr(part), ptr(style), ptr(mr), ptr(ab), N, H, Wd, cin, co, float(eps), stream())
    count_launch(2)
    return z, mr, ab


def try_wgrad(xl, dyl, kind, scale, dW):
    """dW[co][ci][kh][kw] = scale * sum_p dy[p][co] x[p + (kh-1,kw-1)][ci] on the tensor cores."""
    if not _ops._state["use_tc"] or xl.dtype != torch.bfloat16:
        return False
    co, ci, k, _ = dW.shape
    if k != 3:
        return False
    if kind in ("up", "up_fused", "down"):
        # the up/down layers' weight gradient through the 3x3 kernel at the high resolution:
        #   up:   dW = wgrad3x3(nearest_up2(x), dy)   (180-degree rotated taps for 'up_fused')
        #   down: dW = wgrad3x3(x, nearest_up2(dy)/4)
        N, H, Wd, _ = xl.shape
        Hh, Wh = (2 * H, 2 * Wd) if kind != "down" else (H, Wd)
        mode = DOWN if kind == "down" else UP
        if _ops._state.get("fused_updown", True) and lib.sg_wgradud_tc_supported(mode, N, Hh // 2, Wh // 2, ci, co):
            # tap-merged kernel: 16 tap accumulators, stride-2 TMA gathers, then the table's adjoint onto the 3x3 weight
            S = lib.sg_wgradud_tc_splits(mode, N, Hh // 2, Wh // 2, ci, co)
            part = torch.empty((S, 16, co, ci), dtype=torch.float32, device=xl.device)
            lib.sg_wgradud_tc(mode, ptr(xl), ptr(dyl), ptr(part), N, Hh // 2, Wh // 2, ci, co, stream())
            coef = _ops._coef_dev(kind, 3, "fwd", xl.device)
            lib.sg_wprep_bwd(ptr(part), S, ptr(dW), co, ci, 9, 16, ptr(coef), float(scale), 1, 0, stream())
            count_launch(2)
            return True
        if not lib.sg_wgrad3x3_tc_supported(N, Hh, Wh, ci, co):
            return False
        if kind == "down":
            gu = torch.empty((N, Hh, Wh, co), dtype=dyl.dtype, device=dyl.device)
            lib.sg_up2_nhwc(ptr(dyl), ptr(gu), 0.25, BF16, N, Hh // 2, Wh // 2, co, stream())
            count_launch()
            return try_wgrad(xl, gu, "same", scale, dW)
        xu = torch.empty((N, Hh, Wh, ci), dtype=xl.dtype, device=xl.device)
        lib.sg_up2_nhwc(ptr(xl), ptr(xu), 1.0, BF16, N, H, Wd, ci, stream())
        count_launch()
        return try_wgrad(xu, dyl, "same" if kind == "up" else "same_flip", scale, dW)
    if kind not in ("same", "same_flip"):
        return False
    N, H, Wd, _ = xl.shape
    if not lib.sg_wgrad3x3_tc_supported(N, H, Wd, ci, co):
        return False
    if _ops._state.get("wgrad_v2", True) and lib.sg_wgrad3x3_v2_supported(N, H, Wd, ci, co):
        S = lib.sg_wgrad3x3_v2_splits(N, H, Wd, ci, co)
        part = torch.empty((S, 9, co, ci), dtype=torch.float32, device=xl.device)
        lib.sg_wgrad3x3_v2(ptr(xl), ptr(dyl), ptr(part), N, H, Wd, ci, co, stream())
    else:
        S = lib.sg_wgrad3x3_tc_splits(N, H, Wd, ci, co)
        part = torch.empty((S, 9, co, ci), dtype=torch.float32, device=xl.device)
        lib.sg_wgrad3x3_tc(ptr(xl), ptr(dyl), ptr(part), N, H, Wd, ci, co, stream())
    lib.sg_wprep_bwd(ptr(part), S, ptr(dW), co, ci, 9, 9, ptr(_tc_coef(kind == "same_flip", xl.device)), float(scale),
                     1, 0, stream())
    count_launch(2)
    return True
