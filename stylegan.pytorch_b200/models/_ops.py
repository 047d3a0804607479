"""torch.autograd.Function wrappers over the libsgb200 C-ABI.

Activations are NHWC-contiguous tensors [N,H,W,C] in the engine dtype (fp32 or bf16); parameters, images and
statistics are fp32.  Every Function on the discriminator path has a backward built from other Functions, so the R1
gradient penalty (reference models/Losses.py:197-210, autograd.grad(create_graph=True)) differentiates through the
backward pass without PyTorch having to see inside a kernel.  Generator-only Functions are once-differentiable.

PyTorch is plumbing here: it owns memory, streams and the autograd tape; all arithmetic is in the kernels."""
import ctypes
import weakref

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _geom
from ._lib import lib, ptr, stream, code, count_launch

_state = {"dtype": torch.float32, "use_tc": True, "skip_param_grads": False, "fused_updown": True, "use_pair": True,
          "fuse_epilogue": True, "wgrad_v2": True}


class skip_param_grads:
    """Context manager for a backward pass whose parameter gradients are not wanted (the first, create_graph=True
    backward of the R1 penalty differentiates wrt the image only).  autograd cannot tell a custom Function which of
    its gradient outputs will be consumed, so without this every layer would also compute a weight gradient there."""

    def __enter__(self):
        self._old = _state["skip_param_grads"]
        _state["skip_param_grads"] = True

    def __exit__(self, *a):
        _state["skip_param_grads"] = self._old


def _want_params():
    return not _state["skip_param_grads"]


def set_precision(p):
    """'fp32' (parity mode, CUDA-core kernels) or 'bf16' (tcgen05 tensor-core kernels, fp32 accumulation)."""
    _state["dtype"] = {"fp32": torch.float32, "bf16": torch.bfloat16, torch.float32: torch.float32,
                       torch.bfloat16: torch.bfloat16}[p]


def precision():
    return _state["dtype"]


def set_use_tc(flag):
    _state["use_tc"] = bool(flag)


# ---------------------------------------------------------------------------------------------------------------
# per-step values that must not be baked into a captured CUDA graph (models/_graphs.py)
# ---------------------------------------------------------------------------------------------------------------
class DynScalar:
    """A per-step float pair (v, w) in device memory -- e.g. the fade-in (alpha, 1 - alpha).  Axpby(x, y, d, None)
    computes v*x + w*y with both read by the kernel at run time; `float(d)` is the host value of v."""

    def __init__(self, dev, value):
        self.dev, self.value = dev, float(value)           # dev: float32 tensor with >= 1 element

    def __float__(self):
        return self.value

    @property
    def second(self):
        """the DynScalar whose first element is this one's w"""
        return DynScalar(self.dev[1:], float("nan"))


class DynInt:
    """A per-step int in device memory (the style-mixing cut-off)."""

    def __init__(self, dev, value):
        self.dev, self.value = dev, int(value)             # dev: int32 tensor [1]

    def __int__(self):
        return self.value


_capture = [None]          # the active models/_graphs.py capture context, if any


def capture_ctx():
    return _capture[0]


def upload_table(t, device):
    """Small CPU tensor -> device.  During a CUDA-graph capture the copy must come from pinned memory that outlives the
    graph (the copy node re-reads the host address at every replay): the capture context owns such an arena."""
    c = _capture[0]
    if c is not None:
        return c.pinned_copy(t).to(device, non_blocking=True)
    return t.to(device, non_blocking=True)


def lerp_coefs(alpha):
    """(a, b) for a*x + b*y = lerp(y, x, alpha): python floats, or (DynScalar, None) when alpha lives on the device."""
    if isinstance(alpha, DynScalar):
        return alpha, None
    return float(alpha), 1.0 - float(alpha)


def _require_cuda(t):
    if not t.is_cuda:
        raise RuntimeError("sgb200: the StyleGAN hot path runs on CUDA tensors only (no CPU fallback); got a %s "
                           "tensor" % t.device)


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


# ---------------------------------------------------------------------------------------------------------------
# weight preparation cache
# ---------------------------------------------------------------------------------------------------------------
_coef_cache = {}
_prep_cache = {}


def _coef_dev(kind, k, direction, device):
    key = (kind, k, direction, device)
    c = _coef_cache.get(key)
    if c is None:
        c = torch.from_numpy(_geom.tables(kind, k, direction)["coef"]).to(device).contiguous()
        _coef_cache[key] = c
    return c


def _token_of(W):
    """Update counter of whoever rewrites this parameter behind autograd's back: the FusedAdam that owns it or the EMA
    updater (they write through raw pointers, so `_version` does not move); the global epoch for anything else."""
    tok = getattr(W, "_sg_token", None)
    return (tok[0] if tok is not None else -1, _epoch[0])


def cached_weight(W, sub, build):
    """Derived (packed / scaled / bf16) copy of a weight, kept for nn.Parameters only, in a PERSISTENT buffer per
    (parameter object, variant): `build(W, out)` fills `out` (or allocates when out is None); it must not
    capture W (the entry would keep the parameter alive).  An entry belongs to the
    Parameter OBJECT (weak reference: ids and device pointers are recycled once a model is freed) at one (storage,
    version, shape).  The optimizer that rewrites the parameter refreshes all of its variants in place right after the
    update (refresh_derived), so a forward pass normally finds them current and launches nothing -- which is also what
    lets a captured CUDA graph read them: they are refreshed by the graph that contains the optimizer step.  A variant
    whose stamp is older than the parameter's update counter (EMA shadow, global epoch bump) is rebuilt in place on use.
    Code that edits `param.data` behind autograd's back must call clear_weight_cache()."""
    if not isinstance(W, torch.nn.Parameter):
        return build(W, None)
    key = (W.data_ptr(), W._version, tuple(W.shape))
    wid = id(W)
    ent = _prep_cache.get(wid)
    if ent is None or ent[2]() is not W or ent[0] != key:
        def _drop(ref, wid=wid):
            cur = _prep_cache.get(wid)
            if cur is not None and cur[2] is ref:
                del _prep_cache[wid]
        ent = (key, {}, weakref.ref(W, _drop))
        _prep_cache[wid] = ent
    tok = _token_of(W)
    it = ent[1].get(sub)
    if it is None:
        _last_wprep[0] = None
        E = build(W, None)
        rec = _last_wprep[0]                 # the build was one sg_wprep_fwd into E: remember its arguments (batched refresh)
        desc = rec[1] if (rec is not None and isinstance(E, torch.Tensor) and rec[0] == E.data_ptr()) else None
        ent[1][sub] = [E, build, tok, desc]
        return E
    if it[2] != tok:
        build(W, it[0])
        it[2] = tok
    return it[0]


_last_wprep = [None]
_wprep_tables = {}


def wprep_fwd(W, E, dtype, co, ci, KK, npt, coef, scale, transpose):
    """sg_wprep_fwd, recording its arguments for cached_weight (refresh_derived re-issues all recorded bf16 preparations
    of an optimizer's parameters as ONE sg_wprep_fwd_multi launch)."""
    Wc = _c(W.detach())
    lib.sg_wprep_fwd(ptr(Wc), ptr(E), code(dtype), co, ci, KK, npt, ptr(coef), float(scale), int(transpose), stream())
    count_launch()
    if dtype == torch.bfloat16 and Wc.data_ptr() == W.data_ptr() and W.dtype == torch.float32:
        _last_wprep[0] = (E.data_ptr(), (W.data_ptr(), E.data_ptr(), coef.data_ptr(), int(co), int(ci), int(KK), int(npt),
                                         int(transpose), float(scale), coef))
    else:
        _last_wprep[0] = None


def _wprep_multi(descs, device):
    """One launch for a list of recorded sg_wprep_fwd calls; the device table is cached per exact list."""
    import numpy as np
    key = tuple(d[:9] for d in descs)
    ent = _wprep_tables.get(key)
    if ent is None:
        rows = np.zeros(len(descs), dtype=np.dtype([("W", "<i8"), ("E", "<i8"), ("coef", "<i8"), ("Cout", "<i4"),
                                                    ("Cin", "<i4"), ("KK", "<i4"), ("npt", "<i4"), ("transpose", "<i4"),
                                                    ("tile0", "<i4"), ("scale", "<f4"), ("pad", "<i4")]))
        assert rows.dtype.itemsize == 56
        tile0, smem = 0, 0
        for i, (Wp, Ep, Cp, co, ci, KK, npt, tr, scale, _coef) in enumerate(descs):
            rows[i] = (Wp, Ep, Cp, co, ci, KK, npt, tr, tile0, scale, 0)
            tile0 += ((co + 31) // 32) * ((ci + 31) // 32)
            smem = max(smem, npt * KK + 32 * (32 * KK + 1))
        t = torch.from_numpy(rows.view(np.uint8).copy())
        if len(_wprep_tables) > 64:
            _wprep_tables.clear()
        ent = _wprep_tables[key] = (upload_table(t, device), tile0, smem)
    lib.sg_wprep_fwd_multi(ptr(ent[0]), len(descs), ent[1], ent[2], stream())
    count_launch()


def refresh_derived(params):
    """Rebuild, in place, every cached variant of the given parameters (called by the optimizer right after it has
    rewritten them and bumped their update counter)."""
    batch, dev = [], None
    for W in params:
        ent = _prep_cache.get(id(W))
        if ent is None or ent[2]() is not W:
            continue
        if ent[0] != (W.data_ptr(), W._version, tuple(W.shape)):
            del _prep_cache[id(W)]
            continue
        tok = _token_of(W)
        for it in ent[1].values():
            if it[3] is not None and it[3][0] == W.data_ptr() and it[3][1] == it[0].data_ptr():
                batch.append(it[3])                     # a plain sg_wprep_fwd into a bf16 buffer: goes into the one launch
            else:
                it[1](W, it[0])
            it[2] = tok
            dev = W.device
    if batch:
        _wprep_multi(batch, dev)


def prepped_weight(W, kind, direction, dtype, scale):
    """Effective weights E[npt][A][B] of the tap-table convolution in `dtype`."""
    def build(W, out):
        co, ci, k, _ = W.shape
        t = _geom.tables(kind, k, direction)
        npt = t["nphase"] * t["ntaps"]
        a, b = (co, ci) if t["transpose"] else (ci, co)
        E = out if out is not None else torch.empty((npt, a, b), dtype=dtype, device=W.device)
        coef = _coef_dev(kind, k, direction, W.device)
        wprep_fwd(W, E, dtype, co, ci, k * k, npt, coef, scale, t["transpose"])
        return E
    return cached_weight(W, (kind, direction, dtype, float(scale)), build)


_epoch = [0]


def bump_weights_epoch():
    """Called by the fused optimizer after it has rewritten parameters in place."""
    _epoch[0] += 1


def clear_weight_cache():
    _prep_cache.clear()
    _wprep_tables.clear()


# ---------------------------------------------------------------------------------------------------------------
# dense convolutions
# ---------------------------------------------------------------------------------------------------------------
def _conv_raw(x, W, kind, direction, scale, bias=None, bias_mul=1.0, act=False):
    _require_cuda(x)
    x = _c(x)
    co, ci, k, _ = W.shape
    N, H, Wd, C = x.shape
    cin, cout = (ci, co) if direction == "fwd" else (co, ci)
    assert C == cin, "channel mismatch: activation has %d, weight expects %d" % (C, cin)
    Ho, Wo = _geom.out_hw(kind, direction, H, Wd)
    y = torch.empty((N, Ho, Wo, cout), dtype=x.dtype, device=x.device)
    from . import _tc
    if _tc.try_conv(x, W, kind, direction, scale, bias, bias_mul, act, y):
        return y
    E = prepped_weight(W, kind, direction, x.dtype, scale)
    g = _geom.make_geom(kind, k, direction, N, H, Wd, cin, cout)
    lib.sg_conv_generic(ptr(x), ptr(E), ptr(y), ptr(bias), float(bias_mul), int(act), code(x.dtype),
                        ctypes.addressof(g), stream())
    count_launch()
    return y


class Conv(Function):
    """y = conv_{kind,direction}(x; scale*W).  direction 'dgrad' is the adjoint map (used by backward and by R1)."""

    @staticmethod
    def forward(ctx, x, W, kind, direction, scale):
        ctx.save_for_backward(x, W)
        ctx.cfg = (kind, direction, scale)
        return _conv_raw(x, W, kind, direction, scale)

    @staticmethod
    def backward(ctx, gy):
        x, W = ctx.saved_tensors
        kind, direction, scale = ctx.cfg
        gx = gW = None
        if ctx.needs_input_grad[0]:
            gx = Conv.apply(gy, W, kind, "dgrad" if direction == "fwd" else "fwd", scale)
        if ctx.needs_input_grad[1] and _want_params():
            gW = ConvWgrad.apply(x, gy, kind, scale, W.shape) if direction == "fwd" else \
                ConvWgrad.apply(gy, x, kind, scale, W.shape)
        return gx, gW, None, None, None


class ConvWgrad(Function):
    """dW = scale * wgrad(x_like, dy_like) for the layer's forward geometry (bilinear in its two inputs)."""

    @staticmethod
    def forward(ctx, xl, dyl, kind, scale, wshape):
        _require_cuda(xl)
        xl, dyl = _c(xl), _c(dyl)
        co, ci, k, _ = wshape
        N, H, Wd, C = xl.shape
        assert C == ci and dyl.shape[-1] == co
        g = _geom.make_geom(kind, k, "fwd", N, H, Wd, ci, co)
        assert tuple(dyl.shape[1:3]) == (g.Hout, g.Wout)
        t = _geom.tables(kind, k, "fwd")
        npt = t["nphase"] * t["ntaps"]
        dW = torch.empty(wshape, dtype=torch.float32, device=xl.device)
        from . import _tc
        if not _tc.try_wgrad(xl, dyl, kind, scale, dW):
            ns = lib.sg_wgrad_splits(ctypes.addressof(g))
            part = torch.empty((ns, npt, ci, co), dtype=torch.float32, device=xl.device)
            lib.sg_wgrad_generic(ptr(xl), ptr(dyl), ptr(part), ns, code(xl.dtype), ctypes.addressof(g), stream())
            coef = _coef_dev(kind, k, "fwd", xl.device)
            lib.sg_wprep_bwd(ptr(part), ns, ptr(dW), co, ci, k * k, npt, ptr(coef), float(scale), 0, 0, stream())
            count_launch(2)
        ctx.save_for_backward(xl, dyl)
        ctx.cfg = (kind, scale)
        return dW

    @staticmethod
    def backward(ctx, ggW):
        xl, dyl = ctx.saved_tensors
        kind, scale = ctx.cfg
        ggW = _c(ggW)
        gx = Conv.apply(dyl, ggW, kind, "dgrad", scale) if ctx.needs_input_grad[0] else None
        gd = Conv.apply(xl, ggW, kind, "fwd", scale) if ctx.needs_input_grad[1] else None
        return gx, gd, None, None, None


class ConvBiasAct(Function):
    """y = [lrelu](conv(x; scale*W) + bias*bias_mul) with bias/activation fused into the conv epilogue."""

    @staticmethod
    def forward(ctx, x, W, bias, kind, scale, bias_mul, act):
        y = _conv_raw(x, W, kind, "fwd", scale, bias, bias_mul, act)
        ctx.save_for_backward(x, W, y if act else None)
        ctx.cfg = (kind, scale, bias_mul, act, bias is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, W, y = ctx.saved_tensors
        kind, scale, bias_mul, act, has_bias = ctx.cfg
        wp = _want_params()
        want_gb = has_bias and ctx.needs_input_grad[2] and wp
        gb = None
        if act and want_gb and LreluBwdBias.supported(gy):
            g, gb = LreluBwdBias.apply(gy, y, bias_mul)             # one pass: masked gradient + its channel sums
        else:
            g = LreluBwd.apply(gy, y) if act else gy
        gx = Conv.apply(g, W, kind, "dgrad", scale) if ctx.needs_input_grad[0] else None
        gW = ConvWgrad.apply(x, g, kind, scale, W.shape) if (ctx.needs_input_grad[1] and wp) else None
        if want_gb and gb is None:
            gb = ChannelSum.apply(g, None, True, bias_mul).reshape(-1)
        return gx, gW, gb, None, None, None, None


class LreluBwd(Function):
    """gx = g * lrelu'(ref): linear in g, piecewise constant in ref."""

    @staticmethod
    def forward(ctx, g, ref):
        _require_cuda(g)
        g, ref = _c(g), _c(ref)
        assert g.dtype == ref.dtype and g.shape == ref.shape
        out = torch.empty_like(g)
        lib.sg_lrelu_bwd(ptr(g), ptr(ref), ptr(out), g.numel(), code(g.dtype), stream())
        count_launch()
        ctx.save_for_backward(ref)
        return out

    @staticmethod
    def backward(ctx, gg):
        (ref,) = ctx.saved_tensors
        return LreluBwd.apply(gg, ref), None


class LreluBwdBias(Function):
    """(gx, gb) = (g * lrelu'(ref), bias_mul * sum over pixels of gx) in one pass over NHWC tensors; linear in g, gb is a
    leaf gradient (not differentiable, like ChannelSum)."""

    @staticmethod
    def supported(g):
        C = g.shape[-1]
        return g.is_cuda and g.dim() == 4 and bool(lib.sg_lrelu_bwd_bias_supported(code(g.dtype), g.numel() // C, C))

    @staticmethod
    def forward(ctx, g, ref, bias_mul):
        _require_cuda(g)
        g, ref = _c(g), _c(ref)
        assert g.dtype == ref.dtype and g.shape == ref.shape
        C = g.shape[-1]
        P = g.numel() // C
        out = torch.empty_like(g)
        gb = torch.empty((C,), dtype=torch.float32, device=g.device)
        ws = torch.empty((lib.sg_channel_sum_blocks(P), C), dtype=torch.float32, device=g.device)
        lib.sg_lrelu_bwd_bias(ptr(g), ptr(ref), ptr(out), code(g.dtype), P, C, ptr(ws), ptr(gb), float(bias_mul), stream())
        count_launch(2)
        ctx.save_for_backward(ref)
        ctx.mark_non_differentiable(gb)
        return out, gb

    @staticmethod
    def backward(ctx, gg, _):
        (ref,) = ctx.saved_tensors
        return LreluBwd.apply(gg, ref), None, None


class ChannelSum(Function):
    """rows[j][c] = scale * sum_p wt_j(p) x[p][c] (+ an all-ones row).  x NHWC, wt fp32 [N,J,H,W] or None."""

    @staticmethod
    def forward(ctx, x, wt, ones, scale):
        _require_cuda(x)
        x = _c(x)
        N = x.shape[0]
        C = x.shape[-1]
        P = x.numel() // C
        HW = P // N
        J = 0
        if wt is not None:
            wt = _c(wt.float())
            J = wt.numel() // P
        rows = J + (1 if ones else 0)
        out = torch.empty((rows, C), dtype=torch.float32, device=x.device)
        nblk = lib.sg_channel_sum_blocks(P)
        ws = torch.empty((nblk, rows, C), dtype=torch.float32, device=x.device)
        lib.sg_channel_sum(ptr(x), code(x.dtype), ptr(wt), J, int(bool(ones)), P, HW, C, ptr(ws), ptr(out),
                           float(scale), 0, stream())
        count_launch(2)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        raise RuntimeError("sgb200: ChannelSum is a leaf gradient reduction and is not differentiable")


class Blur(Function):
    """Depthwise [1,2,1]^2/16 blur (+ optional per-channel bias)."""

    @staticmethod
    def forward(ctx, x, bias, bias_mul):
        _require_cuda(x)
        x = _c(x)
        N, H, W, C = x.shape
        y = torch.empty_like(x)
        lib.sg_blur3(ptr(x), ptr(y), ptr(bias), float(bias_mul), code(x.dtype), N, H, W, C, stream())
        count_launch()
        ctx.cfg = (bias is not None, bias_mul)
        return y

    @staticmethod
    def backward(ctx, g):
        has_bias, bias_mul = ctx.cfg
        gx = Blur.apply(g, None, 1.0) if ctx.needs_input_grad[0] else None
        gb = ChannelSum.apply(g, None, True, bias_mul).reshape(-1) \
            if (has_bias and ctx.needs_input_grad[1] and _want_params()) else None
        return gx, gb, None


class Axpby(Function):
    """a*x + b*y (y optional)."""

    @staticmethod
    def forward(ctx, x, y, a, b):
        _require_cuda(x)
        x = _c(x)
        if y is not None:
            y = _c(y)
            assert y.shape == x.shape and y.dtype == x.dtype
        out = torch.empty_like(x)
        if isinstance(a, DynScalar):
            lib.sg_axpby_dev(ptr(x), ptr(y), ptr(out), ptr(a.dev), x.numel(), code(x.dtype), stream())
            b = a.second
        else:
            lib.sg_axpby(ptr(x), ptr(y), ptr(out), float(a), float(b), x.numel(), code(x.dtype), stream())
        count_launch()
        ctx.cfg = (a, b, y is not None)
        return out

    @staticmethod
    def backward(ctx, g):
        a, b, has_y = ctx.cfg
        gx = Axpby.apply(g, None, a, 0.0) if ctx.needs_input_grad[0] else None
        gy = Axpby.apply(g, None, b, 0.0) if (has_y and ctx.needs_input_grad[1]) else None
        return gx, gy, None, None


class Cast(Function):
    @staticmethod
    def forward(ctx, x, dtype):
        _require_cuda(x)
        x = _c(x)
        ctx.src = x.dtype
        if x.dtype == dtype:
            return x.view_as(x)
        y = torch.empty(x.shape, dtype=dtype, device=x.device)
        lib.sg_cast(ptr(x), code(x.dtype), ptr(y), code(dtype), x.numel(), stream())
        count_launch()
        return y

    @staticmethod
    def backward(ctx, g):
        return Cast.apply(g, ctx.src), None


# ---------------------------------------------------------------------------------------------------------------
# from_rgb / image helpers (discriminator side: double-differentiable)
# ---------------------------------------------------------------------------------------------------------------
class RgbIn(Function):
    """y[N,H,W,C] = [lrelu](scale * W[C,K] img[N,K,H,W] + bias*bias_mul);  img fp32 NCHW."""

    @staticmethod
    def forward(ctx, img, W, bias, scale, bias_mul, act, dtype):
        _require_cuda(img)
        img = _c(img.float())
        N, K, H, Wd = img.shape
        C = W.shape[0]
        Wm = _c(W.detach().reshape(C, K))
        y = torch.empty((N, H, Wd, C), dtype=dtype, device=img.device)
        lib.sg_rgb_in(ptr(img), ptr(Wm), ptr(bias), float(scale), float(bias_mul), int(act), ptr(y), code(dtype),
                      N, H * Wd, K, C, stream())
        count_launch()
        ctx.save_for_backward(img, W, y if act else None)
        ctx.cfg = (scale, bias_mul, act, bias is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        img, W, y = ctx.saved_tensors
        scale, bias_mul, act, has_bias = ctx.cfg
        g = LreluBwd.apply(gy, y) if act else gy
        gimg = RgbInDgrad.apply(g, W, scale) if ctx.needs_input_grad[0] else None
        gW = gb = None
        if (ctx.needs_input_grad[1] or (has_bias and ctx.needs_input_grad[2])) and _want_params():
            rows = ChannelSum.apply(g, img, True, 1.0)           # [K+1, C]
            K = img.shape[1]
            if ctx.needs_input_grad[1]:
                gW = (rows[:K].t() * scale).reshape(W.shape)
            if has_bias and ctx.needs_input_grad[2]:
                gb = rows[K] * bias_mul
        return gimg, gW, gb, None, None, None, None


class RgbInDgrad(Function):
    """gimg[N,K,H,W] = scale * W^T g  (adjoint of RgbIn wrt the image)."""

    @staticmethod
    def forward(ctx, g, W, scale):
        _require_cuda(g)
        g = _c(g)
        N, H, Wd, C = g.shape
        K = W.shape[1]
        Wm = _c(W.detach().reshape(C, K))
        gimg = torch.empty((N, K, H, Wd), dtype=torch.float32, device=g.device)
        if K == 3 and g.dtype == torch.bfloat16 and Wd % 16 == 0:
            # the adjoint of from_rgb IS a to_rgb with the transposed weight: the K = 3 kernel (csrc/sg_rgb_k3.cu) runs it
            # at 2.5 TB/s against 1.8 for the general dgrad kernel
            lib.sg_rgb_out(ptr(g), code(g.dtype), ptr(_c(Wm.t())), None, float(scale), 1.0, 1.0, 0.0, ptr(gimg), N, H, Wd,
                           C, K, 1, stream())
        else:
            lib.sg_rgb_in_dgrad(ptr(g), code(g.dtype), ptr(Wm), float(scale), ptr(gimg), N, H * Wd, K, C, stream())
        count_launch()
        ctx.save_for_backward(g, W)
        ctx.scale = scale
        return gimg

    @staticmethod
    def backward(ctx, t):
        g, W = ctx.saved_tensors
        scale = ctx.scale
        cg = RgbIn.apply(t, W, None, scale, 1.0, False, g.dtype) if ctx.needs_input_grad[0] else None
        cW = None
        if ctx.needs_input_grad[1]:
            rows = ChannelSum.apply(g, t, False, scale)           # [K, C]
            cW = rows.t().reshape(W.shape)
        return cg, cW, None


class PoolNCHW(Function):
    """fp32 NCHW box filter: out = scale * sum over fxf blocks (scale=1/f^2 -> AvgPool2d(f))."""

    @staticmethod
    def forward(ctx, x, f, scale):
        _require_cuda(x)
        x = _c(x.float())
        N, K, H, W = x.shape
        out = torch.empty((N, K, H // f, W // f), dtype=torch.float32, device=x.device)
        lib.sg_pool_nchw(ptr(x), ptr(out), N * K, H // f, W // f, f, float(scale), stream())
        count_launch()
        ctx.cfg = (f, scale)
        return out

    @staticmethod
    def backward(ctx, g):
        f, scale = ctx.cfg
        return UpsampleNCHW.apply(g, f, scale), None, None


class UpsampleNCHW(Function):
    """fp32 NCHW nearest up-sampling by f, times scale."""

    @staticmethod
    def forward(ctx, x, f, scale):
        _require_cuda(x)
        x = _c(x.float())
        N, K, H, W = x.shape
        out = torch.empty((N, K, H * f, W * f), dtype=torch.float32, device=x.device)
        lib.sg_upsample_nchw(ptr(x), ptr(out), N * K, H, W, f, float(scale), stream())
        count_launch()
        ctx.cfg = (f, scale)
        return out

    @staticmethod
    def backward(ctx, g):
        f, scale = ctx.cfg
        return PoolNCHW.apply(g, f, scale), None, None


# ---------------------------------------------------------------------------------------------------------------
# dense layers (fp32)
# ---------------------------------------------------------------------------------------------------------------
def _rows(x):
    """2-D fp32 view with unit inner stride -> (tensor, ldx)."""
    assert x.dim() == 2 and x.dtype == torch.float32
    if x.stride(1) != 1:
        x = x.contiguous()
    return x, (x.stride(0) if x.shape[0] > 1 else x.shape[1])


class Linear(Function):
    """y = [lrelu](scale * x W^T + bias*bias_mul), fp32."""

    @staticmethod
    def forward(ctx, x, W, bias, scale, bias_mul, act):
        _require_cuda(x)
        x, ldx = _rows(x)
        B, I = x.shape
        O = W.shape[0]
        y = torch.empty((B, O), dtype=torch.float32, device=x.device)
        nws = lib.sg_linear_fwd_ws_floats(B, I, O)
        ws = torch.empty((nws,), dtype=torch.float32, device=x.device) if nws else None
        lib.sg_linear_fwd(ptr(x), ldx, ptr(_c(W.detach())), ptr(bias), float(scale), float(bias_mul), int(act),
                          ptr(y), B, I, O, ptr(ws), stream())
        count_launch()
        ctx.save_for_backward(x, W, y if act else None)
        ctx.cfg = (scale, bias_mul, act, bias is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, W, y = ctx.saved_tensors
        scale, bias_mul, act, has_bias = ctx.cfg
        g = LreluBwd.apply(gy, y) if act else gy
        gx = LinearDgrad.apply(g, W, scale) if ctx.needs_input_grad[0] else None
        gW = gb = None
        if (ctx.needs_input_grad[1] or (has_bias and ctx.needs_input_grad[2])) and _want_params():
            gW, gb = LinearWgrad.apply(g, x, scale, bias_mul, has_bias)
        return gx, gW, (gb if has_bias else None), None, None, None


class LinearDgrad(Function):
    @staticmethod
    def forward(ctx, gy, W, scale):
        _require_cuda(gy)
        gy = _c(gy)
        B, O = gy.shape
        I = W.shape[1]
        gx = torch.empty((B, I), dtype=torch.float32, device=gy.device)
        lib.sg_linear_dgrad(ptr(gy), ptr(_c(W.detach())), float(scale), ptr(gx), B, I, O, stream())
        count_launch()
        ctx.save_for_backward(gy, W)
        ctx.scale = scale
        return gx

    @staticmethod
    def backward(ctx, t):
        gy, W = ctx.saved_tensors
        scale = ctx.scale
        cg = Linear.apply(t, W, None, scale, 1.0, False) if ctx.needs_input_grad[0] else None
        cW = LinearWgrad.apply(gy, t, scale, 1.0, False)[0] if ctx.needs_input_grad[1] else None
        return cg, cW, None


class LinearWgrad(Function):
    @staticmethod
    def forward(ctx, gy, x, scale, bias_mul, want_bias):
        _require_cuda(gy)
        gy = _c(gy)
        x, ldx = _rows(x)
        B, O = gy.shape
        I = x.shape[1]
        gW = torch.empty((O, I), dtype=torch.float32, device=gy.device)
        gb = torch.empty((O,), dtype=torch.float32, device=gy.device) if want_bias else None
        lib.sg_linear_wgrad(ptr(gy), ptr(x), ldx, float(scale), float(bias_mul), ptr(gW), ptr(gb), B, I, O, 0,
                            stream())
        count_launch(2 if want_bias else 1)
        ctx.mark_non_differentiable(gW)
        if gb is None:
            gb = torch.empty((0,), dtype=torch.float32, device=gy.device)
        ctx.mark_non_differentiable(gb)
        return gW, gb

    @staticmethod
    @once_differentiable
    def backward(ctx, a, b):
        raise RuntimeError("sgb200: LinearWgrad is not differentiable")


def _pack_scales(scale, bias_mul):
    import struct
    return struct.unpack("<q", struct.pack("<ff", float(scale), float(bias_mul)))[0]


def _lingroup_table(rows, device):
    """rows: list of 8 python ints per group -> int64 device tensor (one small host-to-device copy)."""
    return upload_table(torch.tensor(rows, dtype=torch.int64).reshape(-1, 8), device)


class StyleAffines(Function):
    """All StyleMod affines of one synthesis pass in one launch (csrc/sg_linear_group.cu; reference
    CustomLayers.py:203-216 evaluates them one by one inside each LayerEpilogue):
        styles[l] = w_mul_l * dlatents[:, idx_l] W_l^T + b_l * b_mul_l          (fp32, [B, 2C_l])
    backward: one launch for the gradient wrt dlatents and one for all weight + bias gradients.
    args: dlatents [B, L, D] fp32 contiguous, idx (tuple of layer indices), cfg (tuple of (w_mul, b_mul)), then
    W_0, b_0, W_1, b_1, ... (b may be None)."""

    @staticmethod
    def supported(dlatents):
        B, L, D = dlatents.shape
        return dlatents.dtype == torch.float32 and bool(lib.sg_lingroup_supported(min(B, 32), D))

    @staticmethod
    def forward(ctx, dlatents, idx, cfg, *params):
        _require_cuda(dlatents)
        dl = _c(dlatents)
        B, L, D = dl.shape
        dev = dl.device
        Ws, bs = params[0::2], params[1::2]
        outs = [torch.empty((B, W.shape[0]), dtype=torch.float32, device=dev) for W in Ws]
        for b0 in range(0, B, 32):
            nb = min(32, B - b0)
            rows, blk = [], 0
            for l, W, b, (w_mul, b_mul), y in zip(idx, Ws, bs, cfg, outs):
                O = W.shape[0]
                rows.append([dl.data_ptr() + 4 * (b0 * L * D + l * D), L * D, W.data_ptr(),
                             b.data_ptr() if b is not None else 0, y.data_ptr() + 4 * b0 * O, O,
                             _pack_scales(w_mul, b_mul), blk])
                blk += lib.sg_lingroup_fwd_blocks(O)
            tab = _lingroup_table(rows, dev)
            lib.sg_lingroup_fwd(ptr(tab), len(rows), blk, nb, D, 0, stream())
            count_launch()
        ctx.save_for_backward(dl, *[p for p in params if p is not None])
        ctx.meta = (idx, cfg, [b is not None for b in bs])
        ctx.mark_non_differentiable()
        return tuple(outs)

    @staticmethod
    @once_differentiable
    def backward(ctx, *gys):
        idx, cfg, has_b = ctx.meta
        saved = ctx.saved_tensors
        dl = saved[0]
        it = iter(saved[1:])
        Ws, bs = [], []
        for hb in has_b:
            Ws.append(next(it))
            bs.append(next(it) if hb else None)
        B, L, D = dl.shape
        dev = dl.device
        gys = [(_c(g) if g is not None else torch.zeros((B, W.shape[0]), dtype=torch.float32, device=dev))
               for g, W in zip(gys, Ws)]
        gdl = None
        if ctx.needs_input_grad[0]:
            assert len(set(idx)) == len(idx), "each dlatent layer feeds one StyleMod"
            gdl = torch.empty_like(dl) if len(idx) == L else torch.zeros_like(dl)
            for b0 in range(0, B, 32):
                nb = min(32, B - b0)
                rows = []
                for l, W, (w_mul, _), g in zip(idx, Ws, cfg, gys):
                    O = W.shape[0]
                    rows.append([gdl.data_ptr() + 4 * (b0 * L * D + l * D), L * D, W.data_ptr(), 0,
                                 g.data_ptr() + 4 * b0 * O, O, _pack_scales(w_mul, 1.0), 0])
                tab = _lingroup_table(rows, dev)
                lib.sg_lingroup_dgrad(ptr(tab), len(rows), nb, D, 0, stream())
                count_launch()
        grads = [None] * (2 * len(Ws))
        if _want_params() and any(ctx.needs_input_grad[3:]):
            assert B <= 32, "sgb200: grouped StyleMod weight gradients are implemented for batch <= 32 per GPU"
            rows, blk = [], 0
            for k, (l, W, b, (w_mul, b_mul), g) in enumerate(zip(idx, Ws, bs, cfg, gys)):
                O = W.shape[0]
                gW = torch.empty_like(W)
                gb = torch.empty_like(b) if b is not None else None
                grads[2 * k], grads[2 * k + 1] = gW, gb
                rows.append([dl.data_ptr() + 4 * l * D, L * D, gW.data_ptr(), gb.data_ptr() if gb is not None else 0,
                             g.data_ptr(), O, _pack_scales(w_mul, b_mul), blk])
                blk += lib.sg_lingroup_wgrad_blocks(O)
            tab = _lingroup_table(rows, dev)
            lib.sg_lingroup_wgrad(ptr(tab), len(rows), blk, B, D, stream())
            count_launch()
        return (gdl, None, None, *grads)


class PixelNorm(Function):
    @staticmethod
    def forward(ctx, x, eps):
        _require_cuda(x)
        x = _c(x.float())
        y = torch.empty_like(x)
        lib.sg_pixelnorm_fwd(ptr(x), ptr(y), x.shape[0], x.shape[1], float(eps), stream())
        count_launch()
        ctx.save_for_backward(x)
        ctx.eps = eps
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        gx = torch.empty_like(x)
        lib.sg_pixelnorm_bwd(ptr(_c(g)), ptr(x), ptr(gx), x.shape[0], x.shape[1], float(ctx.eps), stream())
        count_launch()
        return gx, None


class MbStd(Function):
    """MinibatchStdDev on fp32 NHWC [B,H,W,C] -> [B,H,W,C+1]."""

    @staticmethod
    def forward(ctx, x, group):
        _require_cuda(x)
        x = _c(x)
        assert x.dtype == torch.float32
        B, H, W, C = x.shape
        G = min(group, B)
        out = torch.empty((B, H, W, C + 1), dtype=torch.float32, device=x.device)
        ws = torch.empty((2 * (B // G) + 2,), dtype=torch.float32, device=x.device)
        lib.sg_mbstd_fwd(ptr(x), ptr(out), ptr(ws), B, H * W, C, G, stream())
        count_launch(2)
        ctx.save_for_backward(x)
        ctx.G = G
        return out

    @staticmethod
    def backward(ctx, gout):
        (x,) = ctx.saved_tensors
        return MbStdBwd.apply(gout, x, ctx.G), None


class MbStdBwd(Function):
    @staticmethod
    def forward(ctx, gout, x, G):
        gout = _c(gout)
        B, H, W, C = x.shape
        gx = torch.empty_like(x)
        ws = torch.empty((2 * (B // G) + 2,), dtype=torch.float32, device=x.device)
        lib.sg_mbstd_bwd(ptr(gout), ptr(x), ptr(gx), ptr(ws), B, H * W, C, G, stream())
        count_launch(2)
        ctx.save_for_backward(gout, x)
        ctx.G = G
        return gx

    @staticmethod
    @once_differentiable
    def backward(ctx, t):
        gout, x = ctx.saved_tensors
        t = _c(t)
        B, H, W, C = x.shape
        G = ctx.G
        cg = torch.empty_like(gout)
        cx = torch.empty_like(x)
        ws = torch.empty((2 * (B // G) + 2,), dtype=torch.float32, device=x.device)
        lib.sg_mbstd_bwd2(ptr(t), ptr(gout), ptr(x), ptr(cg), ptr(cx), ptr(ws), B, H * W, C, G, stream())
        count_launch(3)
        return cg, cx, None


# ---------------------------------------------------------------------------------------------------------------
# losses
# ---------------------------------------------------------------------------------------------------------------
class SoftplusMean(Function):
    """mean(softplus(sign * x)) as a 1-element fp32 tensor."""

    @staticmethod
    def forward(ctx, x, sign):
        _require_cuda(x)
        x = _c(x.float())
        n = x.numel()
        out = torch.empty((1,), dtype=torch.float32, device=x.device)
        lib.sg_softplus_sum(ptr(x), n, float(sign), 1.0 / n, ptr(out), 0, stream())
        count_launch()
        ctx.save_for_backward(x)
        ctx.sign = sign
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        n = x.numel()
        gx = torch.empty_like(x)
        lib.sg_softplus_bwd(ptr(x), n, float(ctx.sign), ptr(_c(g.float())), 1.0 / n, ptr(gx), stream())
        count_launch()
        return gx, None


class SqSum(Function):
    """scale * sum(x^2) as a 1-element fp32 tensor (R1 penalty, Losses.py:209)."""

    @staticmethod
    def forward(ctx, x, scale):
        _require_cuda(x)
        x = _c(x.float())
        out = torch.empty((1,), dtype=torch.float32, device=x.device)
        ws = torch.empty((1024,), dtype=torch.float64, device=x.device)
        lib.sg_sqsum(ptr(x), x.numel(), float(scale), ptr(out), 0, ptr(ws), stream())
        count_launch(2)
        ctx.save_for_backward(x)
        ctx.scale = scale
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        gx = torch.empty_like(x)
        lib.sg_scale_by_scalar(ptr(x), ptr(_c(g.float())), 2.0 * ctx.scale, ptr(gx), x.numel(), stream())
        count_launch()
        return gx, None


# ---------------------------------------------------------------------------------------------------------------
# generator-only fused Functions (once-differentiable)
# ---------------------------------------------------------------------------------------------------------------
class Epilogue(Function):
    """LayerEpilogue: out = StyleMod(InstanceNorm(lrelu(y + bias*bmul + noise_w*noise)), style).
    y NHWC; bias [C] or None; noise fp32 [N,1,H,W] or None; noise_w [C]; style fp32 [N,2C] or None."""

    @staticmethod
    def forward(ctx, y, bias, bias_mul, noise, noise_w, style, eps, act=True):
        _require_cuda(y)
        y = _c(y)
        N, H, W, C = y.shape
        HW = H * W
        dev = y.device
        if noise is not None:
            noise = _c(noise.float())
            assert noise.numel() == N * HW, "noise must be [N,1,H,W]"
        if style is not None:
            style = _c(style)
        z = torch.empty_like(y)
        sums = torch.empty((N, C, 3), dtype=torch.float64, device=dev)
        ws = torch.empty((lib.sg_epi_ws_floats(N, HW, C),), dtype=torch.float32, device=dev)
        lib.sg_epi_act_stats(ptr(y), ptr(bias), float(bias_mul), ptr(noise), ptr(noise_w if noise is not None else None),
                             int(act), ptr(z), ptr(sums), ptr(ws), code(y.dtype), N, HW, C, stream())
        mr = torch.empty((N, C, 2), dtype=torch.float32, device=dev)
        ab = torch.empty((N, C, 3), dtype=torch.float32, device=dev)
        lib.sg_epi_finalize(ptr(sums), ptr(style), ptr(mr), ptr(ab), N, C, HW, float(eps), stream())
        out = torch.empty_like(y)
        lib.sg_affine_nc(ptr(z), ptr(ab), ptr(out), code(y.dtype), N, HW, C, stream())
        count_launch(4)
        ctx.save_for_backward(z, mr, style, noise)
        ctx.cfg = (bias is not None, bias_mul, noise is not None, act)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        z, mr, style, noise = ctx.saved_tensors
        has_bias, bias_mul, has_noise, act = ctx.cfg
        gy, gb, gnw, gstyle = _epilogue_backward(g, z, mr, style, noise, act, bias_mul,
                                                 has_bias and ctx.needs_input_grad[1],
                                                 has_noise and ctx.needs_input_grad[4])
        return gy, gb, None, None, gnw, gstyle, None, None


def _epilogue_backward(g, z, mr, style, noise, act, bias_mul, need_b, need_n):
    """Backward of StyleMod(InstanceNorm(z)) o lrelu o (+bias + noise_w*noise): g = grad wrt the epilogue output ->
    (grad wrt the pre-activation conv output, grad bias, grad noise weight, grad style)."""
    g = _c(g)
    N, H, W, C = z.shape
    HW = H * W
    dev = z.device
    sums = torch.empty((N, C, 3), dtype=torch.float64, device=dev)
    ws = torch.empty((lib.sg_epi_ws_floats(N, HW, C),), dtype=torch.float32, device=dev)
    lib.sg_epi_bwd_reduce(ptr(g), ptr(z), ptr(mr), ptr(sums), ptr(ws), code(z.dtype), N, HW, C, stream())
    gstyle = torch.empty((N, 2 * C), dtype=torch.float32, device=dev) if style is not None else None
    coef = torch.empty((N, C, 4), dtype=torch.float32, device=dev)
    lib.sg_epi_bwd_finalize(ptr(sums), ptr(mr), ptr(style), ptr(gstyle), ptr(coef), N, C, HW, stream())
    gy = torch.empty_like(z)
    lib.sg_epi_bwd_apply(ptr(g), ptr(z), ptr(coef), int(act), ptr(gy), code(z.dtype), N, HW, C, stream())
    count_launch(4)
    gb = gnw = None
    if need_b or need_n:
        rows = ChannelSum.apply(gy, noise if need_n else None, need_b, 1.0)
        if need_n:
            gnw = rows[0]
        if need_b:
            gb = rows[-1] * bias_mul
    return gy, gb, gnw, gstyle


class ConvEpilogue(Function):
    """conv3x3 -> LayerEpilogue with the noise add, LeakyReLU and the InstanceNorm statistics fused into the tcgen05
    convolution's epilogue (reference CustomLayers.py:170-171 then :242-248): the convolution result is written once,
    already activated, and the only other forward pass over it is the InstanceNorm + StyleMod affine.
    x NHWC bf16; W [Co,Ci,3,3]; bias [Co]; noise fp32 [N,1,H,W]; noise_w [Co]; style fp32 [N,2Co]."""

    @staticmethod
    def forward(ctx, x, W, bias, scale, bias_mul, noise, noise_w, style, eps):
        from . import _tc
        _require_cuda(x)
        x = _c(x)
        if noise is not None:
            noise = _c(noise.float())
            assert noise.numel() == x.shape[0] * x.shape[1] * x.shape[2], "noise must be [N,1,H,W]"
        if style is not None:
            style = _c(style)
        z, mr, ab = _tc.conv_epilogue_forward(x, W, scale, bias, bias_mul, noise, noise_w, style, eps)
        N, H, Wd, C = z.shape
        out = torch.empty_like(z)
        lib.sg_affine_nc(ptr(z), ptr(ab), ptr(out), code(z.dtype), N, H * Wd, C, stream())
        count_launch()
        ctx.save_for_backward(x, W, z, mr, style, noise)
        ctx.cfg = (scale, bias is not None, bias_mul, noise is not None)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, W, z, mr, style, noise = ctx.saved_tensors
        scale, has_bias, bias_mul, has_noise = ctx.cfg
        gy, gb, gnw, gstyle = _epilogue_backward(g, z, mr, style, noise, True, bias_mul,
                                                 has_bias and ctx.needs_input_grad[2],
                                                 has_noise and ctx.needs_input_grad[6])
        with torch.no_grad():
            gx = _conv_raw(gy, W, "same", "dgrad", scale) if ctx.needs_input_grad[0] else None
            gW = ConvWgrad.apply(x, gy, "same", scale, W.shape) if ctx.needs_input_grad[1] else None
        return gx, gW, gb, None, None, None, gnw, gstyle, None


class BlurEpilogue(Function):
    """BlurLayer (+ the conv bias) -> LayerEpilogue with the blur, bias, noise add, LeakyReLU and the InstanceNorm
    statistics in one pass (csrc/sg_blur_v2.cu: sg_blur3_epi; reference CustomLayers.py:174-180 then :242-248).
    y NHWC bf16 = the up-convolution result WITHOUT bias."""

    @staticmethod
    def supported(y):
        if y.dtype != torch.bfloat16 or not _state.get("fuse_epilogue", True):
            return False
        N, H, W, C = y.shape
        return lib.sg_blur3_epi_slots(N, H, W, C) > 0

    @staticmethod
    def forward(ctx, y, bias, bias_mul, noise, noise_w, style, eps):
        _require_cuda(y)
        y = _c(y)
        N, H, W, C = y.shape
        dev = y.device
        if noise is not None:
            noise = _c(noise.float())
            assert noise.numel() == N * H * W, "noise must be [N,1,H,W]"
        if style is not None:
            style = _c(style)
        slots = lib.sg_blur3_epi_slots(N, H, W, C)
        z = torch.empty_like(y)
        part = torch.empty((N, slots, 2, C), dtype=torch.float32, device=dev)
        lib.sg_blur3_epi(ptr(y), ptr(z), ptr(bias), float(bias_mul), ptr(noise),
                         ptr(noise_w if noise is not None else None), ptr(part), N, H, W, C, stream())
        mr = torch.empty((N, C, 2), dtype=torch.float32, device=dev)
        ab = torch.empty((N, C, 3), dtype=torch.float32, device=dev)
        lib.sg_epi_finalize_slots(ptr(part), slots, ptr(style), ptr(mr), ptr(ab), N, C, H * W, float(eps), stream())
        out = torch.empty_like(y)
        lib.sg_affine_nc(ptr(z), ptr(ab), ptr(out), code(y.dtype), N, H * W, C, stream())
        count_launch(3)
        ctx.save_for_backward(z, mr, style, noise)
        ctx.cfg = (bias is not None, bias_mul, noise is not None)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        z, mr, style, noise = ctx.saved_tensors
        has_bias, bias_mul, has_noise = ctx.cfg
        gp, gb, gnw, gstyle = _epilogue_backward(g, z, mr, style, noise, True, bias_mul,
                                                 has_bias and ctx.needs_input_grad[1],
                                                 has_noise and ctx.needs_input_grad[4])
        gy = None
        if ctx.needs_input_grad[0]:
            N, H, W, C = z.shape
            gy = torch.empty_like(gp)                       # the blur is self-adjoint
            lib.sg_blur3(ptr(gp), ptr(gy), None, 1.0, code(gp.dtype), N, H, W, C, stream())
            count_launch()
        return gy, gb, None, None, gnw, gstyle, None


class ToRgb(Function):
    """Fused to_rgb + fade-in (reference GAN.py:199-204):
        out[N,K,H,W] fp32 = a_hi * rgb(x_hi; W_hi, b_hi) + a_lo * rgb(nearest_up2(x_lo); W_lo, b_lo)
    x_lo/W_lo/b_lo may be None (depth 0 / 'fixed')."""

    @staticmethod
    def forward(ctx, x_hi, W_hi, b_hi, s_hi, a_hi, x_lo, W_lo, b_lo, s_lo, a_lo):
        _require_cuda(x_hi)
        x_hi = _c(x_hi)
        N, H, W, C = x_hi.shape
        K = W_hi.shape[0]
        out = torch.empty((N, K, H, W), dtype=torch.float32, device=x_hi.device)
        Wh = _c(W_hi.detach().reshape(K, C))
        lib.sg_rgb_out(ptr(x_hi), code(x_hi.dtype), ptr(Wh), ptr(b_hi), float(s_hi), 1.0, float(a_hi), 0.0, ptr(out),
                       N, H, W, C, K, 1, stream())
        count_launch()
        if x_lo is not None:
            x_lo = _c(x_lo)
            Cl = x_lo.shape[-1]
            Wl = _c(W_lo.detach().reshape(K, Cl))
            lib.sg_rgb_out(ptr(x_lo), code(x_lo.dtype), ptr(Wl), ptr(b_lo), float(s_lo), 1.0, float(a_lo), 1.0,
                           ptr(out), N, H // 2, W // 2, Cl, K, 2, stream())
            count_launch()
        ctx.save_for_backward(x_hi, W_hi, x_lo, W_lo)
        ctx.cfg = (s_hi, a_hi, s_lo, a_lo, b_hi is not None, b_lo is not None)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x_hi, W_hi, x_lo, W_lo = ctx.saved_tensors
        s_hi, a_hi, s_lo, a_lo, hb_hi, hb_lo = ctx.cfg
        g = _c(g.float())
        N, H, W, C = x_hi.shape
        K = W_hi.shape[0]
        dev = g.device
        st = stream()
        res = [None] * 10
        Wh = _c(W_hi.detach().reshape(K, C))
        if ctx.needs_input_grad[0]:
            gx = torch.empty_like(x_hi)
            lib.sg_rgb_out_dgrad(ptr(g), ptr(Wh), float(a_hi * s_hi), ptr(gx), code(x_hi.dtype), N, H, W, C, K, 1, st)
            count_launch()
            res[0] = gx
        if ctx.needs_input_grad[1]:
            res[1] = ChannelSum.apply(x_hi, g, False, a_hi * s_hi).reshape(W_hi.shape)
        if hb_hi and ctx.needs_input_grad[2]:
            gb = torch.empty((K,), dtype=torch.float32, device=dev)
            ws = torch.empty((K * 64,), dtype=torch.float32, device=dev)
            lib.sg_plane_sum(ptr(g), ptr(gb), ptr(ws), N, K, H * W, float(a_hi), 0, st)
            count_launch(2)
            res[2] = gb
        if x_lo is not None:
            Cl = x_lo.shape[-1]
            Wl = _c(W_lo.detach().reshape(K, Cl))
            if ctx.needs_input_grad[5]:
                gx = torch.empty_like(x_lo)
                lib.sg_rgb_out_dgrad(ptr(g), ptr(Wl), float(a_lo * s_lo), ptr(gx), code(x_lo.dtype), N, H // 2, W // 2,
                                     Cl, K, 2, st)
                count_launch()
                res[5] = gx
            if ctx.needs_input_grad[6]:
                glo = torch.empty((N, K, H // 2, W // 2), dtype=torch.float32, device=dev)
                lib.sg_pool_nchw(ptr(g), ptr(glo), N * K, H // 2, W // 2, 2, 1.0, st)
                count_launch()
                res[6] = ChannelSum.apply(x_lo, glo, False, a_lo * s_lo).reshape(W_lo.shape)
            if hb_lo and ctx.needs_input_grad[7]:
                gb = torch.empty((K,), dtype=torch.float32, device=dev)
                ws = torch.empty((K * 64,), dtype=torch.float32, device=dev)
                lib.sg_plane_sum(ptr(g), ptr(gb), ptr(ws), N, K, H * W, float(a_lo), 0, st)
                count_launch(2)
                res[7] = gb
        return tuple(res)


class Dlatents(Function):
    """[B,L,D] dlatents from w1 (and w2 for style mixing) with truncation towards avg."""

    @staticmethod
    def forward(ctx, w1, w2, avg, L, mix_cutoff, psi, trunc_cutoff):
        _require_cuda(w1)
        w1 = _c(w1)
        B, D = w1.shape
        out = torch.empty((B, L, D), dtype=torch.float32, device=w1.device)
        if isinstance(mix_cutoff, DynInt):
            lib.sg_dlatents_dev(ptr(w1), ptr(_c(w2) if w2 is not None else None), ptr(avg), ptr(out), B, L, D,
                                ptr(mix_cutoff.dev), float(psi), int(trunc_cutoff), stream())
        else:
            lib.sg_dlatents(ptr(w1), ptr(_c(w2) if w2 is not None else None), ptr(avg), ptr(out), B, L, D,
                            int(mix_cutoff), float(psi), int(trunc_cutoff), stream())
        count_launch()
        ctx.cfg = (L, mix_cutoff, psi, trunc_cutoff, w2 is not None, avg is not None)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        L, mix_cutoff, psi, trunc_cutoff, has2, has_trunc = ctx.cfg
        g = _c(g)
        B, _, D = g.shape
        g1 = torch.empty((B, D), dtype=torch.float32, device=g.device)
        g2 = torch.empty((B, D), dtype=torch.float32, device=g.device) if has2 else None
        if isinstance(mix_cutoff, DynInt):
            lib.sg_dlatents_bwd_dev(ptr(g), ptr(g1), ptr(g2), B, L, D, ptr(mix_cutoff.dev), float(psi),
                                    int(trunc_cutoff), int(has_trunc), stream())
        else:
            lib.sg_dlatents_bwd(ptr(g), ptr(g1), ptr(g2), B, L, D, int(mix_cutoff), float(psi), int(trunc_cutoff),
                                int(has_trunc), stream())
        count_launch()
        return g1, g2, None, None, None, None, None


def randn_device(shape, device, seed, offset):
    """N(0,1) fp32 noise from the engine's Philox generator (NoiseLayer, CustomLayers.py:193)."""
    out = torch.empty(shape, dtype=torch.float32, device=device)
    c = _capture[0]
    if c is not None:
        # inside a CUDA-graph capture: counter = (device) base of this step + the offset inside the step
        lib.sg_randn_dev(ptr(out), out.numel(), int(seed) & (2 ** 64 - 1), ptr(c.rng_base), int(offset) - c.rng_offset0,
                         stream())
    else:
        lib.sg_randn(ptr(out), out.numel(), int(seed) & (2 ** 64 - 1), int(offset), stream())
    count_launch()
    return out


def he_scale(fan_in, gain, lrmul=1.0):
    return gain * fan_in ** (-0.5) * lrmul


class BiasNoise(Function):
    """y = x + bias*bias_mul + noise_w*noise (stand-alone NoiseLayer / bias add)."""

    @staticmethod
    def forward(ctx, x, bias, bias_mul, noise, noise_w, act=False):
        _require_cuda(x)
        x = _c(x)
        C = x.shape[-1]
        P = x.numel() // C
        if noise is not None:
            noise = _c(noise.float())
            assert noise.numel() == P
        y = torch.empty_like(x)
        lib.sg_bias_noise(ptr(x), ptr(bias), float(bias_mul), ptr(noise), ptr(noise_w if noise is not None else None),
                          int(act), ptr(y), code(x.dtype), P, C, stream())
        count_launch()
        ctx.save_for_backward(noise, y if act else None)
        ctx.cfg = (bias is not None, bias_mul, noise is not None)
        return y

    @staticmethod
    def backward(ctx, g):
        noise, y = ctx.saved_tensors
        has_bias, bias_mul, has_noise = ctx.cfg
        if y is not None:
            g = LreluBwd.apply(g, y)
        gb = gnw = None
        need_b = has_bias and ctx.needs_input_grad[1]
        need_n = has_noise and ctx.needs_input_grad[4]
        if need_b or need_n:
            rows = ChannelSum.apply(g, noise if need_n else None, need_b, 1.0)
            if need_n:
                gnw = rows[0]
            if need_b:
                gb = rows[-1] * bias_mul
        return g, gb, None, None, gnw, None


class AffineNC(Function):
    """out[n,p,c] = ab[n,c,0]*x + ab[n,c,1] (stand-alone StyleMod / InstanceNorm application)."""

    @staticmethod
    def forward(ctx, x, ab):
        _require_cuda(x)
        x = _c(x)
        ab = _c(ab.float())
        N, C = x.shape[0], x.shape[-1]
        HW = x.numel() // (N * C)
        out = torch.empty_like(x)
        mab = torch.cat([torch.zeros_like(ab[..., :1]), ab], dim=-1).contiguous()      # (mean=0, a, b)
        lib.sg_affine_nc(ptr(x), ptr(mab), ptr(out), code(x.dtype), N, HW, C, stream())
        count_launch()
        ctx.save_for_backward(x, ab)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, ab = ctx.saved_tensors
        g = _c(g)
        N, C = x.shape[0], x.shape[-1]
        HW = x.numel() // (N * C)
        dev = x.device
        gx = gab = None
        if ctx.needs_input_grad[0]:
            z0 = torch.zeros_like(ab[..., 0])
            a0 = torch.stack([z0, ab[..., 0], z0], dim=-1).contiguous()
            gx = torch.empty_like(x)
            lib.sg_affine_nc(ptr(g), ptr(a0), ptr(gx), code(x.dtype), N, HW, C, stream())
            count_launch()
        if ctx.needs_input_grad[1]:
            sums = torch.empty((N, C, 3), dtype=torch.float64, device=dev)
            ws = torch.empty((lib.sg_epi_ws_floats(N, HW, C),), dtype=torch.float32, device=dev)
            mr = torch.zeros((N, C, 2), dtype=torch.float32, device=dev)
            mr[..., 1] = 1.0
            lib.sg_epi_bwd_reduce(ptr(g), ptr(x), ptr(mr), ptr(sums), ptr(ws), code(x.dtype), N, HW, C, stream())
            count_launch()
            gab = torch.stack([sums[..., 1], sums[..., 0]], dim=-1).float()   # d/da = sum g*x, d/db = sum g
        return gx, gab


class UpNHWC(Function):
    """Nearest x2 up-sampling of an NHWC feature map, times scale (adjoint: PoolNHWC)."""

    @staticmethod
    def forward(ctx, x, scale):
        _require_cuda(x)
        x = _c(x)
        N, H, W, C = x.shape
        y = torch.empty((N, 2 * H, 2 * W, C), dtype=x.dtype, device=x.device)
        lib.sg_up2_nhwc(ptr(x), ptr(y), float(scale), code(x.dtype), N, H, W, C, stream())
        count_launch()
        ctx.scale = scale
        return y

    @staticmethod
    def backward(ctx, g):
        return PoolNHWC.apply(g, None, 1.0, False, ctx.scale), None


class PoolNHWC(Function):
    """y = [lrelu](scale * (2x2 box sum of x) + bias*bias_mul) on NHWC feature maps (adjoint: UpNHWC)."""

    @staticmethod
    def forward(ctx, x, bias, bias_mul, act, scale):
        _require_cuda(x)
        x = _c(x)
        N, H, W, C = x.shape
        y = torch.empty((N, H // 2, W // 2, C), dtype=x.dtype, device=x.device)
        lib.sg_pool2_nhwc(ptr(x), ptr(y), ptr(bias), float(bias_mul), int(act), float(scale), code(x.dtype), N, H // 2,
                          W // 2, C, stream())
        count_launch()
        ctx.save_for_backward(y if act else None)
        ctx.cfg = (bias is not None, bias_mul, act, scale)
        return y

    @staticmethod
    def backward(ctx, gy):
        (y,) = ctx.saved_tensors
        has_bias, bias_mul, act, scale = ctx.cfg
        g = LreluBwd.apply(gy, y) if act else gy
        gx = UpNHWC.apply(g, scale) if ctx.needs_input_grad[0] else None
        gb = None
        if has_bias and ctx.needs_input_grad[1] and _want_params():
            gb = ChannelSum.apply(g, None, True, bias_mul).reshape(-1)
        return gx, gb, None, None, None


class PadCrop(Function):
    """Zero-pad / crop an NHWC map to (H, W) (top-left anchored).  Linear and self-adjoint up to the shape swap."""

    @staticmethod
    def forward(ctx, x, H, W):
        _require_cuda(x)
        x = _c(x)
        N, h, w, C = x.shape
        y = torch.empty((N, H, W, C), dtype=x.dtype, device=x.device)
        lib.sg_pad_crop_nhwc(ptr(x), ptr(y), code(x.dtype), N, h, w, H, W, C, stream())
        count_launch()
        ctx.src = (h, w)
        return y

    @staticmethod
    def backward(ctx, g):
        return PadCrop.apply(g, ctx.src[0], ctx.src[1]), None, None


def _tc_small_ok(x, W):
    """4x4 / 8x8 (or 8-wide) maps: run on the tensor-core kernels inside a zero-padded 8x16 frame."""
    co, ci, k, _ = W.shape
    if x.dtype != torch.bfloat16 or not _state["use_tc"] or k != 3:
        return False
    N, H, Wd, _ = x.shape
    if H > 8 or Wd > 8 or (H >= 8 and Wd >= 16):
        return False
    from . import _tc
    return _tc.conv_and_wgrad_supported(N, 8, 16, ci, co)


def _conv_same_padded(x, W, bias, kind, scale, bias_mul, act):
    N, H, Wd, _ = x.shape
    xp = PadCrop.apply(x, 8, 16)
    if bias is None and not act:
        yp = Conv.apply(xp, W, kind, "fwd", scale)
    else:
        yp = ConvBiasAct.apply(xp, W, bias, kind, scale, bias_mul, act)
    return PadCrop.apply(yp, H, Wd)


def conv_layer(x, W, bias, kind, scale, bias_mul, act):
    """One EqualizedConv2d evaluation on NHWC activations.  In bf16 mode the up/down-sampling convolutions run as
    the reference's unfused formulation (nearest-up -> conv3x3, conv3x3 -> 2x2 box; CustomLayers.py:153-154,166-176)
    so that all their FLOPs are on the tcgen05 3x3 kernels; otherwise as one tap-merged convolution."""
    co, ci, k, _ = W.shape
    if kind in ("same", "same_flip") and _tc_small_ok(x, W):
        return _conv_same_padded(x, W, bias, kind, scale, bias_mul, act)
    if kind in ("up", "up_fused", "down") and x.dtype == torch.bfloat16 and _state["use_tc"] and k == 3:
        from . import _tc
        N, H, Wd, _ = x.shape
        Hh, Wh = (2 * H, 2 * Wd) if kind != "down" else (H, Wd)
        if _state["fused_updown"] and _tc.updown_supported(kind, N, Hh // 2, Wh // 2, ci, co):
            # tap-merged tcgen05 kernel (16 tap-GEMMs per low-res pixel, no up-sampled / full-res intermediates)
            if bias is None and not act:
                return Conv.apply(x, W, kind, "fwd", scale)
            return ConvBiasAct.apply(x, W, bias, kind, scale, bias_mul, act)
        if Hh <= 8 and Wh <= 8 and _tc.conv_and_wgrad_supported(N, 8, 16, ci, co):
            # tiny maps: same composition inside the padded 8x16 frame
            if kind == "down":
                y = _conv_same_padded(x, W, None, "same", scale, 1.0, False)
                return PoolNHWC.apply(y, bias, bias_mul, act, 0.25)
            k2 = "same" if kind == "up" else "same_flip"
            return _conv_same_padded(UpNHWC.apply(x, 1.0), W, bias, k2, scale, bias_mul, act)
        if _tc.conv_and_wgrad_supported(N, Hh, Wh, ci, co):
            if kind == "down":
                y = Conv.apply(x, W, "same", "fwd", scale)
                return PoolNHWC.apply(y, bias, bias_mul, act, 0.25)
            k2 = "same" if kind == "up" else "same_flip"
            xu = UpNHWC.apply(x, 1.0)
            if bias is None and not act:
                return Conv.apply(xu, W, k2, "fwd", scale)
            return ConvBiasAct.apply(xu, W, bias, k2, scale, bias_mul, act)
    if bias is None and not act:
        return Conv.apply(x, W, kind, "fwd", scale)
    return ConvBiasAct.apply(x, W, bias, kind, scale, bias_mul, act)
