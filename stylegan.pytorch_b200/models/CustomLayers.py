"""Layer primitives with the reference's names, constructor signatures and state_dict keys
(huangzh13/StyleGAN.pytorch models/CustomLayers.py), executing on the sgb200 CUDA kernels.

Calling convention: every module takes and returns NCHW-shaped tensors like the reference.  Internally feature maps
are NHWC-contiguous; a module output is the NHWC result viewed as NCHW (channels-last strides), so chains of these
modules never copy.  A plain NCHW-contiguous input is re-laid-out once on entry."""
import numpy as np
import torch
import torch.nn as nn

from . import _ops
from ._ops import he_scale

__all__ = ["PixelNormLayer", "Upscale2d", "Downscale2d", "EqualizedLinear", "EqualizedConv2d", "NoiseLayer",
           "StyleMod", "LayerEpilogue", "BlurLayer", "View", "StddevLayer", "Truncation"]


def to_nhwc(x, dtype=None):
    """NCHW-shaped tensor (any strides) -> NHWC-contiguous tensor in the engine dtype."""
    _ops._require_cuda(x)
    y = x.permute(0, 2, 3, 1)
    if not y.is_contiguous():
        y = y.contiguous()
    dtype = dtype or _ops.precision()
    if y.dtype != dtype:
        y = _ops.Cast.apply(y, dtype)
    return y


def to_nchw(y):
    """NHWC-contiguous -> NCHW-shaped view (no copy)."""
    return y.permute(0, 3, 1, 2)


class _Noise:
    """Engine-side RNG stream for NoiseLayer (reference draws torch.randn on the device, CustomLayers.py:193)."""
    seed = None
    offset = 0

    @classmethod
    def draw(cls, shape, device):
        if cls.seed is None:
            cls.seed = torch.initial_seed()
        n = int(np.prod(shape))
        out = _ops.randn_device(shape, device, cls.seed, cls.offset)
        cls.offset += (n + 3) // 4
        return out


def manual_seed(seed):
    _Noise.seed = int(seed)
    _Noise.offset = 0


class PixelNormLayer(nn.Module):
    """x * rsqrt(mean_c(x^2) + eps)  (reference CustomLayers.py:17-23)."""

    def __init__(self, epsilon=1e-8):
        super().__init__()
        self.epsilon = epsilon

    def forward(self, x):
        if x.dim() == 2:
            return _ops.PixelNorm.apply(x, self.epsilon)
        h = to_nhwc(x, torch.float32)
        n, hh, ww, c = h.shape
        y = _ops.PixelNorm.apply(h.reshape(-1, c), self.epsilon).reshape(n, hh, ww, c)
        return to_nchw(_ops.Cast.apply(y, _ops.precision()))


class Upscale2d(nn.Module):
    """Nearest-neighbour up-sampling by an integer factor, times gain (reference CustomLayers.py:26-45)."""

    @staticmethod
    def upscale2d(x, factor=2, gain=1):
        assert x.dim() == 4
        if factor == 1 and gain == 1:
            return x
        y = _ops.UpsampleNCHW.apply(x.float().contiguous(), factor, float(gain))
        return y if x.dtype == torch.float32 else _ops.Cast.apply(y, x.dtype)

    def __init__(self, factor=2, gain=1):
        super().__init__()
        assert isinstance(factor, int) and factor >= 1
        self.gain = gain
        self.factor = factor

    def forward(self, x):
        return self.upscale2d(x, factor=self.factor, gain=self.gain)


class Downscale2d(nn.Module):
    """Box down-sampling (reference CustomLayers.py:48-76).  Keeps the `blur.kernel` buffer of the reference."""

    def __init__(self, factor=2, gain=1):
        super().__init__()
        assert isinstance(factor, int) and factor >= 1
        self.factor = factor
        self.gain = gain
        if factor == 2:
            f = [np.sqrt(gain) / factor] * factor
            self.blur = BlurLayer(kernel=f, normalize=False, stride=factor)
        else:
            self.blur = None

    def forward(self, x):
        assert x.dim() == 4
        if self.factor == 1 and self.gain == 1:
            return x
        f = self.factor
        y = _ops.PoolNCHW.apply(x.float().contiguous(), f, float(self.gain) / (f * f))
        return y if x.dtype == torch.float32 else _ops.Cast.apply(y, x.dtype)


class EqualizedLinear(nn.Module):
    """Dense layer with equalized learning rate (reference CustomLayers.py:79-103)."""

    def __init__(self, input_size, output_size, gain=2 ** 0.5, use_wscale=False, lrmul=1, bias=True):
        super().__init__()
        he_std = gain * input_size ** (-0.5)
        if use_wscale:
            init_std, self.w_mul = 1.0 / lrmul, he_std * lrmul
        else:
            init_std, self.w_mul = he_std / lrmul, lrmul
        self.weight = nn.Parameter(torch.randn(output_size, input_size) * init_std)
        if bias:
            self.bias = nn.Parameter(torch.zeros(output_size))
            self.b_mul = lrmul
        else:
            self.bias = None

    def forward(self, x, act=False):
        b_mul = self.b_mul if self.bias is not None else 1.0
        lead = x.shape[:-1]
        x2 = x.reshape(-1, x.shape[-1]) if x.dim() != 2 else x
        if x2.dtype != torch.float32:
            x2 = _ops.Cast.apply(x2.contiguous(), torch.float32)
        y = _ops.Linear.apply(x2, self.weight, self.bias, self.w_mul, b_mul, act)
        return y.reshape(*lead, y.shape[-1]) if x.dim() != 2 else y


class EqualizedConv2d(nn.Module):
    """Conv layer with equalized learning rate, optional fused up/down-sampling (reference CustomLayers.py:106-180).

    The reference's size-dependent branches are preserved exactly:
      upscale   & out >= 128 : conv_transpose2d of the box-summed 4x4 kernel      -> kind 'up_fused'
      upscale   & out <  128 : nearest x2 then conv3x3                            -> kind 'up'
      downscale             : conv3x3 then 2x2 box == 4x4 stride-2 averaged kernel -> kind 'down' (both branches)
    each evaluated as one tap-table convolution on the low-resolution side (2.25x fewer MACs than the reference)."""

    def __init__(self, input_channels, output_channels, kernel_size, stride=1, gain=2 ** 0.5, use_wscale=False,
                 lrmul=1, bias=True, intermediate=None, upscale=False, downscale=False):
        super().__init__()
        self.upscale = Upscale2d() if upscale else None
        self.downscale = Downscale2d() if downscale else None
        he_std = gain * (input_channels * kernel_size ** 2) ** (-0.5)
        self.kernel_size = kernel_size
        if use_wscale:
            init_std, self.w_mul = 1.0 / lrmul, he_std * lrmul
        else:
            init_std, self.w_mul = he_std / lrmul, lrmul
        self.weight = nn.Parameter(torch.randn(output_channels, input_channels, kernel_size, kernel_size) * init_std)
        if bias:
            self.bias = nn.Parameter(torch.zeros(output_channels))
            self.b_mul = lrmul
        else:
            self.bias = None
        self.intermediate = intermediate

    # ---- functional NHWC core (used by the network fast paths) ----
    def kind_for(self, h, w):
        if self.upscale is not None:
            return "up_fused" if min(h, w) * 2 >= 128 else "up"
        if self.downscale is not None:
            return "down"
        return "same"

    def conv_nhwc(self, x, with_bias=True, act=False):
        """x NHWC -> conv (+bias) (+lrelu), WITHOUT the `intermediate` module."""
        if self.kernel_size != 3 and (self.upscale is not None or self.downscale is not None):
            raise NotImplementedError("sgb200: fused up/down-sampling is implemented for 3x3 kernels")
        kind = self.kind_for(x.shape[1], x.shape[2])
        bias = self.bias if (with_bias and self.bias is not None) else None
        b_mul = self.b_mul if self.bias is not None else 1.0
        return _ops.conv_layer(x, self.weight, bias, kind, self.w_mul, b_mul, act)

    def forward_nhwc(self, x):
        if self.downscale is not None:
            assert self.intermediate is None or x.shape[1] >= 128        # reference CustomLayers.py:167
        if self.intermediate is None:
            return self.conv_nhwc(x, with_bias=True)
        y = self.conv_nhwc(x, with_bias=False)
        b_mul = self.b_mul if self.bias is not None else 1.0
        if isinstance(self.intermediate, BlurLayer) and self.intermediate.is_standard:
            return _ops.Blur.apply(y, self.bias, b_mul)                    # blur and the bias add in one pass
        y = to_nhwc(self.intermediate(to_nchw(y)))
        if self.bias is not None:
            y = _ops.BiasNoise.apply(y, self.bias, b_mul, None, None)
        return y

    def forward(self, x):
        return to_nchw(self.forward_nhwc(to_nhwc(x)))


class NoiseLayer(nn.Module):
    """Per-pixel noise with a learned per-channel weight (reference CustomLayers.py:183-200)."""

    def __init__(self, channels):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(channels))
        self.noise = None

    def get_noise(self, n, h, w, device, noise=None):
        if noise is None and self.noise is None:
            return _Noise.draw((n, 1, h, w), device)
        noise = self.noise if noise is None else noise
        if noise.device != device:
            noise = noise.to(device)
        return noise

    def forward(self, x, noise=None):
        h = to_nhwc(x)
        nz = self.get_noise(h.shape[0], h.shape[1], h.shape[2], h.device, noise)
        return to_nchw(_ops.BiasNoise.apply(h, None, 1.0, nz, self.weight))


class StyleMod(nn.Module):
    """x * (style_scale + 1) + style_shift from an affine map of the dlatent (reference CustomLayers.py:203-216)."""

    def __init__(self, latent_size, channels, use_wscale):
        super(StyleMod, self).__init__()
        self.lin = EqualizedLinear(latent_size, channels * 2, gain=1.0, use_wscale=use_wscale)

    def forward(self, x, latent):
        style = self.lin(latent)
        h = to_nhwc(x)
        c = h.shape[-1]
        ab = torch.stack([style[:, :c] + 1.0, style[:, c:]], dim=2)
        return to_nchw(_ops.AffineNC.apply(h, ab))


class LayerEpilogue(nn.Module):
    """noise -> activation -> [pixel norm] -> [instance norm] -> [style mod] (reference CustomLayers.py:219-248).

    The default StyleGAN configuration (LeakyReLU, instance norm, no pixel norm) runs as the fused Epilogue kernels;
    other configurations fall back to the stand-alone layer kernels in sequence."""

    def __init__(self, channels, dlatent_size, use_wscale, use_noise, use_pixel_norm, use_instance_norm, use_styles,
                 activation_layer):
        super().__init__()
        from collections import OrderedDict
        layers = []
        if use_noise:
            layers.append(("noise", NoiseLayer(channels)))
        layers.append(("activation", activation_layer))
        if use_pixel_norm:
            layers.append(("pixel_norm", PixelNormLayer()))
        if use_instance_norm:
            layers.append(("instance_norm", nn.InstanceNorm2d(channels)))
        self.top_epi = nn.Sequential(OrderedDict(layers))
        self.style_mod = StyleMod(dlatent_size, channels, use_wscale=use_wscale) if use_styles else None
        self._fusable = (not use_pixel_norm) and use_instance_norm and isinstance(activation_layer, nn.LeakyReLU) \
            and abs(activation_layer.negative_slope - 0.2) < 1e-12

    def forward_nhwc(self, x, dlatent=None, pre_bias=None, pre_bias_mul=1.0):
        """x NHWC (conv output; `pre_bias` = that conv's bias, folded into the first pass)."""
        if self.style_mod is None:
            assert dlatent is None
        if not self._fusable:
            if pre_bias is not None:
                x = _ops.BiasNoise.apply(x, pre_bias, pre_bias_mul, None, None)
            y = self.top_epi_general(to_nchw(x))
            if self.style_mod is not None:
                y = self.style_mod(y, dlatent)
            return to_nhwc(y)
        noise_mod = self.top_epi.noise if hasattr(self.top_epi, "noise") else None
        noise = nw = None
        if noise_mod is not None:
            noise = noise_mod.get_noise(x.shape[0], x.shape[1], x.shape[2], x.device)
            nw = noise_mod.weight
        style = self._take_style(dlatent)
        eps = self.top_epi.instance_norm.eps
        return _ops.Epilogue.apply(x, pre_bias, pre_bias_mul, noise, nw, style, eps, True)

    def _take_style(self, dlatent):
        """The StyleMod affine of this layer: precomputed for the whole synthesis pass by GSynthesis (one grouped launch,
        _ops.StyleAffines) or evaluated here."""
        if self.style_mod is None:
            return None
        pre = getattr(self, "_style_pre", None)
        if pre is not None:
            self._style_pre = None
            return pre
        return self.style_mod.lin(dlatent)

    def conv_fused_nhwc(self, conv, x, dlatent):
        """conv3x3 + this epilogue with noise / LeakyReLU / InstanceNorm statistics inside the convolution kernel
        (_ops.ConvEpilogue); None when the configuration or the shape is outside that kernel's domain."""
        from . import _tc
        if not self._fusable or conv.kernel_size != 3 or conv.upscale is not None or conv.downscale is not None \
                or conv.intermediate is not None or not _tc.conv_epilogue_supported(x, conv.weight):
            return None
        noise_mod = self.top_epi.noise if hasattr(self.top_epi, "noise") else None
        noise = nw = None
        if noise_mod is not None:
            noise = noise_mod.get_noise(x.shape[0], x.shape[1], x.shape[2], x.device)
            nw = noise_mod.weight
        style = self._take_style(dlatent)
        b_mul = conv.b_mul if conv.bias is not None else 1.0
        return _ops.ConvEpilogue.apply(x, conv.weight, conv.bias, conv.w_mul, b_mul, noise, nw, style,
                                       self.top_epi.instance_norm.eps)

    def blur_fused_nhwc(self, conv, x, dlatent):
        """up-conv -> [1,2,1] blur -> +bias -> this epilogue, with everything after the convolution up to the
        InstanceNorm statistics in one pass (_ops.BlurEpilogue); None outside that kernel's domain."""
        if not self._fusable or not isinstance(conv.intermediate, BlurLayer) or not conv.intermediate.is_standard:
            return None
        y = conv.conv_nhwc(x, with_bias=False)
        if not _ops.BlurEpilogue.supported(y):
            b_mul = conv.b_mul if conv.bias is not None else 1.0
            return self.forward_nhwc(_ops.Blur.apply(y, conv.bias, b_mul), dlatent)
        noise_mod = self.top_epi.noise if hasattr(self.top_epi, "noise") else None
        noise = nw = None
        if noise_mod is not None:
            noise = noise_mod.get_noise(y.shape[0], y.shape[1], y.shape[2], y.device)
            nw = noise_mod.weight
        style = self._take_style(dlatent)
        b_mul = conv.b_mul if conv.bias is not None else 1.0
        return _ops.BlurEpilogue.apply(y, conv.bias, b_mul, noise, nw, style, self.top_epi.instance_norm.eps)

    def top_epi_general(self, x):
        for name, m in self.top_epi.named_children():
            if name == "activation":
                if isinstance(m, nn.LeakyReLU) and abs(m.negative_slope - 0.2) < 1e-12:
                    x = to_nchw(_ops.BiasNoise.apply(to_nhwc(x), None, 1.0, None, None, True))
                else:
                    raise NotImplementedError("sgb200: only LeakyReLU(0.2) has a kernel")
            elif name == "instance_norm":
                x = to_nchw(_ops.Epilogue.apply(to_nhwc(x), None, 1.0, None, None, None, m.eps, False))
            else:
                x = m(x)
        return x

    def forward(self, x, dlatents_in_slice=None):
        return to_nchw(self.forward_nhwc(to_nhwc(x), dlatents_in_slice))


class BlurLayer(nn.Module):
    """Depthwise separable low-pass filter (reference CustomLayers.py:251-276).  The kernel stays a registered buffer
    (checkpoint key contract); the [1,2,1] stride-1 filter and the 2x2 stride-2 box have kernels."""

    def __init__(self, kernel=None, normalize=True, flip=False, stride=1):
        super(BlurLayer, self).__init__()
        if kernel is None:
            kernel = [1, 2, 1]
        k1 = torch.tensor(kernel, dtype=torch.float32)
        k2 = (k1[:, None] * k1[None, :])[None, None]
        if normalize:
            k2 = k2 / k2.sum()
        if flip:
            k2 = k2.flip(2, 3)
        self.register_buffer("kernel", k2)
        self.stride = stride
        ref = torch.tensor([1.0, 2.0, 1.0])
        ref = (ref[:, None] * ref[None, :] / 16.0)[None, None]
        self.is_standard = stride == 1 and tuple(k2.shape) == (1, 1, 3, 3) and bool(torch.allclose(k2, ref))
        self._box = stride == 2 and tuple(k2.shape) == (1, 1, 2, 2) and bool((k2 == k2.flatten()[0]).all())
        self._box_value = float(k2.flatten()[0])

    def forward(self, x):
        if self.is_standard:
            return to_nchw(_ops.Blur.apply(to_nhwc(x), None, 1.0))
        if self._box:
            y = _ops.PoolNCHW.apply(x.float().contiguous(), 2, self._box_value)
            return y if x.dtype == torch.float32 else _ops.Cast.apply(y, x.dtype)
        raise NotImplementedError("sgb200: BlurLayer kernels other than [1,2,1]/stride 1 and 2x2 box/stride 2")


class View(nn.Module):
    def __init__(self, *shape):
        super().__init__()
        self.shape = shape

    def forward(self, x):
        return x.reshape(x.size(0), *self.shape)


class StddevLayer(nn.Module):
    """Minibatch standard deviation feature (reference CustomLayers.py:288-305)."""

    def __init__(self, group_size=4, num_new_features=1):
        super().__init__()
        self.group_size = group_size
        self.num_new_features = num_new_features
        if num_new_features != 1:
            raise NotImplementedError("sgb200: StddevLayer supports num_new_features=1 (the StyleGAN setting)")

    def forward_nhwc(self, x):
        return _ops.MbStd.apply(x if x.dtype == torch.float32 else _ops.Cast.apply(x, torch.float32),
                                self.group_size)

    def forward(self, x):
        return to_nchw(self.forward_nhwc(to_nhwc(x, torch.float32)))


class Truncation(nn.Module):
    """Truncation trick towards the running average dlatent (reference CustomLayers.py:308-323)."""

    def __init__(self, avg_latent, max_layer=8, threshold=0.7, beta=0.995):
        super().__init__()
        self.max_layer = max_layer
        self.threshold = threshold
        self.beta = beta
        self.register_buffer("avg_latent", avg_latent)

    def update(self, last_avg):
        with torch.no_grad():
            new = _ops.Axpby.apply(self.avg_latent, last_avg.detach().reshape(-1).contiguous(), self.beta,
                                   1.0 - self.beta)
            self.avg_latent.copy_(new)

    def forward(self, x):
        assert x.dim() == 3
        b, l, d = x.shape
        if x.stride(1) == 0:                      # broadcast [B,L,D] from g_mapping: one source row per sample
            return _ops.Dlatents.apply(x[:, 0].contiguous(), None, self.avg_latent, l, l, self.threshold,
                                       self.max_layer)
        # general per-layer dlatents: treat every (b,l) row as its own sample with L=1, cut-off by layer index
        rows = x.reshape(b * l, d).contiguous()
        lo = _ops.Dlatents.apply(rows, None, self.avg_latent, 1, 1, self.threshold, 1).reshape(b, l, d)
        mask = (torch.arange(l, device=x.device) < self.max_layer).view(1, -1, 1)
        return torch.where(mask, lo, x)
