"""Composite blocks with the reference's names and state_dict layout (huangzh13/StyleGAN.pytorch
models/Blocks.py), executing on the sgb200 kernels.  `forward` keeps the reference's NCHW call signature; the
`*_nhwc` methods are the zero-copy paths the networks in GAN.py use."""
from collections import OrderedDict

import torch
import torch.nn as nn

from . import _ops
from .CustomLayers import (BlurLayer, EqualizedConv2d, EqualizedLinear, LayerEpilogue, StddevLayer, View, to_nchw,
                           to_nhwc)


def _conv_then_epilogue(conv, epi, x, dlatent):
    """conv -> [blur] -> (+bias) -> LayerEpilogue with the bias add folded into the epilogue's first pass."""
    b_mul = conv.b_mul if conv.bias is not None else 1.0
    if conv.intermediate is None:
        fused = epi.conv_fused_nhwc(conv, x, dlatent)
        if fused is not None:
            return fused
        y = conv.conv_nhwc(x, with_bias=False)
        return epi.forward_nhwc(y, dlatent, pre_bias=conv.bias, pre_bias_mul=b_mul)
    fused = epi.blur_fused_nhwc(conv, x, dlatent)
    if fused is not None:
        return fused
    return epi.forward_nhwc(conv.forward_nhwc(x), dlatent)


class InputBlock(nn.Module):
    """4x4 block: learned constant (or dense) -> epilogue -> conv3x3 -> epilogue (reference Blocks.py:17-60)."""

    def __init__(self, nf, dlatent_size, const_input_layer, gain, use_wscale, use_noise, use_pixel_norm,
                 use_instance_norm, use_styles, activation_layer):
        super().__init__()
        self.const_input_layer = const_input_layer
        self.nf = nf
        if self.const_input_layer:
            self.const = nn.Parameter(torch.ones(1, nf, 4, 4))
            self.bias = nn.Parameter(torch.ones(nf))
        else:
            self.dense = EqualizedLinear(dlatent_size, nf * 16, gain=gain / 4, use_wscale=use_wscale)
        self.epi1 = LayerEpilogue(nf, dlatent_size, use_wscale, use_noise, use_pixel_norm, use_instance_norm,
                                  use_styles, activation_layer)
        self.conv = EqualizedConv2d(nf, nf, 3, gain=gain, use_wscale=use_wscale)
        self.epi2 = LayerEpilogue(nf, dlatent_size, use_wscale, use_noise, use_pixel_norm, use_instance_norm,
                                  use_styles, activation_layer)

    def forward_nhwc(self, dlatents_in_range):
        batch = dlatents_in_range.size(0)
        dt = _ops.precision()
        if self.const_input_layer:
            # [1,C,4,4] parameter -> NHWC rows, broadcast over the batch; the per-channel bias rides along in the
            # epilogue's first pass (reference Blocks.py:51-52)
            c = self.const.permute(0, 2, 3, 1).expand(batch, -1, -1, -1)
            # This first epilogue always runs in fp32 (16 pixels x 512 channels per image: free).  Its input is the same
            # constant at every pixel plus noise_w[c] * noise, so a channel with a small noise weight has mean >> std;
            # InstanceNorm divides by that std, and rounding z to bf16 BEFORE the normalisation put up to 9 % of relative
            # error into such channels -- it was the dominant term of the bf16 image error at every depth (5 % -> see
            # tests/test_gpu_fullwidth.py).  The normalised output has unit variance and rounds to bf16 harmlessly.
            x = self.epi1.forward_nhwc(c.contiguous(), dlatents_in_range[:, 0], pre_bias=self.bias, pre_bias_mul=1.0)
            x = _ops.Cast.apply(x, dt)
        else:
            x = self.dense(dlatents_in_range[:, 0]).reshape(batch, self.nf, 4, 4)
            x = self.epi1.forward_nhwc(to_nhwc(x), dlatents_in_range[:, 0])
        return _conv_then_epilogue(self.conv, self.epi2, x, dlatents_in_range[:, 1])

    def forward(self, dlatents_in_range):
        return to_nchw(self.forward_nhwc(dlatents_in_range))


class GSynthesisBlock(nn.Module):
    """up-conv (+blur) -> epilogue -> conv3x3 -> epilogue (reference Blocks.py:63-88)."""

    def __init__(self, in_channels, out_channels, blur_filter, dlatent_size, gain, use_wscale, use_noise,
                 use_pixel_norm, use_instance_norm, use_styles, activation_layer):
        super().__init__()
        blur = BlurLayer(blur_filter) if blur_filter else None
        self.conv0_up = EqualizedConv2d(in_channels, out_channels, kernel_size=3, gain=gain, use_wscale=use_wscale,
                                        intermediate=blur, upscale=True)
        self.epi1 = LayerEpilogue(out_channels, dlatent_size, use_wscale, use_noise, use_pixel_norm,
                                  use_instance_norm, use_styles, activation_layer)
        self.conv1 = EqualizedConv2d(out_channels, out_channels, kernel_size=3, gain=gain, use_wscale=use_wscale)
        self.epi2 = LayerEpilogue(out_channels, dlatent_size, use_wscale, use_noise, use_pixel_norm,
                                  use_instance_norm, use_styles, activation_layer)

    def forward_nhwc(self, x, dlatents_in_range):
        x = _conv_then_epilogue(self.conv0_up, self.epi1, x, dlatents_in_range[:, 0])
        return _conv_then_epilogue(self.conv1, self.epi2, x, dlatents_in_range[:, 1])

    def forward(self, x, dlatents_in_range):
        return to_nchw(self.forward_nhwc(to_nhwc(x), dlatents_in_range))


class DiscriminatorTop(nn.Sequential):
    """minibatch-stddev -> conv3x3 -> lrelu -> flatten -> dense -> lrelu -> dense (reference Blocks.py:91-134).
    Runs in fp32 regardless of the engine precision (0.2 % of the FLOPs, and it feeds the loss)."""

    def __init__(self, mbstd_group_size, mbstd_num_features, in_channels, intermediate_channels, gain, use_wscale,
                 activation_layer, resolution=4, in_channels2=None, output_features=1, last_gain=1):
        layers = []
        if mbstd_group_size > 1:
            layers.append(("stddev_layer", StddevLayer(mbstd_group_size, mbstd_num_features)))
        if in_channels2 is None:
            in_channels2 = in_channels
        layers.append(("conv", EqualizedConv2d(in_channels + mbstd_num_features, in_channels2, kernel_size=3,
                                               gain=gain, use_wscale=use_wscale)))
        layers.append(("act0", activation_layer))
        layers.append(("view", View(-1)))
        layers.append(("dense0", EqualizedLinear(in_channels2 * resolution * resolution, intermediate_channels,
                                                 gain=gain, use_wscale=use_wscale)))
        layers.append(("act1", activation_layer))
        layers.append(("dense1", EqualizedLinear(intermediate_channels, output_features, gain=last_gain,
                                                 use_wscale=use_wscale)))
        super().__init__(OrderedDict(layers))

    def forward_nhwc(self, x):
        conv = self.conv
        b_mul = conv.b_mul if conv.bias is not None else 1.0
        if hasattr(self, "stddev_layer") and conv.weight.shape[1] == x.shape[-1] + 1 and x.shape[-1] % 64 == 0:
            # conv over [x, stddev] = conv_x(x) + conv_s(stddev): the 513-channel kernel is split so that the wide
            # part keeps the vector / tensor-core path (an odd channel count would force the scalar kernel) and only
            # the single stddev channel goes through the generic one.  Same arithmetic, summed in fp32.
            c = x.shape[-1]
            xf = x if x.dtype == torch.float32 else _ops.Cast.apply(x, torch.float32)
            std = self.stddev_layer.forward_nhwc(xf)[..., c:].contiguous()
            main = _ops.conv_layer(x, conv.weight[:, :c], None, "same", conv.w_mul, 1.0, False)
            if main.dtype != torch.float32:
                main = _ops.Cast.apply(main, torch.float32)
            extra = _ops.Conv.apply(std, conv.weight[:, c:], "same", "fwd", conv.w_mul)
            x = _ops.BiasNoise.apply(_ops.Axpby.apply(main, extra, 1.0, 1.0), conv.bias, b_mul, None, None, True)
        else:
            if x.dtype != torch.float32:
                x = _ops.Cast.apply(x, torch.float32)
            if hasattr(self, "stddev_layer"):
                x = self.stddev_layer.forward_nhwc(x)
            x = conv.conv_nhwc(x, with_bias=True, act=True)
        x = x.permute(0, 3, 1, 2).reshape(x.shape[0], -1)       # flatten in the reference's NCHW feature order
        x = self.dense0(x, act=True)
        return self.dense1(x)

    def forward(self, x):
        return self.forward_nhwc(to_nhwc(x, torch.float32))


class DiscriminatorBlock(nn.Sequential):
    """conv3x3 -> lrelu -> blur -> conv3x3+down -> lrelu (reference Blocks.py:137-146)."""

    def __init__(self, in_channels, out_channels, gain, use_wscale, activation_layer, blur_kernel):
        super().__init__(OrderedDict([
            ("conv0", EqualizedConv2d(in_channels, in_channels, kernel_size=3, gain=gain, use_wscale=use_wscale)),
            ("act0", activation_layer),
            ("blur", BlurLayer(kernel=blur_kernel)),
            ("conv1_down", EqualizedConv2d(in_channels, out_channels, kernel_size=3, gain=gain,
                                           use_wscale=use_wscale, downscale=True)),
            ("act1", activation_layer)]))

    def forward_nhwc(self, x):
        x = self.conv0.conv_nhwc(x, with_bias=True, act=True)
        if self.blur.is_standard:
            x = _ops.Blur.apply(x, None, 1.0)
        else:
            x = to_nhwc(self.blur(to_nchw(x)))
        return self.conv1_down.conv_nhwc(x, with_bias=True, act=True)

    def forward(self, x):
        return to_nchw(self.forward_nhwc(to_nhwc(x)))
