"""Dispatch to the tcgen05 tensor-core convolution kernels (csrc/sg_conv_tc.cu) for bf16 activations.

Returns False when the layer shape is not covered; the caller then runs the generic CUDA-core kernel.  Both are
hand-written sm_100a kernels of this library -- there is no library or CPU fallback."""
import torch

from ._lib import lib, ptr, stream, count_launch
from . import _ops

_tc_weights = {}


def try_conv(x, W, kind, direction, scale, bias, bias_mul, act, y):
    return False


def try_wgrad(xl, dyl, kind, scale, dW):
    return False
