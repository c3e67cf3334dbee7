"""FilterInterpolationLayer -- adaptive warping op (reference:
my_package/functions/FilterInterpolationLayer.py:6-71).

forward(input1 [B,C,H,W], input2 = flow [B,2,H,W], input3 = per-pixel kernel [B,fs*fs,H,W])
-> [B,C,H,W]; backward returns gradients for ALL three inputs (reference :71).

Unlike the reference Function (which zero-fills output and three gradient tensors with
separate memset kernels, :28,46-48) the buffers here are `torch.empty` and the library is
called with MEMC_B200_OVERWRITE: it writes every element itself and zero-fills only the
scatter target (gradinput1).  How gradinput1 is accumulated inside a tile (fixed point / fp32 atomics) is
`memc_b200.lib.set_fi_accumulation()` / the environment variable MEMC_B200_FI_ACCUM.
"""
import math

import torch
from torch.autograd import Function

from memc_b200 import lib as _lib
from ._base import fast_call, prep


class _FilterInterpolationFunction(Function):
    @staticmethod
    def forward(ctx, input1, input2, input3):
        input1, input2, input3 = prep(input1, "input1"), prep(input2, "input2"), prep(input3, "input3")
        _lib.check_same_device(input1, input2, input3)
        B, C, H, W = input1.shape
        if input2.shape != (B, 2, H, W) or input3.shape[0] != B or input3.shape[2:] != (H, W):
            raise _lib.MemcB200Error("FilterInterpolation: inconsistent shapes %s %s %s" % (
                tuple(input1.shape), tuple(input2.shape), tuple(input3.shape)))
        fs = int(math.sqrt(float(input3.size(1))))  # my_lib_cuda.c:619-620
        output = torch.empty_like(input1)
        fast_call("memc_b200_filter_interpolation_forward", _lib.stream_ptr(input1), B, C, H, W, fs,
                  _lib.strides_of(input1), _lib.strides_of(input2), _lib.strides_of(input3),
                  _lib.strides_of(output), _lib.ptr(input1), _lib.ptr(input2), _lib.ptr(input3),
                  _lib.ptr(output), _lib.OVERWRITE)
        ctx.save_for_backward(input1, input2, input3)
        ctx.fs = fs
        return output

    @staticmethod
    def backward(ctx, gradoutput):
        input1, input2, input3 = ctx.saved_tensors
        gradoutput = prep(gradoutput, "gradoutput")
        _lib.check_same_device(input1, gradoutput)
        B, C, H, W = input1.shape
        gi1, gi2, gi3 = torch.empty_like(input1), torch.empty_like(input2), torch.empty_like(input3)
        fast_call("memc_b200_filter_interpolation_backward", _lib.stream_ptr(input1), B, C, H, W, ctx.fs,
                  _lib.strides_of(input1), _lib.strides_of(input2), _lib.strides_of(input3),
                  _lib.strides_of(gradoutput), _lib.strides_of(gi1), _lib.strides_of(gi2),
                  _lib.strides_of(gi3), _lib.ptr(input1), _lib.ptr(input2), _lib.ptr(input3),
                  _lib.ptr(gradoutput), _lib.ptr(gi1), _lib.ptr(gi2), _lib.ptr(gi3), _lib.fi_backward_flags())
        return gi1, gi2, gi3


class FilterInterpolationLayer(object):
    """Reference-named, reference-constructed (`FilterInterpolationLayer()`) callable."""

    apply = staticmethod(_FilterInterpolationFunction.apply)

    def __init__(self):
        pass

    def __call__(self, input1, input2, input3):
        return _FilterInterpolationFunction.apply(input1, input2, input3)

    forward = __call__
