"""InterpolationLayer -- plain bilinear backward warp (reference:
my_package/functions/InterpolationLayer.py:7-69).  The reference wrapper only admits C == 3
(my_lib_cuda.c:373); other channel counts go through the same kernel under its
"InterpolationCh" name (my_lib_cuda.c:481-533), which is what this class does for C != 3.
"""
import torch
from torch.autograd import Function

from memc_b200 import lib as _lib
from ._base import fast_call, prep


class _InterpolationFunction(Function):
    @staticmethod
    def forward(ctx, input1, input2):
        input1, input2 = prep(input1, "input1"), prep(input2, "input2")
        _lib.check_same_device(input1, input2)
        B, C, H, W = input1.shape
        if input2.shape != (B, 2, H, W):
            raise _lib.MemcB200Error("Interpolation: flow must be [B,2,H,W]")
        output = torch.empty_like(input1)
        fast_call("memc_b200_interpolation_forward", _lib.stream_ptr(input1), B, C, H, W,
                  _lib.strides_of(input1), _lib.strides_of(input2), _lib.strides_of(output),
                  _lib.ptr(input1), _lib.ptr(input2), _lib.ptr(output), _lib.OVERWRITE)
        ctx.save_for_backward(input1, input2)
        return output

    @staticmethod
    def backward(ctx, gradoutput):
        input1, input2 = ctx.saved_tensors
        gradoutput = prep(gradoutput, "gradoutput")
        _lib.check_same_device(gradoutput, *ctx.saved_tensors)
        B, C, H, W = input1.shape
        gi1, gi2 = torch.empty_like(input1), torch.empty_like(input2)
        fast_call("memc_b200_interpolation_backward", _lib.stream_ptr(input1), B, C, H, W,
                  _lib.strides_of(input1), _lib.strides_of(input2), _lib.strides_of(gradoutput),
                  _lib.strides_of(gi1), _lib.strides_of(gi2), _lib.ptr(input1), _lib.ptr(input2),
                  _lib.ptr(gradoutput), _lib.ptr(gi1), _lib.ptr(gi2), _lib.OVERWRITE)
        return gi1, gi2


class InterpolationLayer(object):
    apply = staticmethod(_InterpolationFunction.apply)

    def __init__(self):
        pass

    def __call__(self, input1, input2):
        return _InterpolationFunction.apply(input1, input2)

    forward = __call__
