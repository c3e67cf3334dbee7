"""SeparableConvLayer -- per-pixel separable local convolution (reference:
my_package/functions/SeparableConvLayer.py:10-88; that file cannot even be imported in the
reference, `import ___ext.my_lib` at :4, and has no Module or caller).

`SeparableConvLayer(filtersize)`; forward(input1 [B,3,H,W], vertical [B,fs,Ho,Wo],
horizontal [B,fs,Ho,Wo]) -> [B,3,Ho,Wo], Ho = H-fs+1, Wo = W-fs+1 (reference :16-32).
"""
import torch
from torch.autograd import Function

from memc_b200 import lib as _lib
from ._base import fast_call, prep


class _SeparableConvFunction(Function):
    @staticmethod
    def forward(ctx, input1, input2, input3, filtersize):
        input1, input2, input3 = prep(input1, "input1"), prep(input2, "input2"), prep(input3, "input3")
        _lib.check_same_device(input1, input2, input3)
        B, C, H, W = input1.shape
        fs = min(input2.size(1), input3.size(1))
        Ho, Wo = min(input2.size(2), input3.size(2)), min(input2.size(3), input3.size(3))
        # reference :22-24
        assert H - filtersize == Ho - 1
        assert W - filtersize == Wo - 1
        assert fs == filtersize
        if input2.shape != (B, fs, Ho, Wo) or input3.shape != (B, fs, Ho, Wo):
            raise _lib.MemcB200Error("SeparableConv: filters must both be [B,fs,H-fs+1,W-fs+1]")
        output = torch.empty((B, C, Ho, Wo), dtype=input1.dtype, device=input1.device)
        fast_call("memc_b200_separable_conv_forward", _lib.stream_ptr(input1), B, C, H, W, fs,
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
        _lib.check_same_device(gradoutput, *ctx.saved_tensors)
        B, C, H, W = input1.shape
        gi1, gi2, gi3 = torch.empty_like(input1), torch.empty_like(input2), torch.empty_like(input3)
        fast_call("memc_b200_separable_conv_backward", _lib.stream_ptr(input1), B, C, H, W, ctx.fs,
                  _lib.strides_of(input1), _lib.strides_of(input2), _lib.strides_of(input3),
                  _lib.strides_of(gradoutput), _lib.strides_of(gi1), _lib.strides_of(gi2),
                  _lib.strides_of(gi3), _lib.ptr(input1), _lib.ptr(input2), _lib.ptr(input3),
                  _lib.ptr(gradoutput), _lib.ptr(gi1), _lib.ptr(gi2), _lib.ptr(gi3), _lib.OVERWRITE)
        return gi1, gi2, gi3, None


class SeparableConvLayer(object):
    def __init__(self, filtersize):
        self.filtersize = filtersize

    def __call__(self, input1, input2, input3):
        return _SeparableConvFunction.apply(input1, input2, input3, self.filtersize)

    forward = __call__
