"""FlowProjectionLayer -- forward-splat the flow to the mid time step (reference:
my_package/functions/FlowProjectionLayer.py:6-70).

`FlowProjectionLayer(requires_grad)`: fill-hole runs only when the input did not require
grad (reference :15, `fillhole = 1 if requires_grad == False else 0`) -- callers reproduce
the reference's inference behaviour by running under `torch.no_grad()`.
The per-pixel hit `count` of the forward pass is kept for backward (reference :37, :53);
backward ignores the hole fill, exactly as the reference kernel does.
"""
import torch
from torch.autograd import Function

from memc_b200 import lib as _lib
from ._base import fast_call, prep


class _FlowProjectionFunction(Function):
    @staticmethod
    def forward(ctx, input1, fillhole):
        input1 = prep(input1, "input1")
        B, C, H, W = input1.shape
        if C != 2:  # my_lib_cuda.c:763
            raise _lib.MemcB200Error("FlowProjection: input must have 2 channels, got %d" % C)
        count = torch.empty((B, 1, H, W), dtype=input1.dtype, device=input1.device)
        output = torch.empty_like(input1)
        fast_call("memc_b200_flow_projection_forward", _lib.stream_ptr(input1), B, H, W, int(fillhole),
                  _lib.strides_of(input1), _lib.strides_of(count), _lib.strides_of(output),
                  _lib.ptr(input1), _lib.ptr(count), _lib.ptr(output), _lib.OVERWRITE)
        ctx.save_for_backward(input1, count)
        ctx.mark_non_differentiable(count)
        return output, count

    @staticmethod
    def backward(ctx, gradoutput, _gradcount):
        input1, count = ctx.saved_tensors
        gradoutput = prep(gradoutput, "gradoutput")
        _lib.check_same_device(gradoutput, input1)
        B, _, H, W = input1.shape
        gi = torch.empty_like(input1)
        fast_call("memc_b200_flow_projection_backward", _lib.stream_ptr(input1), B, H, W,
                  _lib.strides_of(input1), _lib.strides_of(count), _lib.strides_of(gradoutput),
                  _lib.strides_of(gi), _lib.ptr(input1), _lib.ptr(count), _lib.ptr(gradoutput),
                  _lib.ptr(gi), _lib.OVERWRITE)
        return gi, None


class FlowProjectionLayer(object):
    """Reference-named callable; `.count` holds the last forward's hit counts (reference :37)."""

    def __init__(self, requires_grad):
        self.requires_grad = requires_grad
        self.fillhole = 1 if self.requires_grad == False else 0  # noqa: E712  (reference :15)
        self.count = None

    def __call__(self, input1):
        output, count = _FlowProjectionFunction.apply(input1, self.fillhole)
        self.count = count
        return output

    forward = __call__

    @staticmethod
    def apply(input1, requires_grad=None):
        rg = input1.requires_grad if requires_grad is None else requires_grad
        return FlowProjectionLayer(rg)(input1)
