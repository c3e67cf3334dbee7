"""DepthFlowProjectionLayer -- FlowProjection with a per-source weight.

The reference ships the C side of this op (my_package/src/my_lib_kernel.cu:2053-2497, FFI entry
my_lib_cuda.c:857-985) but no Python class; this one follows the conventions of the reference's
FlowProjectionLayer (my_package/functions/FlowProjectionLayer.py:6-70): `DepthFlowProjectionLayer(requires_grad)`,
fill-hole only when the input did not require grad, the accumulated weight `count` and the forward's `output` are
kept for backward (the backward kernel reads both, my_lib_kernel.cu:2312-2357), gradients for both inputs.

    input1  [B, 2, H, W]  flow frame0 -> frame2
    input2  [B, 1, H, W]  weight of every source pixel (an inverse depth: nearer pixels win the vote)
    ->      [B, 2, H, W]  sum(-w * flow) / sum(w) over the sources that land on a pixel
"""
import torch
from torch.autograd import Function

from memc_b200 import lib as _lib
from ._base import fast_call, prep


class _DepthFlowProjectionFunction(Function):
    @staticmethod
    def forward(ctx, input1, input2, fillhole):
        input1, input2 = prep(input1, "input1"), prep(input2, "input2")
        _lib.check_same_device(input2, input1)
        B, C, H, W = input1.shape
        if C != 2:  # my_lib_cuda.c:869
            raise _lib.MemcB200Error("DepthFlowProjection: input1 must have 2 channels, got %d" % C)
        if tuple(input2.shape) != (B, 1, H, W):  # my_lib_cuda.c:875
            raise _lib.MemcB200Error("DepthFlowProjection: input2 must be [B,1,H,W], got %s" % (tuple(input2.shape),))
        count = torch.empty((B, 1, H, W), dtype=input1.dtype, device=input1.device)
        output = torch.empty_like(input1)
        fast_call("memc_b200_depth_flow_projection_forward", _lib.stream_ptr(input1), B, H, W, int(fillhole),
                  _lib.strides_of(input1), _lib.strides_of(input2), _lib.strides_of(count), _lib.strides_of(output),
                  _lib.ptr(input1), _lib.ptr(input2), _lib.ptr(count), _lib.ptr(output), _lib.OVERWRITE)
        ctx.save_for_backward(input1, input2, count, output)
        ctx.mark_non_differentiable(count)
        return output, count

    @staticmethod
    def backward(ctx, gradoutput, _gradcount):
        input1, input2, count, output = ctx.saved_tensors
        gradoutput = prep(gradoutput, "gradoutput")
        _lib.check_same_device(gradoutput, input1)
        B, _, H, W = input1.shape
        gi1, gi2 = torch.empty_like(input1), torch.empty_like(input2)
        S, P = _lib.strides_of, _lib.ptr
        fast_call("memc_b200_depth_flow_projection_backward", _lib.stream_ptr(input1), B, H, W,
                  S(input1), S(input2), S(count), S(output), S(gradoutput), S(gi1), S(gi2),
                  P(input1), P(input2), P(count), P(output), P(gradoutput), P(gi1), P(gi2), _lib.OVERWRITE)
        return gi1, gi2, None


class DepthFlowProjectionLayer(object):
    """`.count` holds the last forward's accumulated weights (as FlowProjectionLayer keeps its hit counts)."""

    def __init__(self, requires_grad):
        self.requires_grad = requires_grad
        self.fillhole = 1 if self.requires_grad == False else 0  # noqa: E712  (FlowProjectionLayer.py:15)
        self.count = None

    def __call__(self, input1, input2):
        output, count = _DepthFlowProjectionFunction.apply(input1, input2, self.fillhole)
        self.count = count
        return output

    forward = __call__

    @staticmethod
    def apply(input1, input2, requires_grad=None):
        rg = input1.requires_grad if requires_grad is None else requires_grad
        return DepthFlowProjectionLayer(rg)(input1, input2)
