"""WeightedFlowProjectionLayer -- FlowProjection gated by brightness constancy.

The reference ships the C side of this op (my_package/src/my_lib_kernel.cu:2499-3024, FFI entry
my_lib_cuda.c:986-1140) but no Python class; this one follows the conventions of the reference's FlowProjectionLayer
(my_package/functions/FlowProjectionLayer.py:6-70): `WeightedFlowProjectionLayer(requires_grad, threshold)`, fill-hole only
when the flow did not require grad, `count` kept for backward; the frames get no gradient (the reference computes none).

    input1  [B, 2, H, W]  flow frame0 -> frame2
    input2  [B, 3, H, W]  frame0;   input3 [B, 3, H, W] frame2
    ->      [B, 2, H, W]  mean of -flow over the sources that land on a pixel AND whose brightness-constancy error
                          mean_c |frame0[p] - frame2[p + 2 flow]| + 1e-8 is <= threshold;
            `.weight` [B, 1, H, W] holds the mean error of those sources.
"""
import torch
from torch.autograd import Function

from memc_b200 import lib as _lib
from ._base import fast_call, prep


class _WeightedFlowProjectionFunction(Function):
    @staticmethod
    def forward(ctx, input1, input2, input3, fillhole, threshold):
        input1, input2, input3 = prep(input1, "input1"), prep(input2, "input2"), prep(input3, "input3")
        _lib.check_same_device(input2, input1)
        _lib.check_same_device(input3, input1)
        B, C, H, W = input1.shape
        if C != 2:  # my_lib_cuda.c:1001
            raise _lib.MemcB200Error("WeightedFlowProjection: input1 must have 2 channels, got %d" % C)
        for name, t in (("input2", input2), ("input3", input3)):  # my_lib_cuda.c:1002-1003
            if tuple(t.shape) != (B, 3, H, W):
                raise _lib.MemcB200Error("WeightedFlowProjection: %s must be [B,3,H,W], got %s" % (name, tuple(t.shape)))
        count = torch.empty((B, 1, H, W), dtype=input1.dtype, device=input1.device)
        weight = torch.empty_like(count)
        output = torch.empty_like(input1)
        S, P = _lib.strides_of, _lib.ptr
        fast_call("memc_b200_weighted_flow_projection_forward", _lib.stream_ptr(input1), B, H, W, int(fillhole), float(threshold),
                  S(input1), S(input2), S(input3), S(count), S(weight), S(output),
                  P(input1), P(input2), P(input3), P(count), P(weight), P(output), _lib.OVERWRITE)
        ctx.save_for_backward(input1, input2, input3, count)
        ctx.threshold = float(threshold)
        ctx.mark_non_differentiable(count, weight)
        return output, count, weight

    @staticmethod
    def backward(ctx, gradoutput, _gradcount, _gradweight):
        input1, input2, input3, count = ctx.saved_tensors
        gradoutput = prep(gradoutput, "gradoutput")
        _lib.check_same_device(gradoutput, input1)
        B, _, H, W = input1.shape
        gi = torch.empty_like(input1)
        S, P = _lib.strides_of, _lib.ptr
        fast_call("memc_b200_weighted_flow_projection_backward", _lib.stream_ptr(input1), B, H, W, ctx.threshold,
                  S(input1), S(input2), S(input3), S(count), S(gradoutput), S(gi),
                  P(input1), P(input2), P(input3), P(count), P(gradoutput), P(gi), _lib.OVERWRITE)
        return gi, None, None, None, None


class WeightedFlowProjectionLayer(object):
    """`.count` / `.weight` hold the last forward's vote counts and mean brightness errors."""

    def __init__(self, requires_grad, threshold=2.0):
        self.requires_grad = requires_grad
        self.fillhole = 1 if self.requires_grad == False else 0  # noqa: E712  (FlowProjectionLayer.py:15)
        self.threshold = float(threshold)
        self.count = None
        self.weight = None

    def __call__(self, input1, input2, input3):
        output, count, weight = _WeightedFlowProjectionFunction.apply(input1, input2, input3, self.fillhole, self.threshold)
        self.count, self.weight = count, weight
        return output

    forward = __call__
