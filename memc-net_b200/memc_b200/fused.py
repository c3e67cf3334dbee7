"""Fused call sites of the reference's networks -- optional fast
paths NEXT TO the unchanged `my_package` API.

    FilterInterpolate(ref0, ref2, offset, filter, occlusion)
        = occlusion[0] * FilterInterpolationModule()(ref0, offset[0], filter[0])
        + occlusion[1] * FilterInterpolationModule()(ref2, offset[1], filter[1])

    FlowProjectPair(flow_a, flow_b)
        = (FlowProject(flow_a), FlowProject(flow_b))     the bidirectional pair of networks/MEMC_Net.py:109-113

`FilterInterpolate` is the static method of the same name in networks/MEMC_Net.py:258-264 and
networks/MEMC_Net_star.py:272-278 (their sixth argument `filter_size2` is unused there and optional
here), so a network can rebind it:  `MEMC_Net.FilterInterpolate = staticmethod(fused.FilterInterpolate)`.

`FilterInterpolateMean(ref0, ref2, offset, filter)` is the same call site in networks/MEMC_Net_s.py:258-264, where the two
warps are averaged instead of occlusion-blended (`ref0_offset / 2.0 + ref2_offset / 2.0`): the same fused kernel with
the constant weight 0.5 (no occlusion maps are read).

`FilterInterpolateShared(input_a, input_b, offset, filter)` warps two images that share flow and filter -- the RGB frame and
its 64-channel context features in networks/MEMC_Net_star.py:272-285 -- in one kernel: flow, filter, geometry and bounding box
are fetched / computed once.

`FlowProjectPair` runs both directions as ONE FlowProjection call on the concatenated batch (frames are
independent, so each half equals the separate call; with 2 x B >= 3 frames the persistent pipeline has twice the
frames to overlap), and returns the two halves as views.

`FilterInterpolate` forward: one kernel warps both references and blends them in registers
(memc_b200_filter_interpolation_blend_forward), bit-identical to the composition.  Backward: composed
from the plain ops (FilterInterpolation backward per reference; the two warps are recomputed for the
occlusion gradients), so training code keeps working.  CUDA only: like every op of this package it
raises on CPU tensors -- there is no fallback.
"""
import math

import torch
from torch.autograd import Function

from . import lib as _lib


def _prep(t, name):
    return _lib.check_tensor(t, name).contiguous()


def _fi_forward(in1, flow, filt, fs):
    B, C, H, W = in1.shape
    out = torch.empty_like(in1)
    _lib.call("memc_b200_filter_interpolation_forward", _lib.stream_ptr(in1), B, C, H, W, fs, _lib.strides_of(in1),
              _lib.strides_of(flow), _lib.strides_of(filt), _lib.strides_of(out), _lib.ptr(in1), _lib.ptr(flow),
              _lib.ptr(filt), _lib.ptr(out), _lib.OVERWRITE)
    return out


def _fi_backward(in1, flow, filt, gout, fs):
    B, C, H, W = in1.shape
    g1, g2, g3 = torch.empty_like(in1), torch.empty_like(flow), torch.empty_like(filt)
    _lib.call("memc_b200_filter_interpolation_backward", _lib.stream_ptr(in1), B, C, H, W, fs, _lib.strides_of(in1),
              _lib.strides_of(flow), _lib.strides_of(filt), _lib.strides_of(gout), _lib.strides_of(g1), _lib.strides_of(g2),
              _lib.strides_of(g3), _lib.ptr(in1), _lib.ptr(flow), _lib.ptr(filt), _lib.ptr(gout), _lib.ptr(g1),
              _lib.ptr(g2), _lib.ptr(g3), _lib.fi_backward_flags())
    return g1, g2, g3


class _FilterInterpolateBlend(Function):
    @staticmethod
    def forward(ctx, ref0, flow0, filt0, occ0, ref1, flow1, filt1, occ1):
        mean = occ0 is None  # FilterInterpolateMean: constant weights 0.5, no occlusion maps
        ts = [None if t is None else _prep(t, n) for t, n in zip(
            (ref0, flow0, filt0, occ0, ref1, flow1, filt1, occ1),
            ("ref0", "offset[0]", "filter[0]", "occlusion[0]", "ref2", "offset[1]", "filter[1]", "occlusion[1]"))]
        ref0, flow0, filt0, occ0, ref1, flow1, filt1, occ1 = ts
        _lib.check_same_device(*ts)
        B, C, H, W = ref0.shape
        ok = (ref1.shape == ref0.shape and flow0.shape == (B, 2, H, W) == flow1.shape and filt0.shape == filt1.shape and
              filt0.shape[0] == B and filt0.shape[2:] == (H, W) and
              (mean or occ0.shape == (B, 1, H, W) == occ1.shape))
        if not ok:
            raise _lib.MemcB200Error("FilterInterpolate: inconsistent shapes " +
                                     " ".join(str(tuple(t.shape)) for t in ts if t is not None))
        fs = int(math.sqrt(float(filt0.size(1))))  # my_lib_cuda.c:619-620
        out = torch.empty_like(ref0)
        S, P = _lib.strides_of, _lib.ptr
        null, s0 = _lib.ctypes.c_void_p(None), _lib.Strides(0, 0, 0)
        _lib.call("memc_b200_filter_interpolation_blend_forward", _lib.stream_ptr(ref0), B, C, H, W, fs,
                  S(ref0), S(flow0), S(filt0), S(ref1), S(flow1), S(filt1), s0 if mean else S(occ0), s0 if mean else S(occ1),
                  S(out), P(ref0), P(flow0), P(filt0), P(ref1), P(flow1), P(filt1), null if mean else P(occ0),
                  null if mean else P(occ1), P(out), _lib.OVERWRITE)
        ctx.mean = mean
        ctx.save_for_backward(*(t for t in ts if t is not None))
        ctx.fs = fs
        return out

    @staticmethod
    def backward(ctx, gout):
        gout = _prep(gout, "gradoutput")
        grads = []
        if ctx.mean:
            ref0, flow0, filt0, ref1, flow1, filt1 = ctx.saved_tensors
            half = (gout * 0.5).contiguous()
            for ref, flow, filt in ((ref0, flow0, filt0), (ref1, flow1, filt1)):
                grads += list(_fi_backward(ref, flow, filt, half, ctx.fs)) + [None]
            return tuple(grads)
        ref0, flow0, filt0, occ0, ref1, flow1, filt1, occ1 = ctx.saved_tensors
        for ref, flow, filt, occ in ((ref0, flow0, filt0, occ0), (ref1, flow1, filt1, occ1)):
            warp = _fi_forward(ref, flow, filt, ctx.fs)                      # recomputed, not stored
            g_occ = (gout * warp).sum(dim=1, keepdim=True)                   # d(occ * warp) / d occ
            g1, g2, g3 = _fi_backward(ref, flow, filt, (gout * occ).contiguous(), ctx.fs)
            grads += [g1, g2, g3, g_occ]
        return tuple(grads)


def FilterInterpolate(ref0, ref2, offset, filter, occlusion, filter_size2=None):  # noqa: A002 (reference's names)
    """Drop-in for the networks' static method (networks/MEMC_Net.py:258-264)."""
    return _FilterInterpolateBlend.apply(ref0, offset[0], filter[0], occlusion[0], ref2, offset[1], filter[1], occlusion[1])


def FilterInterpolateMean(ref0, ref2, offset, filter, filter_size2=None):  # noqa: A002 (reference's names)
    """Drop-in for MEMC_Net_s.FilterInterpolate (networks/MEMC_Net_s.py:258-264): the mean of the two warps."""
    return _FilterInterpolateBlend.apply(ref0, offset[0], filter[0], None, ref2, offset[1], filter[1], None)


class _FilterInterpolateShared(Function):
    @staticmethod
    def forward(ctx, in_a, in_b, flow, filt):
        in_a, in_b, flow, filt = (_prep(t, n) for t, n in zip((in_a, in_b, flow, filt), ("input_a", "input_b", "offset", "filter")))
        _lib.check_same_device(in_a, in_b, flow, filt)
        B, Ca, H, W = in_a.shape
        Cb = in_b.size(1)
        if in_b.shape != (B, Cb, H, W) or flow.shape != (B, 2, H, W) or filt.shape[0] != B or filt.shape[2:] != (H, W):
            raise _lib.MemcB200Error("FilterInterpolateShared: inconsistent shapes")
        fs = int(math.sqrt(float(filt.size(1))))
        out_a, out_b = torch.empty_like(in_a), torch.empty_like(in_b)
        S, P = _lib.strides_of, _lib.ptr
        _lib.call("memc_b200_filter_interpolation_forward_pair", _lib.stream_ptr(in_a), B, Ca, Cb, H, W, fs,
                  S(in_a), S(in_b), S(flow), S(filt), S(out_a), S(out_b), P(in_a), P(in_b), P(flow), P(filt),
                  P(out_a), P(out_b), _lib.OVERWRITE)
        ctx.save_for_backward(in_a, in_b, flow, filt)
        ctx.fs = fs
        return out_a, out_b

    @staticmethod
    def backward(ctx, g_a, g_b):
        in_a, in_b, flow, filt = ctx.saved_tensors
        ga1, ga2, ga3 = _fi_backward(in_a, flow, filt, _prep(g_a, "gradoutput_a"), ctx.fs)
        gb1, gb2, gb3 = _fi_backward(in_b, flow, filt, _prep(g_b, "gradoutput_b"), ctx.fs)
        return ga1, gb1, ga2 + gb2, ga3 + gb3


def FilterInterpolateShared(input_a, input_b, offset, filter):  # noqa: A002 (reference's names)
    """(FilterInterpolationModule()(input_a, offset, filter), FilterInterpolationModule()(input_b, offset, filter)) in one
    pass over the flow / filter planes: the RGB frame and its context features of networks/MEMC_Net_star.py:272-285."""
    return _FilterInterpolateShared.apply(input_a, input_b, offset, filter)


def FlowProjectPair(flow_a, flow_b, requires_grad=None):
    """(FlowProject(flow_a), FlowProject(flow_b)) of networks/MEMC_Net.py:109-113, 252-256 in one call.
    fill-hole follows the reference rule (on iff the inputs do not require grad, FlowProjectionLayer.py:15)."""
    from my_package.functions.FlowProjectionLayer import FlowProjectionLayer
    if flow_a.shape != flow_b.shape:
        raise _lib.MemcB200Error("FlowProjectPair: shapes differ %s %s" % (tuple(flow_a.shape), tuple(flow_b.shape)))
    _lib.check_same_device(flow_a, flow_b)
    if requires_grad is None and flow_a.requires_grad != flow_b.requires_grad:
        # the reference decides fill-hole PER DIRECTION (FlowProjectionModule(input.requires_grad)): one fused call
        # cannot honour two different answers, so the pair is run as the two separate calls it stands for
        return FlowProjectionLayer(flow_a.requires_grad)(flow_a), FlowProjectionLayer(flow_b.requires_grad)(flow_b)
    rg = flow_a.requires_grad if requires_grad is None else requires_grad
    # (the concatenation copies both inputs once: 16 B/px on top of the op's 20 B/px; pass one [2B,2,H,W] tensor to
    # FlowProjectionLayer directly to avoid it)
    both = FlowProjectionLayer(rg)(torch.cat((_lib.check_tensor(flow_a, "flow_a"), _lib.check_tensor(flow_b, "flow_b")), dim=0))
    B = flow_a.size(0)
    return both[:B], both[B:]
