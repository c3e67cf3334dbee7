"""`my_package._ext.my_lib` -- the FFI namespace of the reference, re-hosted on ctypes.

In the reference this module wraps a cffi-built `_my_lib.so` (my_package/_ext/my_lib/
__init__.py:1-12) whose C side is my_package/src/my_lib_cuda.c.  Here every function keeps
the reference's name, argument order, CONTRACT and error convention
(my_package/src/my_lib_cuda.h:19-99):

  * the caller allocates AND zero-fills every output / gradient tensor;
  * return value 0 = ok, -1 = shape / stride violation or launch failure; nothing raises
    for those (the reference's Python prints a non-zero code and carries on);
  * the work is enqueued on the current CUDA stream of the tensors' device, no sync.

Shape / stride checks restate my_lib_cuda.c (line refs per function).  The launch goes to
the reference-named extern "C" launchers of libmemc_b200.so.  CPU tensors are rejected
loudly: this build has no CPU path (the reference's *_cpu_* symbols are not provided).
"""
import math

from memc_b200 import lib as _lib

__all__ = [
    "FilterInterpolationLayer_gpu_forward", "FilterInterpolationLayer_gpu_backward",
    "FlowProjectionLayer_gpu_forward", "FlowProjectionLayer_gpu_backward",
    "DepthFlowProjectionLayer_gpu_forward", "DepthFlowProjectionLayer_gpu_backward",
    "WeightedFlowProjectionLayer_gpu_forward", "WeightedFlowProjectionLayer_gpu_backward",
    "PixelValueLayer_gpu_forward", "PixelValueLayer_gpu_backward",
    "SeparableConvFlowLayer_gpu_forward", "SeparableConvFlowLayer_gpu_backward",
    "WeightLayer_gpu_forward", "WeightLayer_gpu_backward",
    "PixelWeightLayer_gpu_forward", "PixelWeightLayer_gpu_backward",
    "ReliableWeightLayer_gpu_forward", "ReliableWeightLayer_gpu_backward",
    "InterpolationLayer_gpu_forward", "InterpolationLayer_gpu_backward",
    "InterpolationChLayer_gpu_forward", "InterpolationChLayer_gpu_backward",
    "SeparableConvLayer_gpu_forward", "SeparableConvLayer_gpu_backward",
]

_ERR = -1


def _ints(*tensors):
    out = []
    for t in tensors:
        out.extend(int(s) for s in t.stride())
    return out


def _go(name, head, tensors_for_strides, pointers):
    for t in pointers:
        _lib.check_tensor(t, name)
    fn = getattr(_lib.load(), name)
    return int(fn(_lib.stream_ptr(pointers[0]), *head, *_ints(*tensors_for_strides),
                  *[_lib.ptr(t) for t in pointers]))


def _fs_of(input3):
    # my_lib_cuda.c:619-620: filter_size = (int) sqrt((float) size(1))
    return int(math.sqrt(float(input3.size(1))))


# ---------------------------------------------------------------- FilterInterpolation
def FilterInterpolationLayer_gpu_forward(input1, input2, input3, output):
    """my_lib_cuda.c:598-668."""
    batch, channel, h, w = input1.size()
    if input2.size(0) != batch or input2.size(1) != 2:
        return _ERR
    if input2.size(2) != h or input2.size(3) != w:
        return _ERR
    if input1.stride(3) != 1 or input2.stride(3) != 1 or input3.stride(3) != 1:
        return _ERR
    if input1.stride(0) != output.stride(0) or input1.stride(1) != output.stride(1):
        return _ERR
    return _go("FilterInterpolationLayer_gpu_forward_kernel",
               [output.numel(), w, h, channel, batch, _fs_of(input3)],
               [input1, input2, input3], [input1, input2, input3, output])


def FilterInterpolationLayer_gpu_backward(input1, input2, input3, gradoutput,
                                          gradinput1, gradinput2, gradinput3):
    """my_lib_cuda.c:669-749."""
    batch, channel, h, w = input1.size()
    if input2.size(0) != batch or input2.size(1) != 2:
        return _ERR
    if input2.size(2) != h or input2.size(3) != w:
        return _ERR
    if input1.stride(3) != 1 or input2.stride(3) != 1 or input3.stride(3) != 1:
        return _ERR
    if input1.stride(0) != gradinput1.stride(0) or input2.stride(0) != gradinput2.stride(0):
        return _ERR
    if input1.stride(1) != gradinput1.stride(1) or input2.stride(1) != gradinput2.stride(1):
        return _ERR
    if input3.stride(1) != gradinput3.stride(1):
        return _ERR
    return _go("FilterInterpolationLayer_gpu_backward_kernel",
               [gradoutput.numel(), w, h, channel, batch, _fs_of(input3)],
               [input1, input2, input3],
               [input1, input2, input3, gradoutput, gradinput1, gradinput2, gradinput3])


# --------------------------------------------------------------------- FlowProjection
def FlowProjectionLayer_gpu_forward(input1, count, output, fillhole):
    """my_lib_cuda.c:752-799."""
    batch, channel, h, w = input1.size()
    if channel != 2:
        return _ERR
    if input1.stride(0) != output.stride(0) or input1.stride(1) != output.stride(1):
        return _ERR
    return _go("FlowProjection_gpu_forward_kernel",
               [output.numel(), w, h, channel, batch, int(fillhole)],
               [input1, count], [input1, count, output])


def FlowProjectionLayer_gpu_backward(input1, count, gradoutput, gradinput1):
    """my_lib_cuda.c:801-855."""
    batch, channel, h, w = input1.size()
    if channel != 2:
        return _ERR
    if count.size(0) != batch or count.size(1) != 1:
        return _ERR
    if count.size(2) != h or count.size(3) != w:
        return _ERR
    if input1.stride(0) != gradinput1.stride(0) or input1.stride(1) != gradinput1.stride(1):
        return _ERR
    return _go("FlowProjection_gpu_backward_kernel",
               [gradoutput.numel(), w, h, channel, batch],
               [input1, count], [input1, count, gradoutput, gradinput1])


# ---------------------------------------------------------------- DepthFlowProjection
def DepthFlowProjectionLayer_gpu_forward(input1, input2, count, output, fillhole):
    """my_lib_cuda.c:857-914 (declared my_lib_cuda.h:101-107)."""
    batch, channel, h, w = input1.size()
    if channel != 2:
        return _ERR
    if input2.size(1) != 1:
        return _ERR
    if input1.stride(0) != output.stride(0) or input1.stride(1) != output.stride(1):
        return _ERR
    return _go("DepthFlowProjection_gpu_forward_kernel",
               [output.numel(), w, h, channel, batch, int(fillhole)],
               [input1, input2, count], [input1, input2, count, output])


def DepthFlowProjectionLayer_gpu_backward(input1, input2, count, output, gradoutput, gradinput1, gradinput2):
    """my_lib_cuda.c:916-985 (declared my_lib_cuda.h:109-117)."""
    batch, channel, h, w = input1.size()
    if channel != 2:
        return _ERR
    if input2.size(1) != 1:
        return _ERR
    if count.size(0) != batch or count.size(1) != 1:
        return _ERR
    if count.size(2) != h or count.size(3) != w:
        return _ERR
    if input1.stride(0) != gradinput1.stride(0) or input1.stride(1) != gradinput1.stride(1):
        return _ERR
    return _go("DepthFlowProjection_gpu_backward_kernel",
               [gradoutput.numel(), w, h, channel, batch],
               [input1, input2, count],
               [input1, input2, count, output, gradoutput, gradinput1, gradinput2])


# ------------------------------------------------------------- WeightedFlowProjection
def _go_thr(name, head, threshold, tensors_for_strides, pointers):
    import ctypes
    for t in pointers:
        _lib.check_tensor(t, name)
    fn = getattr(_lib.load(), name)
    return int(fn(_lib.stream_ptr(pointers[0]), *head, ctypes.c_float(threshold), *_ints(*tensors_for_strides),
                  *[_lib.ptr(t) for t in pointers]))


def WeightedFlowProjectionLayer_gpu_forward(input1, input2, input3, count, weight, output, fillhole, threshhold):
    """my_lib_cuda.c:986-1060 (declared my_lib_cuda.h:119-128)."""
    batch, channel, h, w = input1.size()
    if channel != 2:
        return _ERR
    if input2.size(1) != 3 or input3.size(1) != 3:
        return _ERR
    if input1.stride(0) != output.stride(0) or input1.stride(1) != output.stride(1):
        return _ERR
    return _go_thr("WeightedFlowProjection_gpu_forward_kernel",
                   [output.numel(), w, h, channel, batch, int(fillhole)], float(threshhold),
                   [input1, input2, input3, count, weight], [input1, input2, input3, count, weight, output])


def WeightedFlowProjectionLayer_gpu_backward(input1, input2, input3, count, weight, gradoutput, gradinput1, threshhold):
    """my_lib_cuda.c:1062-1140 (declared my_lib_cuda.h:130-139)."""
    batch, channel, h, w = input1.size()
    if channel != 2:
        return _ERR
    if input2.size(1) != 3 or input3.size(1) != 3:
        return _ERR
    if count.size(0) != batch or count.size(1) != 1:
        return _ERR
    if count.size(2) != h or count.size(3) != w:
        return _ERR
    if input1.stride(0) != gradinput1.stride(0) or input1.stride(1) != gradinput1.stride(1):
        return _ERR
    return _go_thr("WeightedFlowProjection_gpu_backward_kernel",
                   [gradoutput.numel(), w, h, channel, batch], float(threshhold),
                   [input1, input2, input3, count, weight],
                   [input1, input2, input3, count, weight, gradoutput, gradinput1])


# ------------------------------------- PixelValue / PixelWeight / ReliableWeight (4x4 splat family)
def _go_tail(name, head, tensors_for_strides, pointers, floats):
    import ctypes
    for t in pointers:
        _lib.check_tensor(t, name)
    fn = getattr(_lib.load(), name)
    return int(fn(_lib.stream_ptr(pointers[0]), *head, *_ints(*tensors_for_strides), *[_lib.ptr(t) for t in pointers],
                  *[ctypes.c_float(v) for v in floats]))


def PixelValueLayer_gpu_forward(input1, input3, flow_weights, output, sigma_d, tao_r, Prowindow):
    """my_lib_cuda.c:1335-1433 (declared my_lib_cuda.h:163-166)."""
    if Prowindow != 2.0:
        return _ERR
    batch, channel, h, w = input1.size()
    if channel != 3:
        return _ERR
    if input3.size(1) != 2 or flow_weights.size(1) != 1 or output.size(1) != 3:
        return _ERR
    if input1.stride(3) != 1 or input3.stride(3) != 1 or flow_weights.stride(3) != 1:
        return _ERR
    if input1.stride(0) != output.stride(0) or input1.stride(1) != output.stride(1):
        return _ERR
    return _go_tail("PixelValueLayer_gpu_forward_kernel", [output.numel(), w, h, channel, batch],
                    [input1, input3, flow_weights, output], [input1, input3, flow_weights, output],
                    [sigma_d, tao_r, Prowindow])


def PixelValueLayer_gpu_backward(input1, input3, flow_weights, gradoutput, gradinput1, gradinput3, gradflow_weights,
                                 sigma_d, tao_r, Prowindow):
    """my_lib_cuda.c:1435-1538 (declared my_lib_cuda.h:167-172)."""
    if Prowindow != 2.0:
        return _ERR
    batch, channel, h, w = input1.size()
    if channel != 3:
        return _ERR
    if input3.size(1) != 2 or flow_weights.size(1) != 1:
        return _ERR
    if input1.stride(3) != 1 or input3.stride(3) != 1:
        return _ERR
    if input1.stride(0) != gradinput1.stride(0) or input1.stride(1) != gradinput1.stride(1):
        return _ERR
    if input3.stride(1) != gradinput3.stride(1) or flow_weights.stride(0) != gradflow_weights.stride(0):
        return _ERR
    return _go_tail("PixelValueLayer_gpu_backward_kernel", [gradoutput.numel(), w, h, channel, batch],
                    [input1, input3, flow_weights, gradoutput],
                    [input1, input3, flow_weights, gradoutput, gradinput1, gradinput3, gradflow_weights],
                    [sigma_d, tao_r, Prowindow])


def PixelWeightLayer_gpu_forward(input3, flow_weights, output, sigma_d, tao_r, Prowindow):
    """my_lib_cuda.c:1540-1624 (declared my_lib_cuda.h:173-176)."""
    if Prowindow != 2.0:
        return _ERR
    batch, channel, h, w = input3.size()
    if channel != 2 or flow_weights.size(1) != 1 or output.size(1) != 1:
        return _ERR
    if input3.stride(3) != 1 or flow_weights.stride(3) != 1:
        return _ERR
    return _go_tail("PixelWeightLayer_gpu_forward_kernel", [output.numel(), w, h, batch],
                    [input3, flow_weights, output], [input3, flow_weights, output], [sigma_d, tao_r, Prowindow])


def PixelWeightLayer_gpu_backward(input3, flow_weights, output, gradoutput, gradinput3, gradflow_weights,
                                  threshhold, sigma_d, tao_r, Prowindow):
    """my_lib_cuda.c:1626-1726 (declared my_lib_cuda.h:177-183)."""
    if Prowindow != 2.0:
        return _ERR
    batch, channel, h, w = input3.size()
    if channel != 2 or flow_weights.size(1) != 1:
        return _ERR
    if input3.stride(3) != 1 or input3.stride(1) != gradinput3.stride(1):
        return _ERR
    return _go_tail("PixelWeightLayer_gpu_backward_kernel", [gradoutput.numel(), w, h, batch],
                    [input3, flow_weights, output],
                    [input3, flow_weights, output, gradoutput, gradinput3, gradflow_weights],
                    [threshhold, sigma_d, tao_r, Prowindow])


def ReliableWeightLayer_gpu_forward(input3, output, sigma_d, tao_r, Prowindow):
    """my_lib_cuda.c:1744-1829 (declared my_lib_cuda.h:194-197)."""
    if Prowindow != 2.0:
        return _ERR
    batch, channel, h, w = input3.size()
    if channel != 2 or output.size(1) != 1:
        return _ERR
    if input3.stride(3) != 1:
        return _ERR
    return _go_tail("ReliableWeightLayer_gpu_forward_kernel", [output.numel(), w, h, batch],
                    [input3, output], [input3, output], [sigma_d, tao_r, Prowindow])


def ReliableWeightLayer_gpu_backward(input3, output, gradoutput, gradinput3, threshhold, sigma_d, tao_r, Prowindow):
    """my_lib_cuda.c:1831-1932 (declared my_lib_cuda.h:198-203)."""
    if Prowindow != 2.0:
        return _ERR
    batch, channel, h, w = input3.size()
    if channel != 2:
        return _ERR
    if input3.stride(3) != 1 or input3.stride(1) != gradinput3.stride(1):
        return _ERR
    return _go_tail("ReliableWeightLayer_gpu_backward_kernel", [gradoutput.numel(), w, h, batch],
                    [input3, output], [input3, output, gradoutput, gradinput3],
                    [threshhold, sigma_d, tao_r, Prowindow])


# ------------------------------------------------------------------------ WeightLayer
def _wl_checks(input1, input2, input3, output, Nw):
    # my_lib_cuda.c:1160-1215
    if Nw != 3.0:
        return None
    batch, channel, h, w = input1.size()
    if channel != 3 or input2.size(0) != batch or input2.size(2) != h or input2.size(3) != w:
        return None
    if input3.size(1) != 2 or output.size(1) != 1:
        return None
    if input1.stride(3) != 1 or input2.stride(3) != 1 or input3.stride(3) != 1:
        return None
    if input1.stride(0) != input2.stride(0):
        return None
    return batch, channel, h, w


def WeightLayer_gpu_forward(input1, input2, input3, output, lambda_e, lambda_v, Nw):
    """my_lib_cuda.c:1142-1238 (declared my_lib_cuda.h:147-152)."""
    d = _wl_checks(input1, input2, input3, output, Nw)
    if d is None:
        return _ERR
    batch, channel, h, w = d
    return _go_tail("WeightLayer_gpu_forward_kernel", [output.numel(), w, h, channel, batch],
                    [input1, input2, input3, output], [input1, input2, input3, output], [lambda_e, lambda_v, Nw])


def WeightLayer_gpu_backward(input1, input2, input3, output, gradoutput, gradinput1, gradinput2, gradinput3,
                             lambda_e, lambda_v, Nw):
    """my_lib_cuda.c:1240-1333 (declared my_lib_cuda.h:153-160)."""
    d = _wl_checks(input1, input2, input3, output, Nw)
    if d is None:
        return _ERR
    batch, channel, h, w = d
    if input1.stride(0) != gradinput1.stride(0) or input1.stride(1) != gradinput1.stride(1):
        return _ERR
    if input2.stride(0) != gradinput2.stride(0) or input2.stride(1) != gradinput2.stride(1):
        return _ERR
    if input3.stride(0) != gradinput3.stride(0) or input3.stride(1) != gradinput3.stride(1):
        return _ERR
    return _go_tail("WeightLayer_gpu_backward_kernel", [gradoutput.numel(), w, h, channel, batch],
                    [input1, input2, input3, output],
                    [input1, input2, input3, output, gradoutput, gradinput1, gradinput2, gradinput3],
                    [lambda_e, lambda_v, Nw])


# ------------------------------------------------------------------ SeparableConvFlow
def _scf_checks(input1, input2, input3, flow_output):
    # my_lib_cuda.c:24-34, 65-73
    batch, channel, h, w = input1.size()
    fs = input2.size(1)
    if channel != 3 or input2.size(0) != batch:
        return None
    if input2.size(2) != h - fs + 1 or input2.size(3) != w - fs + 1:
        return None
    if input1.stride(3) != 1 or input2.stride(3) != 1 or input3.stride(3) != 1 or flow_output.stride(3) != 1:
        return None
    if input2.stride(0) != input3.stride(0) or input2.stride(1) != input3.stride(1):
        return None
    return batch, channel, h, w, fs


def SeparableConvFlowLayer_gpu_forward(input1, input2, input3, flow_output):
    """my_lib_cuda.c:12-102 (declared my_lib_cuda.h:2-8)."""
    d = _scf_checks(input1, input2, input3, flow_output)
    if d is None:
        return _ERR
    batch, channel, h, w, fs = d
    return _go("SeparableConvFlowLayer_gpu_forward_kernel", [flow_output.numel(), w, h, channel, batch, fs],
               [input1, input2, input3, flow_output], [input1, input2, input3, flow_output])


def SeparableConvFlowLayer_gpu_backward(input1, input2, input3, gradflow_output, gradinput1, gradinput2, gradinput3):
    """my_lib_cuda.c:104-198 (declared my_lib_cuda.h:10-17)."""
    d = _scf_checks(input1, input2, input3, gradflow_output)
    if d is None:
        return _ERR
    batch, channel, h, w, fs = d
    if input2.stride(0) != gradinput2.stride(0) or input2.stride(1) != gradinput2.stride(1):
        return _ERR
    if input3.stride(0) != gradinput3.stride(0) or input3.stride(1) != gradinput3.stride(1):
        return _ERR
    return _go("SeparableConvFlowLayer_gpu_backward_kernel", [gradflow_output.numel(), w, h, channel, batch, fs],
               [input1, input2, input3, gradflow_output],
               [input1, input2, input3, gradflow_output, gradinput1, gradinput2, gradinput3])


# ---------------------------------------------------------------------- Interpolation
def _interp_fwd(name, need3, input1, input2, output):
    batch, channel, h, w = input1.size()
    if need3 and channel != 3:  # my_lib_cuda.c:373 (dropped in the Ch variant, :490)
        return _ERR
    if input2.size(0) != batch or input2.size(1) != 2:
        return _ERR
    if input2.size(2) != h or input2.size(3) != w:
        return _ERR
    if input1.stride(0) != output.stride(0) or input1.stride(1) != output.stride(1):
        return _ERR
    return _go(name, [output.numel(), w, h, channel, batch], [input1, input2], [input1, input2, output])


def _interp_bwd(name, need3, input1, input2, gradoutput, gradinput1, gradinput2):
    batch, channel, h, w = input1.size()
    if need3 and channel != 3:
        return _ERR
    if input2.size(0) != batch or input2.size(1) != 2:
        return _ERR
    if input2.size(2) != h or input2.size(3) != w:
        return _ERR
    if input1.stride(0) != gradinput1.stride(0) or input2.stride(0) != gradinput2.stride(0):
        return _ERR
    if input1.stride(1) != gradinput1.stride(1) or input2.stride(1) != gradinput2.stride(1):
        return _ERR
    return _go(name, [gradoutput.numel(), w, h, channel, batch], [input1, input2],
               [input1, input2, gradoutput, gradinput1, gradinput2])


def InterpolationLayer_gpu_forward(input1, input2, output):
    """my_lib_cuda.c:364-416."""
    return _interp_fwd("InterpolationLayer_gpu_forward_kernel", True, input1, input2, output)


def InterpolationLayer_gpu_backward(input1, input2, gradoutput, gradinput1, gradinput2):
    """my_lib_cuda.c:419-479."""
    return _interp_bwd("InterpolationLayer_gpu_backward_kernel", True, input1, input2, gradoutput,
                       gradinput1, gradinput2)


def InterpolationChLayer_gpu_forward(input1, input2, output):
    """my_lib_cuda.c:481-533."""
    return _interp_fwd("InterpolationChLayer_gpu_forward_kernel", False, input1, input2, output)


def InterpolationChLayer_gpu_backward(input1, input2, gradoutput, gradinput1, gradinput2):
    """my_lib_cuda.c:536-596."""
    return _interp_bwd("InterpolationChLayer_gpu_backward_kernel", False, input1, input2, gradoutput,
                       gradinput1, gradinput2)


# ---------------------------------------------------------------------- SeparableConv
def _sepconv_checks(input1, input2, input3):
    batch, channel, h, w = input1.size()
    if channel != 3:  # my_lib_cuda.c:211
        return False
    if input2.size(0) != batch:
        return False
    fs = input2.size(1)
    if input2.size(2) != h - fs + 1 or input2.size(3) != w - fs + 1:  # :218-219
        return False
    return True


def SeparableConvLayer_gpu_forward(input1, input2, input3, output):
    """my_lib_cuda.c:200-276."""
    if not _sepconv_checks(input1, input2, input3):
        return _ERR
    if any(t.stride(3) != 1 for t in (input1, input2, input3, output)):
        return _ERR
    if input2.stride(0) != input3.stride(0) or input2.stride(1) != input3.stride(1):
        return _ERR
    batch, channel, h, w = input1.size()
    return _go("SeparableConvLayer_gpu_forward_kernel",
               [output.numel(), w, h, channel, batch, input2.size(1)],
               [input1, input2, input3, output], [input1, input2, input3, output])


def SeparableConvLayer_gpu_backward(input1, input2, input3, gradoutput,
                                    gradinput1, gradinput2, gradinput3):
    """my_lib_cuda.c:277-362."""
    if not _sepconv_checks(input1, input2, input3):
        return _ERR
    if any(t.stride(3) != 1 for t in (input1, input2, input3, gradoutput)):
        return _ERR
    if input1.stride(0) != gradinput1.stride(0) or input2.stride(0) != gradinput2.stride(0):
        return _ERR
    if input1.stride(1) != gradinput1.stride(1) or input2.stride(1) != gradinput2.stride(1):
        return _ERR
    if input3.stride(1) != gradinput3.stride(1):
        return _ERR
    batch, channel, h, w = input1.size()
    return _go("SeparableConvLayer_gpu_backward_kernel",
               [gradoutput.numel(), w, h, channel, batch, input2.size(1)],
               [input1, input2, input3, gradoutput],
               [input1, input2, input3, gradoutput, gradinput1, gradinput2, gradinput3])
