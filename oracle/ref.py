"""ctypes bindings of oracle/_ref/*.so = the REFERENCE'S OWN sources, compiled unchanged
(TEST INFRASTRUCTURE ONLY, see oracle/__init__.py and oracle/Makefile).

  cpu_*   -> libmemc_ref_cpu.so  (my_package/src/my_lib.c against oracle/th_stub/TH.h);
             takes/returns numpy arrays; runs anywhere
  gpu_*   -> libmemc_ref_gpu.so  (my_package/src/my_lib_kernel.cu for sm_100a) through its
             extern "C" launchers (my_lib_kernel.h); takes torch CUDA tensors, launches on
             the current torch stream.  This is both the GPU parity oracle (<=1e-5) and
             the "legacy kernels recompiled on the same box" baseline.

The libraries are built HERE (where /root/reference exists) and travel to the GPU box as
binaries; `available_cpu()/available_gpu()` say whether they are present.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(_HERE, "_ref")
_CPU_SO = os.path.join(_REF_DIR, "libmemc_ref_cpu.so")
_GPU_SO = os.path.join(_REF_DIR, "libmemc_ref_gpu.so")
_cache = {}


def build(reference_root="/root/reference"):
    """(Re)build oracle/_ref from the reference sources if they are present."""
    if os.path.isdir(os.path.join(reference_root, "my_package", "src")):
        subprocess.run(["make", "-s", "-C", _HERE, "ref", "REF=%s" % reference_root], check=True)
    return available_cpu(), available_gpu()


def available_cpu():
    return os.path.exists(_CPU_SO)


def available_gpu():
    return os.path.exists(_GPU_SO)


# ------------------------------------------------------------------------------ CPU side
class _THFloatTensor(ctypes.Structure):
    # must match oracle/th_stub/TH.h
    _fields_ = [("size", ctypes.POINTER(ctypes.c_long)),
                ("stride", ctypes.POINTER(ctypes.c_long)),
                ("nDimension", ctypes.c_int),
                ("data", ctypes.POINTER(ctypes.c_float))]


class _TH:
    """Keeps the numpy array and the size/stride vectors alive next to the struct."""

    def __init__(self, arr):
        assert arr.dtype == np.float32
        self.arr = arr
        n = arr.ndim
        self.size = (ctypes.c_long * n)(*arr.shape)
        self.stride = (ctypes.c_long * n)(*[s // 4 for s in arr.strides])
        self.t = _THFloatTensor(self.size, self.stride, n,
                                arr.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))

    @property
    def ref(self):
        return ctypes.byref(self.t)


def _cpu():
    if "cpu" not in _cache:
        if not available_cpu():
            raise FileNotFoundError(_CPU_SO + " (run `make -C oracle ref` where /root/reference exists)")
        _cache["cpu"] = ctypes.CDLL(_CPU_SO)
    return _cache["cpu"]


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def cpu_filter_interpolation_forward(in1, flow, filt):
    in1, flow, filt = _c(in1), _c(flow), _c(filt)
    out = np.zeros_like(in1)
    rc = _cpu().FilterInterpolationLayer_cpu_forward(_TH(in1).ref, _TH(flow).ref, _TH(filt).ref, _TH(out).ref)
    assert rc == 0, rc
    return out


def cpu_filter_interpolation_backward(in1, flow, filt, gout):
    in1, flow, filt, gout = _c(in1), _c(flow), _c(filt), _c(gout)
    g1, g2, g3 = np.zeros_like(in1), np.zeros_like(flow), np.zeros_like(filt)
    rc = _cpu().FilterInterpolationLayer_cpu_backward(
        _TH(in1).ref, _TH(flow).ref, _TH(filt).ref, _TH(gout).ref, _TH(g1).ref, _TH(g2).ref, _TH(g3).ref)
    assert rc == 0, rc
    return g1, g2, g3


def cpu_flow_projection_forward(flow):
    """scatter + average only: the reference CPU path has no fill-hole (my_lib.c:1539-1543)."""
    flow = _c(flow)
    B, _, H, W = flow.shape
    count = np.zeros((B, 1, H, W), np.float32)
    out = np.zeros_like(flow)
    rc = _cpu().FlowProjectionLayer_cpu_forward(_TH(flow).ref, _TH(count).ref, _TH(out).ref, ctypes.c_int(0))
    assert rc == 0, rc
    return out, count


def cpu_flow_projection_backward(flow, count, gout):
    flow, count, gout = _c(flow), _c(count), _c(gout)
    gi = np.zeros_like(flow)
    rc = _cpu().FlowProjectionLayer_cpu_backward(_TH(flow).ref, _TH(count).ref, _TH(gout).ref, _TH(gi).ref)
    assert rc == 0, rc
    return gi


def cpu_depth_flow_projection_forward(flow, depth):
    """scatter + average only: the reference CPU path has no fill-hole (my_lib.c:1738-1740)."""
    flow, depth = _c(flow), _c(depth)
    B, _, H, W = flow.shape
    count = np.zeros((B, 1, H, W), np.float32)
    out = np.zeros_like(flow)
    rc = _cpu().DepthFlowProjectionLayer_cpu_forward(_TH(flow).ref, _TH(depth).ref, _TH(count).ref, _TH(out).ref,
                                                     ctypes.c_int(0))
    assert rc == 0, rc
    return out, count


def cpu_depth_flow_projection_backward(flow, depth, count, fout, gout):
    flow, depth, count, fout, gout = _c(flow), _c(depth), _c(count), _c(fout), _c(gout)
    g1, g2 = np.zeros_like(flow), np.zeros_like(depth)
    rc = _cpu().DepthFlowProjectionLayer_cpu_backward(_TH(flow).ref, _TH(depth).ref, _TH(count).ref, _TH(fout).ref,
                                                      _TH(gout).ref, _TH(g1).ref, _TH(g2).ref)
    assert rc == 0, rc
    return g1, g2


def cpu_weighted_flow_projection_forward(flow, im0, im1, threshold):
    """scatter + average only (no fill-hole in the reference's CPU path, my_lib.c:2012-2014)."""
    flow, im0, im1 = _c(flow), _c(im0), _c(im1)
    B, _, H, W = flow.shape
    count, weight = np.zeros((B, 1, H, W), np.float32), np.zeros((B, 1, H, W), np.float32)
    out = np.zeros_like(flow)
    rc = _cpu().WeightedFlowProjectionLayer_cpu_forward(_TH(flow).ref, _TH(im0).ref, _TH(im1).ref, _TH(count).ref,
                                                        _TH(weight).ref, _TH(out).ref, ctypes.c_int(0),
                                                        ctypes.c_float(threshold))
    assert rc == 0, rc
    return out, count, weight


def cpu_weighted_flow_projection_backward(flow, im0, im1, count, weight, gout, threshold):
    flow, im0, im1, count, weight, gout = _c(flow), _c(im0), _c(im1), _c(count), _c(weight), _c(gout)
    gi = np.zeros_like(flow)
    rc = _cpu().WeightedFlowProjectionLayer_cpu_backward(_TH(flow).ref, _TH(im0).ref, _TH(im1).ref, _TH(count).ref,
                                                         _TH(weight).ref, _TH(gout).ref, _TH(gi).ref,
                                                         ctypes.c_float(threshold))
    assert rc == 0, rc
    return gi


def _f(v):
    return ctypes.c_float(v)


def cpu_pixel_splat_forward(mode, flow, in1=None, fw=None, sigma_d=1.0):
    """PixelValueLayer / PixelWeightLayer / ReliableWeightLayer _cpu_forward (my_lib.c:2615-3400); tao_r = 0, Prowindow = 2."""
    flow = _c(flow)
    B, _, H, W = flow.shape
    if mode == "value":
        in1, fw = _c(in1), _c(fw)
        out = np.zeros_like(in1)
        rc = _cpu().PixelValueLayer_cpu_forward(_TH(in1).ref, _TH(flow).ref, _TH(fw).ref, _TH(out).ref, _f(sigma_d), _f(0), _f(2))
    elif mode == "weight":
        fw = _c(fw)
        out = np.zeros((B, 1, H, W), np.float32)
        rc = _cpu().PixelWeightLayer_cpu_forward(_TH(flow).ref, _TH(fw).ref, _TH(out).ref, _f(sigma_d), _f(0), _f(2))
    else:
        out = np.zeros((B, 1, H, W), np.float32)
        rc = _cpu().ReliableWeightLayer_cpu_forward(_TH(flow).ref, _TH(out).ref, _f(sigma_d), _f(0), _f(2))
    assert rc == 0, rc
    return out


def cpu_pixel_splat_backward(mode, flow, gout, in1=None, fw=None, fout=None, sigma_d=1.0, threshold=0.0):
    flow, gout = _c(flow), _c(gout)
    B, _, H, W = flow.shape
    g3 = np.zeros_like(flow)
    if mode == "value":
        in1, fw = _c(in1), _c(fw)
        g1, gw = np.zeros_like(in1), np.zeros_like(fw)
        rc = _cpu().PixelValueLayer_cpu_backward(_TH(in1).ref, _TH(flow).ref, _TH(fw).ref, _TH(gout).ref, _TH(g1).ref,
                                                 _TH(g3).ref, _TH(gw).ref, _f(sigma_d), _f(0), _f(2))
        assert rc == 0, rc
        return g1, g3, gw
    fout = _c(fout)
    if mode == "weight":
        fw = _c(fw)
        gw = np.zeros_like(fw)
        rc = _cpu().PixelWeightLayer_cpu_backward(_TH(flow).ref, _TH(fw).ref, _TH(fout).ref, _TH(gout).ref, _TH(g3).ref,
                                                  _TH(gw).ref, _f(threshold), _f(sigma_d), _f(0), _f(2))
        assert rc == 0, rc
        return None, g3, gw
    rc = _cpu().ReliableWeightLayer_cpu_backward(_TH(flow).ref, _TH(fout).ref, _TH(gout).ref, _TH(g3).ref, _f(threshold),
                                                 _f(sigma_d), _f(0), _f(2))
    assert rc == 0, rc
    return None, g3, None


def cpu_weight_layer_forward(in1, in2, flow, lambda_e):
    in1, in2, flow = _c(in1), _c(in2), _c(flow)
    out = np.zeros((in1.shape[0], 1) + in1.shape[2:], np.float32)
    rc = _cpu().WeightLayer_cpu_forward(_TH(in1).ref, _TH(in2).ref, _TH(flow).ref, _TH(out).ref, _f(lambda_e), _f(0), _f(3))
    assert rc == 0, rc
    return out


def cpu_weight_layer_backward(in1, in2, flow, fout, gout, lambda_e):
    in1, in2, flow, fout, gout = _c(in1), _c(in2), _c(flow), _c(fout), _c(gout)
    g1, g2, g3 = np.zeros_like(in1), np.zeros_like(in2), np.zeros_like(flow)
    rc = _cpu().WeightLayer_cpu_backward(_TH(in1).ref, _TH(in2).ref, _TH(flow).ref, _TH(fout).ref, _TH(gout).ref,
                                         _TH(g1).ref, _TH(g2).ref, _TH(g3).ref, _f(lambda_e), _f(0), _f(3))
    assert rc == 0, rc
    return g1, g2, g3


def cpu_separable_conv_flow_forward(in1, vert, horiz):
    in1, vert, horiz = _c(in1), _c(vert), _c(horiz)
    flow = np.zeros((vert.shape[0], 2) + vert.shape[2:], np.float32)
    rc = _cpu().SeparableConvFlowLayer_cpu_forward(_TH(in1).ref, _TH(vert).ref, _TH(horiz).ref, _TH(flow).ref)
    assert rc == 0, rc
    return flow


def cpu_separable_conv_flow_backward(in1, vert, horiz, gflow):
    in1, vert, horiz, gflow = _c(in1), _c(vert), _c(horiz), _c(gflow)
    g1, gv, gh = np.zeros_like(in1), np.zeros_like(vert), np.zeros_like(horiz)
    rc = _cpu().SeparableConvFlowLayer_cpu_backward(_TH(in1).ref, _TH(vert).ref, _TH(horiz).ref, _TH(gflow).ref,
                                                    _TH(g1).ref, _TH(gv).ref, _TH(gh).ref)
    assert rc == 0, rc
    return gv, gh


def cpu_interpolation_forward(in1, flow):
    in1, flow = _c(in1), _c(flow)
    out = np.zeros_like(in1)
    fn = _cpu().InterpolationLayer_cpu_forward if in1.shape[1] == 3 else _cpu().InterpolationChLayer_cpu_forward
    rc = fn(_TH(in1).ref, _TH(flow).ref, _TH(out).ref)
    assert rc == 0, rc
    return out


def cpu_interpolation_backward(in1, flow, gout):
    in1, flow, gout = _c(in1), _c(flow), _c(gout)
    g1, g2 = np.zeros_like(in1), np.zeros_like(flow)
    fn = _cpu().InterpolationLayer_cpu_backward if in1.shape[1] == 3 else _cpu().InterpolationChLayer_cpu_backward
    rc = fn(_TH(in1).ref, _TH(flow).ref, _TH(gout).ref, _TH(g1).ref, _TH(g2).ref)
    assert rc == 0, rc
    return g1, g2


def cpu_separable_conv_forward(in1, vert, horiz):
    in1, vert, horiz = _c(in1), _c(vert), _c(horiz)
    B, C, H, W = in1.shape
    fs = vert.shape[1]
    out = np.zeros((B, C, H - fs + 1, W - fs + 1), np.float32)
    rc = _cpu().SeparableConvLayer_cpu_forward(_TH(in1).ref, _TH(vert).ref, _TH(horiz).ref, _TH(out).ref)
    assert rc == 0, rc
    return out


def cpu_separable_conv_backward(in1, vert, horiz, gout):
    in1, vert, horiz, gout = _c(in1), _c(vert), _c(horiz), _c(gout)
    g1, g2, g3 = np.zeros_like(in1), np.zeros_like(vert), np.zeros_like(horiz)
    rc = _cpu().SeparableConvLayer_cpu_backward(
        _TH(in1).ref, _TH(vert).ref, _TH(horiz).ref, _TH(gout).ref, _TH(g1).ref, _TH(g2).ref, _TH(g3).ref)
    assert rc == 0, rc
    return g1, g2, g3


# ------------------------------------------------------------------------------ GPU side
def _gpu():
    if "gpu" not in _cache:
        if not available_gpu():
            raise FileNotFoundError(_GPU_SO + " (run `make -C oracle ref` where /root/reference exists)")
        _cache["gpu"] = ctypes.CDLL(_GPU_SO)
    return _cache["gpu"]


def _stream():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _s(t):
    return [ctypes.c_int(int(x)) for x in t.stride()]


def _d(t):
    return ctypes.c_void_p(t.data_ptr())


def _i(x):
    return ctypes.c_int(int(x))


def gpu_filter_interpolation_forward(in1, flow, filt, out=None):
    import torch
    B, C, H, W = in1.shape
    fs = int(np.sqrt(np.float32(filt.shape[1])))
    if out is None:
        out = torch.zeros_like(in1)
    rc = _gpu().FilterInterpolationLayer_gpu_forward_kernel(
        _stream(), _i(out.numel()), _i(W), _i(H), _i(C), _i(B), _i(fs),
        *_s(in1), *_s(flow), *_s(filt), _d(in1), _d(flow), _d(filt), _d(out))
    assert rc == 0, rc
    return out


def gpu_filter_interpolation_backward(in1, flow, filt, gout, grads=None):
    import torch
    B, C, H, W = in1.shape
    fs = int(np.sqrt(np.float32(filt.shape[1])))
    if grads is None:
        grads = (torch.zeros_like(in1), torch.zeros_like(flow), torch.zeros_like(filt))
    g1, g2, g3 = grads
    rc = _gpu().FilterInterpolationLayer_gpu_backward_kernel(
        _stream(), _i(gout.numel()), _i(W), _i(H), _i(C), _i(B), _i(fs),
        *_s(in1), *_s(flow), *_s(filt), _d(in1), _d(flow), _d(filt), _d(gout), _d(g1), _d(g2), _d(g3))
    assert rc == 0, rc
    return g1, g2, g3


def gpu_flow_projection_forward(flow, fillhole, bufs=None):
    import torch
    B, _, H, W = flow.shape
    if bufs is None:
        bufs = (torch.zeros(B, 1, H, W, device=flow.device), torch.zeros_like(flow))
    count, out = bufs
    rc = _gpu().FlowProjection_gpu_forward_kernel(
        _stream(), _i(out.numel()), _i(W), _i(H), _i(2), _i(B), _i(fillhole),
        *_s(flow), *_s(count), _d(flow), _d(count), _d(out))
    assert rc == 0, rc
    return out, count


def gpu_flow_projection_backward(flow, count, gout, gi=None):
    import torch
    B, _, H, W = flow.shape
    if gi is None:
        gi = torch.zeros_like(flow)
    rc = _gpu().FlowProjection_gpu_backward_kernel(
        _stream(), _i(gout.numel()), _i(W), _i(H), _i(2), _i(B),
        *_s(flow), *_s(count), _d(flow), _d(count), _d(gout), _d(gi))
    assert rc == 0, rc
    return gi


def gpu_depth_flow_projection_forward(flow, depth, fillhole, bufs=None):
    import torch
    B, _, H, W = flow.shape
    if bufs is None:
        bufs = (torch.zeros(B, 1, H, W, device=flow.device), torch.zeros_like(flow))
    count, out = bufs
    rc = _gpu().DepthFlowProjection_gpu_forward_kernel(
        _stream(), _i(out.numel()), _i(W), _i(H), _i(2), _i(B), _i(fillhole),
        *_s(flow), *_s(depth), *_s(count), _d(flow), _d(depth), _d(count), _d(out))
    assert rc == 0, rc
    return out, count


def gpu_depth_flow_projection_backward(flow, depth, count, fout, gout, grads=None):
    import torch
    B, _, H, W = flow.shape
    g1, g2 = grads if grads is not None else (torch.zeros_like(flow), torch.zeros_like(depth))
    rc = _gpu().DepthFlowProjection_gpu_backward_kernel(
        _stream(), _i(gout.numel()), _i(W), _i(H), _i(2), _i(B),
        *_s(flow), *_s(depth), *_s(count), _d(flow), _d(depth), _d(count), _d(fout), _d(gout), _d(g1), _d(g2))
    assert rc == 0, rc
    return g1, g2


def gpu_weighted_flow_projection_forward(flow, im0, im1, fillhole, threshold, bufs=None):
    import torch
    B, _, H, W = flow.shape
    if bufs is None:
        bufs = (torch.zeros(B, 1, H, W, device=flow.device), torch.zeros(B, 1, H, W, device=flow.device), torch.zeros_like(flow))
    count, weight, out = bufs
    rc = _gpu().WeightedFlowProjection_gpu_forward_kernel(
        _stream(), _i(out.numel()), _i(W), _i(H), _i(2), _i(B), _i(fillhole), ctypes.c_float(threshold),
        *_s(flow), *_s(im0), *_s(im1), *_s(count), *_s(weight),
        _d(flow), _d(im0), _d(im1), _d(count), _d(weight), _d(out))
    assert rc == 0, rc
    return out, count, weight


def gpu_weighted_flow_projection_backward(flow, im0, im1, count, weight, gout, threshold, gi=None):
    import torch
    B, _, H, W = flow.shape
    if gi is None:
        gi = torch.zeros_like(flow)
    rc = _gpu().WeightedFlowProjection_gpu_backward_kernel(
        _stream(), _i(gout.numel()), _i(W), _i(H), _i(2), _i(B), ctypes.c_float(threshold),
        *_s(flow), *_s(im0), *_s(im1), *_s(count), *_s(weight),
        _d(flow), _d(im0), _d(im1), _d(count), _d(weight), _d(gout), _d(gi))
    assert rc == 0, rc
    return gi


def gpu_pixel_splat_forward(mode, flow, in1=None, fw=None, sigma_d=1.0, out=None):
    """reference kernels of the 4x4 splat family (my_lib_kernel.h:297-398); tao_r = 0, Prowindow = 2."""
    import torch
    B, _, H, W = flow.shape
    f3 = (_f(sigma_d), _f(0), _f(2))
    if mode == "value":
        C = in1.shape[1]
        out = torch.zeros_like(in1) if out is None else out
        rc = _gpu().PixelValueLayer_gpu_forward_kernel(
            _stream(), _i(out.numel()), _i(W), _i(H), _i(C), _i(B), *_s(in1), *_s(flow), *_s(fw), *_s(out),
            _d(in1), _d(flow), _d(fw), _d(out), *f3)
    elif mode == "weight":
        out = torch.zeros(B, 1, H, W, device=flow.device) if out is None else out
        rc = _gpu().PixelWeightLayer_gpu_forward_kernel(
            _stream(), _i(out.numel()), _i(W), _i(H), _i(B), *_s(flow), *_s(fw), *_s(out), _d(flow), _d(fw), _d(out), *f3)
    else:
        out = torch.zeros(B, 1, H, W, device=flow.device) if out is None else out
        rc = _gpu().ReliableWeightLayer_gpu_forward_kernel(
            _stream(), _i(out.numel()), _i(W), _i(H), _i(B), *_s(flow), *_s(out), _d(flow), _d(out), *f3)
    assert rc == 0, rc
    return out


def gpu_pixel_splat_backward(mode, flow, gout, in1=None, fw=None, fout=None, sigma_d=1.0, threshold=0.0):
    import torch
    B, _, H, W = flow.shape
    f3 = (_f(sigma_d), _f(0), _f(2))
    g3 = torch.zeros_like(flow)
    if mode == "value":
        C = in1.shape[1]
        g1, gw = torch.zeros_like(in1), torch.zeros_like(fw)
        rc = _gpu().PixelValueLayer_gpu_backward_kernel(
            _stream(), _i(gout.numel()), _i(W), _i(H), _i(C), _i(B), *_s(in1), *_s(flow), *_s(fw), *_s(gout),
            _d(in1), _d(flow), _d(fw), _d(gout), _d(g1), _d(g3), _d(gw), *f3)
        assert rc == 0, rc
        return g1, g3, gw
    if mode == "weight":
        gw = torch.zeros_like(fw)
        rc = _gpu().PixelWeightLayer_gpu_backward_kernel(
            _stream(), _i(gout.numel()), _i(W), _i(H), _i(B), *_s(flow), *_s(fw), *_s(fout),
            _d(flow), _d(fw), _d(fout), _d(gout), _d(g3), _d(gw), _f(threshold), *f3)
        assert rc == 0, rc
        return None, g3, gw
    rc = _gpu().ReliableWeightLayer_gpu_backward_kernel(
        _stream(), _i(gout.numel()), _i(W), _i(H), _i(B), *_s(flow), *_s(fout),
        _d(flow), _d(fout), _d(gout), _d(g3), _f(threshold), *f3)
    assert rc == 0, rc
    return None, g3, None


def gpu_weight_layer_forward(in1, in2, flow, lambda_e):
    import torch
    B, C, H, W = in1.shape
    out = torch.zeros(B, 1, H, W, device=in1.device)
    rc = _gpu().WeightLayer_gpu_forward_kernel(
        _stream(), _i(out.numel()), _i(W), _i(H), _i(C), _i(B), *_s(in1), *_s(in2), *_s(flow), *_s(out),
        _d(in1), _d(in2), _d(flow), _d(out), _f(lambda_e), _f(0), _f(3))
    assert rc == 0, rc
    return out


def gpu_weight_layer_backward(in1, in2, flow, fout, gout, lambda_e):
    import torch
    B, C, H, W = in1.shape
    g1, g2, g3 = torch.zeros_like(in1), torch.zeros_like(in2), torch.zeros_like(flow)
    rc = _gpu().WeightLayer_gpu_backward_kernel(
        _stream(), _i(gout.numel()), _i(W), _i(H), _i(C), _i(B), *_s(in1), *_s(in2), *_s(flow), *_s(fout),
        _d(in1), _d(in2), _d(flow), _d(fout), _d(gout), _d(g1), _d(g2), _d(g3), _f(lambda_e), _f(0), _f(3))
    assert rc == 0, rc
    return g1, g2, g3


def gpu_separable_conv_flow_forward(in1, vert, horiz):
    import torch
    B, C, H, W = in1.shape
    fs = vert.shape[1]
    flow = torch.zeros(B, 2, H - fs + 1, W - fs + 1, device=in1.device)
    rc = _gpu().SeparableConvFlowLayer_gpu_forward_kernel(
        _stream(), _i(flow.numel()), _i(W), _i(H), _i(C), _i(B), _i(fs), *_s(in1), *_s(vert), *_s(horiz), *_s(flow),
        _d(in1), _d(vert), _d(horiz), _d(flow))
    assert rc == 0, rc
    return flow


def gpu_separable_conv_flow_backward(in1, vert, horiz, gflow):
    import torch
    B, C, H, W = in1.shape
    fs = vert.shape[1]
    g1, gv, gh = torch.zeros_like(in1), torch.zeros_like(vert), torch.zeros_like(horiz)
    rc = _gpu().SeparableConvFlowLayer_gpu_backward_kernel(
        _stream(), _i(gflow.numel()), _i(W), _i(H), _i(C), _i(B), _i(fs), *_s(in1), *_s(vert), *_s(horiz), *_s(gflow),
        _d(in1), _d(vert), _d(horiz), _d(gflow), _d(g1), _d(gv), _d(gh))
    assert rc == 0, rc
    return gv, gh


def gpu_interpolation_forward(in1, flow, out=None):
    import torch
    B, C, H, W = in1.shape
    if out is None:
        out = torch.zeros_like(in1)
    rc = _gpu().InterpolationLayer_gpu_forward_kernel(
        _stream(), _i(out.numel()), _i(W), _i(H), _i(C), _i(B),
        *_s(in1), *_s(flow), _d(in1), _d(flow), _d(out))
    assert rc == 0, rc
    return out


def gpu_interpolation_backward(in1, flow, gout, grads=None):
    import torch
    B, C, H, W = in1.shape
    if grads is None:
        grads = (torch.zeros_like(in1), torch.zeros_like(flow))
    g1, g2 = grads
    rc = _gpu().InterpolationLayer_gpu_backward_kernel(
        _stream(), _i(gout.numel()), _i(W), _i(H), _i(C), _i(B),
        *_s(in1), *_s(flow), _d(in1), _d(flow), _d(gout), _d(g1), _d(g2))
    assert rc == 0, rc
    return g1, g2


def gpu_separable_conv_forward(in1, vert, horiz, out=None):
    import torch
    B, C, H, W = in1.shape
    fs = vert.shape[1]
    if out is None:
        out = torch.zeros(B, C, H - fs + 1, W - fs + 1, device=in1.device)
    rc = _gpu().SeparableConvLayer_gpu_forward_kernel(
        _stream(), _i(out.numel()), _i(W), _i(H), _i(C), _i(B), _i(fs),
        *_s(in1), *_s(vert), *_s(horiz), *_s(out), _d(in1), _d(vert), _d(horiz), _d(out))
    assert rc == 0, rc
    return out


def gpu_separable_conv_backward(in1, vert, horiz, gout, grads=None):
    import torch
    B, C, H, W = in1.shape
    fs = vert.shape[1]
    if grads is None:
        grads = (torch.zeros_like(in1), torch.zeros_like(vert), torch.zeros_like(horiz))
    g1, g2, g3 = grads
    rc = _gpu().SeparableConvLayer_gpu_backward_kernel(
        _stream(), _i(gout.numel()), _i(W), _i(H), _i(C), _i(B), _i(fs),
        *_s(in1), *_s(vert), *_s(horiz), *_s(gout),
        _d(in1), _d(vert), _d(horiz), _d(gout), _d(g1), _d(g2), _d(g3))
    assert rc == 0, rc
    return g1, g2, g3
