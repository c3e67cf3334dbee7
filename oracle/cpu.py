"""ctypes bindings of oracle/memc_oracle.c  (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Every function takes/returns dense NCHW float32 numpy arrays (outputs are float32 for the
f32 build and float64 for the f64 build) and mirrors one reference CPU entry point:

  filter_interpolation_forward/backward   my_lib.c:904-1079 / 1082-1444
  flow_projection_forward/backward        my_lib.c:1447-1547 / 1549-1634 (+ fill-hole from
                                          my_lib_kernel.cu:1742-1836)
  interpolation_forward/backward          my_lib.c:440-533 / 534-667
  separable_conv_forward/backward         my_lib.c:250-339 / 340-439
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def build(force=False):
    """Compile memc_oracle.c (both precisions) with gcc via oracle/Makefile."""
    want = [os.path.join(_HERE, "liboracle_f32.so"), os.path.join(_HERE, "liboracle_f64.so")]
    src = os.path.join(_HERE, "memc_oracle.c")
    stale = force or any(
        (not os.path.exists(p)) or os.path.getmtime(p) < os.path.getmtime(src) for p in want)
    if stale:
        subprocess.run(["make", "-s", "-C", _HERE, "oracle"] + (["-B"] if force else []), check=True)
    return want


def _lib(precision):
    if precision not in ("f32", "f64"):
        raise ValueError(precision)
    if precision not in _LIBS:
        build()
        lib = ctypes.CDLL(os.path.join(_HERE, "liboracle_%s.so" % precision))
        assert lib.oracle_real_bytes() == (4 if precision == "f32" else 8)
        _LIBS[precision] = lib
    return _LIBS[precision]


def _real(precision):
    return np.float32 if precision == "f32" else np.float64


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _check(rc, name):
    if rc != 0:
        raise ValueError("%s: precondition violated (rc=%d)" % (name, rc))


def filter_interpolation_forward(in1, flow, filt, precision="f32"):
    in1, flow, filt = _f32(in1), _f32(flow), _f32(filt)
    B, C, H, W = in1.shape
    fs = int(np.sqrt(np.float32(filt.shape[1])))  # my_lib_cuda.c:619-620
    assert flow.shape == (B, 2, H, W) and filt.shape[0] == B and filt.shape[2:] == (H, W)
    out = np.zeros((B, C, H, W), dtype=_real(precision))
    _check(_lib(precision).oracle_filter_interpolation_forward(
        B, C, H, W, fs, _p(in1), _p(flow), _p(filt), _p(out)), "filter_interpolation_forward")
    return out


def filter_interpolation_backward(in1, flow, filt, gout, precision="f32"):
    in1, flow, filt, gout = _f32(in1), _f32(flow), _f32(filt), _f32(gout)
    B, C, H, W = in1.shape
    fs = int(np.sqrt(np.float32(filt.shape[1])))
    r = _real(precision)
    gi1, gi2, gi3 = np.zeros(in1.shape, r), np.zeros(flow.shape, r), np.zeros(filt.shape, r)
    _check(_lib(precision).oracle_filter_interpolation_backward(
        B, C, H, W, fs, _p(in1), _p(flow), _p(filt), _p(gout), _p(gi1), _p(gi2), _p(gi3)),
        "filter_interpolation_backward")
    return gi1, gi2, gi3


def flow_projection_forward(flow, fillhole, precision="f32"):
    flow = _f32(flow)
    B, two, H, W = flow.shape
    assert two == 2
    count = np.zeros((B, 1, H, W), np.float32)
    out = np.zeros((B, 2, H, W), _real(precision))
    _check(_lib(precision).oracle_flow_projection_forward(
        B, H, W, _p(flow), _p(count), _p(out), int(fillhole)), "flow_projection_forward")
    return out, count


def flow_projection_backward(flow, count, gout, precision="f32"):
    flow, count, gout = _f32(flow), _f32(count), _f32(gout)
    B, _, H, W = flow.shape
    gi = np.zeros(flow.shape, _real(precision))
    _check(_lib(precision).oracle_flow_projection_backward(
        B, H, W, _p(flow), _p(count), _p(gout), _p(gi)), "flow_projection_backward")
    return gi


def depth_flow_projection_forward(flow, depth, fillhole, precision="f32"):
    flow, depth = _f32(flow), _f32(depth)
    B, two, H, W = flow.shape
    assert two == 2 and depth.shape == (B, 1, H, W)
    count = np.zeros((B, 1, H, W), np.float32)
    out = np.zeros((B, 2, H, W), _real(precision))
    _check(_lib(precision).oracle_depth_flow_projection_forward(
        B, H, W, _p(flow), _p(depth), _p(count), _p(out), int(fillhole)), "depth_flow_projection_forward")
    return out, count


def depth_flow_projection_backward(flow, depth, count, fout, gout, precision="f32"):
    flow, depth, count, fout, gout = _f32(flow), _f32(depth), _f32(count), _f32(fout), _f32(gout)
    B, _, H, W = flow.shape
    r = _real(precision)
    gi1, gi2 = np.zeros(flow.shape, r), np.zeros(depth.shape, r)
    _check(_lib(precision).oracle_depth_flow_projection_backward(
        B, H, W, _p(flow), _p(depth), _p(count), _p(fout), _p(gout), _p(gi1), _p(gi2)), "depth_flow_projection_backward")
    return gi1, gi2


def weighted_flow_projection_forward(flow, im0, im1, fillhole, threshold, precision="f32"):
    flow, im0, im1 = _f32(flow), _f32(im0), _f32(im1)
    B, two, H, W = flow.shape
    assert two == 2 and im0.shape == (B, 3, H, W) and im1.shape == (B, 3, H, W)
    r = _real(precision)
    count = np.zeros((B, 1, H, W), np.float32)
    weight, out = np.zeros((B, 1, H, W), r), np.zeros((B, 2, H, W), r)
    _check(_lib(precision).oracle_weighted_flow_projection_forward(
        B, H, W, _p(flow), _p(im0), _p(im1), _p(count), _p(weight), _p(out), int(fillhole), ctypes.c_float(threshold)),
        "weighted_flow_projection_forward")
    return out, count, weight


def weighted_flow_projection_backward(flow, im0, im1, count, gout, threshold, precision="f32"):
    flow, im0, im1, count, gout = _f32(flow), _f32(im0), _f32(im1), _f32(count), _f32(gout)
    B, _, H, W = flow.shape
    gi = np.zeros(flow.shape, _real(precision))
    _check(_lib(precision).oracle_weighted_flow_projection_backward(
        B, H, W, _p(flow), _p(im0), _p(im1), _p(count), _p(gout), _p(gi), ctypes.c_float(threshold)),
        "weighted_flow_projection_backward")
    return gi


_PX_MODES = {"value": 0, "weight": 1, "reliable": 2}


def pixel_splat_forward(mode, flow, in1=None, fw=None, sigma_d=1.0, precision="f32"):
    """PixelValue ("value": in1 [B,C,H,W] + fw [B,1,H,W]), PixelWeight ("weight": fw), ReliableWeight ("reliable")."""
    flow = _f32(flow)
    B, _, H, W = flow.shape
    C = in1.shape[1] if mode == "value" else 1
    in1 = _f32(in1) if in1 is not None else np.zeros(1, np.float32)
    fw = _f32(fw) if fw is not None else np.zeros(1, np.float32)
    out = np.zeros((B, C, H, W), _real(precision))
    _check(_lib(precision).oracle_pixel_splat_forward(_PX_MODES[mode], B, C, H, W, _p(in1), _p(flow), _p(fw), _p(out),
                                                      ctypes.c_float(sigma_d)), "pixel_splat_forward")
    return out


def pixel_splat_backward(mode, flow, gout, in1=None, fw=None, fout=None, sigma_d=1.0, threshold=0.0, precision="f32"):
    """-> (gi1 | None, gi3, gfw | None) as the op has them."""
    flow, gout = _f32(flow), _f32(gout)
    B, _, H, W = flow.shape
    r = _real(precision)
    C = in1.shape[1] if mode == "value" else 1
    in1_ = _f32(in1) if in1 is not None else np.zeros(1, np.float32)
    fw_ = _f32(fw) if fw is not None else np.zeros(1, np.float32)
    fout_ = _f32(fout) if fout is not None else np.zeros(1, np.float32)
    gi1 = np.zeros((B, C, H, W), r)
    gi3, gfw = np.zeros((B, 2, H, W), r), np.zeros((B, 1, H, W), r)
    _check(_lib(precision).oracle_pixel_splat_backward(
        _PX_MODES[mode], B, C, H, W, _p(in1_), _p(flow), _p(fw_), _p(fout_), _p(gout), _p(gi1), _p(gi3), _p(gfw),
        ctypes.c_float(sigma_d), ctypes.c_float(threshold)), "pixel_splat_backward")
    return (gi1 if mode == "value" else None), gi3, (gfw if mode != "reliable" else None)


def weight_layer_forward(in1, in2, flow, lambda_e, Nw=3.0, precision="f32"):
    in1, in2, flow = _f32(in1), _f32(in2), _f32(flow)
    B, C, H, W = in1.shape
    out = np.zeros((B, 1, H, W), _real(precision))
    _check(_lib(precision).oracle_weight_layer_forward(B, C, H, W, _p(in1), _p(in2), _p(flow), _p(out),
                                                       ctypes.c_float(lambda_e), ctypes.c_float(Nw)), "weight_layer_forward")
    return out


def weight_layer_backward(in1, in2, flow, fout, gout, lambda_e, Nw=3.0, precision="f32"):
    in1, in2, flow, fout, gout = _f32(in1), _f32(in2), _f32(flow), _f32(fout), _f32(gout)
    B, C, H, W = in1.shape
    r = _real(precision)
    g1, g2, g3 = np.zeros(in1.shape, r), np.zeros(in2.shape, r), np.zeros(flow.shape, r)
    _check(_lib(precision).oracle_weight_layer_backward(B, C, H, W, _p(in1), _p(in2), _p(flow), _p(fout), _p(gout), _p(g1), _p(g2),
                                                        _p(g3), ctypes.c_float(lambda_e), ctypes.c_float(Nw)), "weight_layer_backward")
    return g1, g2, g3


def separable_conv_flow_forward(vert, horiz, precision="f32"):
    vert, horiz = _f32(vert), _f32(horiz)
    B, fs, Ho, Wo = vert.shape
    flow = np.zeros((B, 2, Ho, Wo), _real(precision))
    _check(_lib(precision).oracle_separable_conv_flow_forward(B, fs, Ho, Wo, _p(vert), _p(horiz), _p(flow)),
           "separable_conv_flow_forward")
    return flow


def separable_conv_flow_backward(vert, horiz, gflow, precision="f32"):
    vert, horiz, gflow = _f32(vert), _f32(horiz), _f32(gflow)
    B, fs, Ho, Wo = vert.shape
    r = _real(precision)
    gv, gh = np.zeros(vert.shape, r), np.zeros(horiz.shape, r)
    _check(_lib(precision).oracle_separable_conv_flow_backward(B, fs, Ho, Wo, _p(vert), _p(horiz), _p(gflow), _p(gv), _p(gh)),
           "separable_conv_flow_backward")
    return gv, gh


def interpolation_forward(in1, flow, precision="f32"):
    in1, flow = _f32(in1), _f32(flow)
    B, C, H, W = in1.shape
    out = np.zeros(in1.shape, _real(precision))
    _check(_lib(precision).oracle_interpolation_forward(
        B, C, H, W, _p(in1), _p(flow), _p(out)), "interpolation_forward")
    return out


def interpolation_backward(in1, flow, gout, precision="f32"):
    in1, flow, gout = _f32(in1), _f32(flow), _f32(gout)
    B, C, H, W = in1.shape
    r = _real(precision)
    gi1, gi2 = np.zeros(in1.shape, r), np.zeros(flow.shape, r)
    _check(_lib(precision).oracle_interpolation_backward(
        B, C, H, W, _p(in1), _p(flow), _p(gout), _p(gi1), _p(gi2)), "interpolation_backward")
    return gi1, gi2


def separable_conv_forward(in1, vert, horiz, precision="f32"):
    in1, vert, horiz = _f32(in1), _f32(vert), _f32(horiz)
    B, C, H, W = in1.shape
    fs = vert.shape[1]
    out = np.zeros((B, C, H - fs + 1, W - fs + 1), _real(precision))
    _check(_lib(precision).oracle_separable_conv_forward(
        B, C, H, W, fs, _p(in1), _p(vert), _p(horiz), _p(out)), "separable_conv_forward")
    return out


def separable_conv_backward(in1, vert, horiz, gout, precision="f32"):
    in1, vert, horiz, gout = _f32(in1), _f32(vert), _f32(horiz), _f32(gout)
    B, C, H, W = in1.shape
    fs = vert.shape[1]
    r = _real(precision)
    gi1, gi2, gi3 = np.zeros(in1.shape, r), np.zeros(vert.shape, r), np.zeros(horiz.shape, r)
    _check(_lib(precision).oracle_separable_conv_backward(
        B, C, H, W, fs, _p(in1), _p(vert), _p(horiz), _p(gout), _p(gi1), _p(gi2), _p(gi3)),
        "separable_conv_backward")
    return gi1, gi2, gi3
