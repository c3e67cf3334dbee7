"""ctypes binding of libmemc_b200.so -- the ONLY compute path of this package.

There is deliberately no CPU or PyTorch fallback: if the CUDA library is missing or a
tensor is not a CUDA fp32 tensor the call raises.  (The CPU oracle under /oracle is test
infrastructure and is never imported from here.)
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libmemc_b200.so")

OVERWRITE = 1  # MEMC_B200_OVERWRITE
NO_FAST = 2    # MEMC_B200_NO_FAST
NO_ZERO = 4    # MEMC_B200_NO_ZERO
FLOAT_ACCUM = 8  # MEMC_B200_FLOAT_ACCUM


def variant(n):
    """MEMC_B200_VARIANT(n): select a non-production kernel variant (A/B measurements, cross-checks)."""
    return (int(n) & 0xFF) << 16

_lib = None

# How FilterInterpolation backward accumulates gradinput1 inside a tile (include/memc_b200.h):
#   "fixed"  per-tile int32 fixed point: every contribution rounded to <= 2^-22 x the tile's largest contribution
#            (order independent, the speed path; absolute error of gradinput1 ~4e-6 on the benchmark workload)
#   "float"  fp32 shared-memory atomics (MEMC_B200_FLOAT_ACCUM): keeps fp32 RELATIVE precision when gradient
#            magnitudes inside one 32x8 tile differ by many orders (masked / occlusion-weighted losses); slower
# Default from the environment variable MEMC_B200_FI_ACCUM, read once at import; set_fi_accumulation() at run time.
_fi_accum = os.environ.get("MEMC_B200_FI_ACCUM", "fixed").strip().lower()


def set_fi_accumulation(mode):
    global _fi_accum
    if mode not in ("fixed", "float"):
        raise ValueError("accumulation mode must be 'fixed' or 'float', got %r" % (mode,))
    _fi_accum = mode


def get_fi_accumulation():
    return "float" if _fi_accum == "float" else "fixed"


def fi_backward_flags():
    """Flags the autograd Functions pass to memc_b200_filter_interpolation_backward."""
    return OVERWRITE | (FLOAT_ACCUM if _fi_accum == "float" else 0)


class MemcB200Error(RuntimeError):
    pass


class Strides(ctypes.Structure):
    """memc_strides of include/memc_b200.h: element strides (b, c, h); w-stride is 1."""
    _fields_ = [("b", ctypes.c_int64), ("c", ctypes.c_int64), ("h", ctypes.c_int64)]


_I, _P, _S, _F = ctypes.c_int, ctypes.c_void_p, Strides, ctypes.c_float

# name -> argtypes, straight from include/memc_b200.h
_EXTENDED = {
    "memc_b200_filter_interpolation_forward": [_P, _I, _I, _I, _I, _I, _S, _S, _S, _S, _P, _P, _P, _P, _I],
    "memc_b200_filter_interpolation_backward": [_P, _I, _I, _I, _I, _I, _S, _S, _S, _S, _S, _S, _S,
                                                _P, _P, _P, _P, _P, _P, _P, _I],
    "memc_b200_filter_interpolation_blend_forward": [_P, _I, _I, _I, _I, _I] + [_S] * 9 + [_P] * 9 + [_I],
    "memc_b200_filter_interpolation_forward_pair": [_P, _I, _I, _I, _I, _I, _I] + [_S] * 6 + [_P] * 6 + [_I],
    "memc_b200_flow_projection_forward": [_P, _I, _I, _I, _I, _S, _S, _S, _P, _P, _P, _I],
    "memc_b200_flow_projection_backward": [_P, _I, _I, _I, _S, _S, _S, _S, _P, _P, _P, _P, _I],
    "memc_b200_depth_flow_projection_forward": [_P, _I, _I, _I, _I, _S, _S, _S, _S, _P, _P, _P, _P, _I],
    "memc_b200_depth_flow_projection_backward": [_P, _I, _I, _I] + [_S] * 7 + [_P] * 7 + [_I],
    "memc_b200_weighted_flow_projection_forward": [_P, _I, _I, _I, _I, _F] + [_S] * 6 + [_P] * 6 + [_I],
    "memc_b200_weighted_flow_projection_backward": [_P, _I, _I, _I, _F] + [_S] * 6 + [_P] * 6 + [_I],
    "memc_b200_weight_layer_forward": [_P, _I, _I, _I, _I, _F, _F] + [_S] * 4 + [_P] * 4 + [_I],
    "memc_b200_weight_layer_backward": [_P, _I, _I, _I, _I, _F, _F] + [_S] * 4 + [_P] * 8 + [_I],
    "memc_b200_separable_conv_flow_forward": [_P, _I, _I, _I, _I, _S, _S, _S, _P, _P, _P, _I],
    "memc_b200_separable_conv_flow_backward": [_P, _I, _I, _I, _I] + [_S] * 5 + [_P] * 5 + [_I],
    "memc_b200_pixel_value_forward": [_P, _I, _I, _I, _I, _F] + [_S] * 4 + [_P] * 4 + [_I],
    "memc_b200_pixel_value_backward": [_P, _I, _I, _I, _I, _F] + [_S] * 7 + [_P] * 7 + [_I],
    "memc_b200_pixel_weight_forward": [_P, _I, _I, _I, _F] + [_S] * 3 + [_P] * 3 + [_I],
    "memc_b200_pixel_weight_backward": [_P, _I, _I, _I, _F, _F] + [_S] * 6 + [_P] * 6 + [_I],
    "memc_b200_reliable_weight_forward": [_P, _I, _I, _I, _F] + [_S] * 2 + [_P] * 2 + [_I],
    "memc_b200_reliable_weight_backward": [_P, _I, _I, _I, _F, _F] + [_S] * 4 + [_P] * 4 + [_I],
    "memc_b200_interpolation_forward": [_P, _I, _I, _I, _I, _S, _S, _S, _P, _P, _P, _I],
    "memc_b200_interpolation_backward": [_P, _I, _I, _I, _I, _S, _S, _S, _S, _S, _P, _P, _P, _P, _P, _I],
    "memc_b200_separable_conv_forward": [_P, _I, _I, _I, _I, _I, _S, _S, _S, _S, _P, _P, _P, _P, _I],
    "memc_b200_separable_conv_backward": [_P, _I, _I, _I, _I, _I, _S, _S, _S, _S, _S, _S, _S,
                                          _P, _P, _P, _P, _P, _P, _P, _I],
}
# reference-named launchers: (stream, nElement, w, h, channel, batch[, fs|fillhole]), n stride ints, n pointers
_NAMED = {
    "FilterInterpolationLayer_gpu_forward_kernel": [_P] + [_I] * 6 + [_I] * 12 + [_P] * 4,
    "FilterInterpolationLayer_gpu_backward_kernel": [_P] + [_I] * 6 + [_I] * 12 + [_P] * 7,
    "FlowProjection_gpu_forward_kernel": [_P] + [_I] * 6 + [_I] * 8 + [_P] * 3,
    "FlowProjection_gpu_backward_kernel": [_P] + [_I] * 5 + [_I] * 8 + [_P] * 4,
    "DepthFlowProjection_gpu_forward_kernel": [_P] + [_I] * 6 + [_I] * 12 + [_P] * 4,
    "DepthFlowProjection_gpu_backward_kernel": [_P] + [_I] * 5 + [_I] * 12 + [_P] * 7,
    "WeightedFlowProjection_gpu_forward_kernel": [_P] + [_I] * 6 + [_F] + [_I] * 20 + [_P] * 6,
    "WeightedFlowProjection_gpu_backward_kernel": [_P] + [_I] * 5 + [_F] + [_I] * 20 + [_P] * 7,
    "WeightLayer_gpu_forward_kernel": [_P] + [_I] * 5 + [_I] * 16 + [_P] * 4 + [_F] * 3,
    "WeightLayer_gpu_backward_kernel": [_P] + [_I] * 5 + [_I] * 16 + [_P] * 8 + [_F] * 3,
    "SeparableConvFlowLayer_gpu_forward_kernel": [_P] + [_I] * 6 + [_I] * 16 + [_P] * 4,
    "SeparableConvFlowLayer_gpu_backward_kernel": [_P] + [_I] * 6 + [_I] * 16 + [_P] * 7,
    "PixelValueLayer_gpu_forward_kernel": [_P] + [_I] * 5 + [_I] * 16 + [_P] * 4 + [_F] * 3,
    "PixelValueLayer_gpu_backward_kernel": [_P] + [_I] * 5 + [_I] * 16 + [_P] * 7 + [_F] * 3,
    "PixelWeightLayer_gpu_forward_kernel": [_P] + [_I] * 4 + [_I] * 12 + [_P] * 3 + [_F] * 3,
    "PixelWeightLayer_gpu_backward_kernel": [_P] + [_I] * 4 + [_I] * 12 + [_P] * 6 + [_F] * 4,
    "ReliableWeightLayer_gpu_forward_kernel": [_P] + [_I] * 4 + [_I] * 8 + [_P] * 2 + [_F] * 3,
    "ReliableWeightLayer_gpu_backward_kernel": [_P] + [_I] * 4 + [_I] * 8 + [_P] * 4 + [_F] * 4,
    "InterpolationLayer_gpu_forward_kernel": [_P] + [_I] * 5 + [_I] * 8 + [_P] * 3,
    "InterpolationLayer_gpu_backward_kernel": [_P] + [_I] * 5 + [_I] * 8 + [_P] * 5,
    "InterpolationChLayer_gpu_forward_kernel": [_P] + [_I] * 5 + [_I] * 8 + [_P] * 3,
    "InterpolationChLayer_gpu_backward_kernel": [_P] + [_I] * 5 + [_I] * 8 + [_P] * 5,
    "SeparableConvLayer_gpu_forward_kernel": [_P] + [_I] * 6 + [_I] * 16 + [_P] * 4,
    "SeparableConvLayer_gpu_backward_kernel": [_P] + [_I] * 6 + [_I] * 16 + [_P] * 7,
}
EXPORTS = sorted(list(_EXTENDED) + list(_NAMED) +
                 ["memc_b200_abi_version", "memc_b200_build_info", "memc_b200_launch_count", "memc_b200_scratch_trim"])


def load():
    """Load libmemc_b200.so (once).  Raises MemcB200Error if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MemcB200Error(
            "libmemc_b200.so is missing (%s). Build it with `python memc-net_b200/build.py`; "
            "this package has no CPU / PyTorch fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in list(_EXTENDED.items()) + list(_NAMED.items()):
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = ctypes.c_int
    lib.memc_b200_abi_version.restype = ctypes.c_int
    lib.memc_b200_build_info.restype = ctypes.c_char_p
    lib.memc_b200_launch_count.restype = ctypes.c_ulonglong
    lib.memc_b200_scratch_trim.restype = ctypes.c_int
    _lib = lib
    return lib


def launch_count():
    return int(load().memc_b200_launch_count())


def scratch_trim():
    """Release the library's cached scratch memory (stream-ordered pool) back to the driver."""
    return int(load().memc_b200_scratch_trim())


def build_info():
    return load().memc_b200_build_info().decode()


# ------------------------------------------------------------------------------ helpers
def check_tensor(t, name, ndim=4):
    if not isinstance(t, torch.Tensor):
        raise MemcB200Error("%s: expected a torch.Tensor, got %r" % (name, type(t)))
    if not t.is_cuda:
        raise MemcB200Error("%s: must be a CUDA tensor (this package has no CPU path)" % name)
    if t.dtype != torch.float32:
        raise MemcB200Error("%s: must be float32, got %s" % (name, t.dtype))
    if t.dim() != ndim:
        raise MemcB200Error("%s: must be %d-D (NCHW), got %d-D" % (name, ndim, t.dim()))
    return t


def check_same_device(*tensors):
    """Every operand of one call must live on ONE device (the library launches on the device that owns them)."""
    devs = {t.device for t in tensors if isinstance(t, torch.Tensor)}
    if len(devs) > 1:
        raise MemcB200Error("operands are on different devices: %s" % sorted(str(d) for d in devs))


def strides_of(t):
    sb, sc, sh, sw = t.stride()
    if sw != 1 and t.size(3) != 1:
        raise MemcB200Error("w-stride must be 1 (got %d); call .contiguous()" % sw)
    return Strides(sb, sc, sh)


def stream_ptr(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def call(name, *args):
    """Invoke an entry point; raise on a non-zero return (launch error / bad layout)."""
    rc = getattr(load(), name)(*args)
    if rc != 0:
        raise MemcB200Error("%s returned %d" % (name, rc))
    return rc
