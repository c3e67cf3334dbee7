"""Pins oracle/memc_oracle.c (our restatement) to the reference's OWN CPU code.

oracle/_ref/libmemc_ref_cpu.so is my_package/src/my_lib.c compiled unchanged (oracle/Makefile);
the float build of the oracle must agree with it BIT FOR BIT on seeded inputs, including the
edge cases of tests/cases.py.  Skipped where oracle/_ref is absent (it is built wherever
/root/reference exists and travels with the repo snapshot); the committed fixtures under
tests/golden/ (test_golden.py) carry the same pin everywhere.
"""
import numpy as np
import pytest

from oracle import cpu, ref
from tests.cases import fi_case, flow_case, sepconv_case

needs_ref = pytest.mark.skipif(not ref.available_cpu(), reason="oracle/_ref/libmemc_ref_cpu.so not built")

FI_SHAPES = [  # B, C, H, W, fs, sigma
    (1, 3, 64, 64, 4, 3.0),    # BASELINE.json configs[0]
    (3, 3, 64, 64, 4, 3.0),
    (2, 3, 37, 53, 4, 8.0),    # ragged
    (1, 5, 20, 31, 5, 2.0),    # odd filter size, C != 3
    (1, 2, 16, 16, 2, 1.0),
    (1, 3, 24, 24, 6, 40.0),   # mostly out-of-range flow
    (1, 64, 16, 24, 4, 2.0),   # MEMC_Net_star context channels
    (1, 1, 1, 1, 4, 0.0),      # degenerate 1x1
]


@needs_ref
@pytest.mark.parametrize("shape", FI_SHAPES)
def test_filter_interpolation_bit_exact(shape):
    B, C, H, W, fs, sigma = shape
    in1, flow, filt, gout = fi_case(B, C, H, W, fs, sigma, seed=hash(shape) % 1000)
    assert np.array_equal(cpu.filter_interpolation_forward(in1, flow, filt),
                          ref.cpu_filter_interpolation_forward(in1, flow, filt))
    for a, b in zip(cpu.filter_interpolation_backward(in1, flow, filt, gout),
                    ref.cpu_filter_interpolation_backward(in1, flow, filt, gout)):
        assert np.array_equal(a, b)


@needs_ref
@pytest.mark.parametrize("shape", [(1, 64, 64, 3.0), (2, 37, 53, 8.0), (1, 24, 24, 40.0), (1, 1, 1, 0.0)])
def test_flow_projection_bit_exact(shape):
    B, H, W, sigma = shape
    flow = flow_case(B, H, W, sigma, seed=11)
    out, count = cpu.flow_projection_forward(flow, fillhole=0)
    rout, rcount = ref.cpu_flow_projection_forward(flow)
    assert np.array_equal(count, rcount) and np.array_equal(out, rout)
    gout = np.random.default_rng(5).standard_normal(flow.shape).astype(np.float32)
    assert np.array_equal(cpu.flow_projection_backward(flow, count, gout),
                          ref.cpu_flow_projection_backward(flow, count, gout))


@needs_ref
@pytest.mark.parametrize("shape", [(1, 64, 64, 3.0), (2, 37, 53, 8.0), (1, 24, 24, 40.0), (1, 1, 1, 0.0)])
@pytest.mark.parametrize("weights", ["inverse_depth", "signed"])
def test_depth_flow_projection_bit_exact(shape, weights):
    """the depth-weighted splat against my_lib.c:1637-1877 (forward without fill-hole: the
    reference's CPU twin has none; backward reads the forward's output)."""
    B, H, W, sigma = shape
    flow = flow_case(B, H, W, sigma, seed=13)
    rng = np.random.default_rng(17)
    depth = (1e-6 + 1.0 / rng.uniform(0.5, 20.0, (B, 1, H, W))).astype(np.float32)
    if weights == "signed":  # not an inverse depth: negative and zero weights, cells whose accumulated weight is <= 0
        depth = rng.standard_normal((B, 1, H, W)).astype(np.float32)
        depth[rng.random(depth.shape) < 0.1] = 0.0
    out, count = cpu.depth_flow_projection_forward(flow, depth, fillhole=0)
    rout, rcount = ref.cpu_depth_flow_projection_forward(flow, depth)
    assert np.array_equal(count, rcount) and np.array_equal(out, rout, equal_nan=True)
    gout = rng.standard_normal(flow.shape).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        for a, b in zip(cpu.depth_flow_projection_backward(flow, depth, count, out, gout),
                        ref.cpu_depth_flow_projection_backward(flow, depth, count, out, gout)):
            assert np.array_equal(a, b, equal_nan=True)


@needs_ref
@pytest.mark.parametrize("shape", [(1, 64, 64, 3.0), (2, 37, 53, 8.0), (1, 24, 24, 40.0), (1, 1, 1, 0.0)])
@pytest.mark.parametrize("threshold", [0.0, 0.25, 0.4, 2.0])
def test_weighted_flow_projection_bit_exact(shape, threshold):
    """the brightness-gated splat against my_lib.c:1879-2250 (forward without fill-hole: the
    reference's CPU twin has none).  Thresholds: nobody votes / about a third / most / everybody."""
    B, H, W, sigma = shape
    flow = flow_case(B, H, W, sigma, seed=19)
    rng = np.random.default_rng(23)
    im0, im1 = rng.random((B, 3, H, W), dtype=np.float32), rng.random((B, 3, H, W), dtype=np.float32)
    out, count, weight = cpu.weighted_flow_projection_forward(flow, im0, im1, 0, threshold)
    rout, rcount, rweight = ref.cpu_weighted_flow_projection_forward(flow, im0, im1, threshold)
    assert np.array_equal(count, rcount) and np.array_equal(out, rout) and np.array_equal(weight, rweight)
    if threshold == 2.0:
        assert count.sum() > 0
    gout = rng.standard_normal(flow.shape).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        assert np.array_equal(cpu.weighted_flow_projection_backward(flow, im0, im1, count, gout, threshold),
                              ref.cpu_weighted_flow_projection_backward(flow, im0, im1, count, weight, gout, threshold),
                              equal_nan=True)


@needs_ref
@pytest.mark.parametrize("shape", [(1, 64, 64, 3.0), (2, 37, 53, 8.0), (1, 24, 24, 40.0), (1, 1, 1, 0.0)])
@pytest.mark.parametrize("mode", ["value", "weight", "reliable"])
def test_pixel_splat_family_bit_exact(shape, mode):
    """PixelValue / PixelWeight / ReliableWeight against my_lib.c:2615-3400, forward and
    backward (the threshold of the weight ops' backward set so that about a third of the cells are skipped)."""
    B, H, W, sigma = shape
    flow = flow_case(B, H, W, sigma, seed=29)
    rng = np.random.default_rng(31)
    in1 = rng.random((B, 3, H, W), dtype=np.float32) if mode == "value" else None
    fw = rng.random((B, 1, H, W), dtype=np.float32) if mode != "reliable" else None
    sigma_d = 1.3
    out = cpu.pixel_splat_forward(mode, flow, in1, fw, sigma_d)
    rout = ref.cpu_pixel_splat_forward(mode, flow, in1, fw, sigma_d)
    assert np.array_equal(out, rout)
    gout = rng.standard_normal(out.shape).astype(np.float32)
    thr = float(np.quantile(out, 0.35)) if mode != "value" else 0.0
    got = cpu.pixel_splat_backward(mode, flow, gout, in1, fw, out, sigma_d, thr)
    exp = ref.cpu_pixel_splat_backward(mode, flow, gout, in1, fw, out, sigma_d, thr)
    for a, b in zip(got, exp):
        assert (a is None) == (b is None)
        if a is not None:
            assert np.array_equal(a, b)


@needs_ref
@pytest.mark.parametrize("shape", [(1, 64, 64, 3.0), (2, 37, 53, 8.0), (1, 24, 24, 40.0), (1, 1, 1, 0.0)])
def test_weight_layer_bit_exact(shape):
    """WeightLayer against my_lib.c:2251-2614, forward and backward (sign decisions included: same fp32 expressions)."""
    B, H, W, sigma = shape
    flow = flow_case(B, H, W, sigma, seed=41)
    rng = np.random.default_rng(43)
    in1, in2 = rng.random((B, 3, H, W), dtype=np.float32), rng.random((B, 3, H, W), dtype=np.float32)
    lam = 0.9
    out = cpu.weight_layer_forward(in1, in2, flow, lam)
    assert np.array_equal(out, ref.cpu_weight_layer_forward(in1, in2, flow, lam))
    gout = rng.standard_normal(out.shape).astype(np.float32)
    for a, b in zip(cpu.weight_layer_backward(in1, in2, flow, out, gout, lam),
                    ref.cpu_weight_layer_backward(in1, in2, flow, out, gout, lam)):
        assert np.array_equal(a, b)


@needs_ref
@pytest.mark.parametrize("shape", [(1, 3, 32, 32, 4), (2, 3, 21, 35, 5), (1, 3, 9, 9, 3), (1, 3, 4, 4, 4)])
def test_separable_conv_flow_bit_exact(shape):
    """SeparableConvFlow against my_lib.c:13-249 on filters with a positive tap sum (the reference's C source divides by
    |sum|, its CUDA source -- which the oracle follows -- by the signed sum: they agree there), plus taps that sum to
    exactly 0 (-2000 / no gradient in both)."""
    B, C, H, W, fs = shape
    in1, v, hz, _ = sepconv_case(B, C, H, W, fs, seed=37)
    v, hz = np.abs(v) + np.float32(0.01), np.abs(hz) + np.float32(0.01)
    v[0, :, 0, 0] = 0.0
    hz[-1, :, -1, -1] = 0.0
    assert np.array_equal(cpu.separable_conv_flow_forward(v, hz), ref.cpu_separable_conv_flow_forward(in1, v, hz))
    gflow = np.random.default_rng(7).standard_normal((B, 2, H - fs + 1, W - fs + 1)).astype(np.float32)
    for a, b in zip(cpu.separable_conv_flow_backward(v, hz, gflow), ref.cpu_separable_conv_flow_backward(in1, v, hz, gflow)):
        assert np.array_equal(a, b)


@needs_ref
@pytest.mark.parametrize("shape", [(1, 3, 64, 64, 3.0), (2, 3, 37, 53, 8.0), (1, 7, 20, 31, 2.0)])
def test_interpolation_bit_exact(shape):
    B, C, H, W, sigma = shape
    in1, flow, _, gout = fi_case(B, C, H, W, 4, sigma, seed=7)
    assert np.array_equal(cpu.interpolation_forward(in1, flow), ref.cpu_interpolation_forward(in1, flow))
    for a, b in zip(cpu.interpolation_backward(in1, flow, gout), ref.cpu_interpolation_backward(in1, flow, gout)):
        assert np.array_equal(a, b)


@needs_ref
@pytest.mark.parametrize("shape", [(1, 3, 32, 32, 4), (2, 3, 21, 35, 5), (1, 3, 9, 9, 3)])
def test_separable_conv_bit_exact(shape):
    B, C, H, W, fs = shape
    in1, v, hz, gout = sepconv_case(B, C, H, W, fs, seed=3)
    assert np.array_equal(cpu.separable_conv_forward(in1, v, hz), ref.cpu_separable_conv_forward(in1, v, hz))
    for a, b in zip(cpu.separable_conv_backward(in1, v, hz, gout),
                    ref.cpu_separable_conv_backward(in1, v, hz, gout)):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("shape", [(2, 3, 37, 53, 4, 5.0), (1, 64, 16, 24, 4, 2.0)])
def test_f64_build_agrees_with_f32(shape):
    """The f64 build makes the same geometric decisions, so it differs only by rounding."""
    B, C, H, W, fs, sigma = shape
    in1, flow, filt, gout = fi_case(B, C, H, W, fs, sigma, seed=1)
    o32 = cpu.filter_interpolation_forward(in1, flow, filt, "f32")
    o64 = cpu.filter_interpolation_forward(in1, flow, filt, "f64")
    assert o64.dtype == np.float64 and np.abs(o32 - o64).max() < 2e-6
    for a, b in zip(cpu.filter_interpolation_backward(in1, flow, filt, gout, "f32"),
                    cpu.filter_interpolation_backward(in1, flow, filt, gout, "f64")):
        assert np.abs(a - b).max() < 1e-4 * max(1.0, np.abs(b).max())
