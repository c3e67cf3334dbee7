"""Analytic known-answer / property tests of the oracle; they hold
for the reference by construction of its algorithm and need no reference binary."""
import numpy as np

from oracle import cpu, pyloop


def _rng(seed=0):
    return np.random.default_rng(seed)


def test_zero_flow_one_hot_centre_tap_is_identity():
    # flow 0: ix = w, window origin L = w-1, T = h-1; tap k = 1*4+1 = 5 is the pixel itself
    B, C, H, W = 2, 3, 17, 23
    in1 = _rng().random((B, C, H, W), dtype=np.float32)
    flow = np.zeros((B, 2, H, W), np.float32)
    filt = np.zeros((B, 16, H, W), np.float32)
    filt[:, 5] = 1.0
    assert np.array_equal(cpu.filter_interpolation_forward(in1, flow, filt), in1)


def test_bilinear_taps_equal_plain_interpolation():
    # taps {5,6,9,10} = 1 -> TL,TR,BL,BR are the 4 bilinear neighbours -> equals Interpolation
    # wherever both ops are valid and no clamping differs (interior targets)
    B, C, H, W = 1, 3, 32, 40
    rng = _rng(1)
    in1 = rng.random((B, C, H, W), dtype=np.float32)
    flow = (rng.standard_normal((B, 2, H, W)) * 2).astype(np.float32)
    filt = np.zeros((B, 16, H, W), np.float32)
    filt[:, [5, 6, 9, 10]] = 1.0
    fi = cpu.filter_interpolation_forward(in1, flow, filt)
    ip = cpu.interpolation_forward(in1, flow)
    ys, xs = np.mgrid[0:H, 0:W].astype(np.float32)
    x2, y2 = xs + flow[0, 0], ys + flow[0, 1]
    inside = (x2 >= 0) & (y2 >= 0) & (x2 <= W - 2) & (y2 <= H - 2)
    assert inside.mean() > 0.5
    assert np.abs(fi - ip)[0][:, inside].max() < 1e-6


def test_integer_flow_uses_top_left_quadrant_only():
    B, C, H, W = 1, 2, 12, 12
    rng = _rng(2)
    in1 = rng.random((B, C, H, W), dtype=np.float32)
    flow = np.zeros((B, 2, H, W), np.float32)
    flow[:, 0] = 2.0
    flow[:, 1] = -1.0
    filt = rng.standard_normal((B, 16, H, W)).astype(np.float32)
    out = cpu.filter_interpolation_forward(in1, flow, filt)
    filt2 = filt.copy()
    filt2[:, [2, 3, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15]] = 123.0  # TR, BL, BR taps: weight alpha/beta = 0
    assert np.array_equal(out, cpu.filter_interpolation_forward(in1, flow, filt2))


def test_out_of_range_copies_input_and_kills_gradients():
    B, C, H, W = 1, 3, 10, 14
    rng = _rng(3)
    in1 = rng.random((B, C, H, W), dtype=np.float32)
    flow = np.zeros((B, 2, H, W), np.float32)
    flow[:, 0, :, :] = W / 2.0            # |fx| >= W/2 everywhere -> invalid everywhere
    filt = rng.standard_normal((B, 16, H, W)).astype(np.float32)
    gout = rng.standard_normal((B, C, H, W)).astype(np.float32)
    assert np.array_equal(cpu.filter_interpolation_forward(in1, flow, filt), in1)
    for g in cpu.filter_interpolation_backward(in1, flow, filt, gout):
        assert not g.any()


def test_flow_projection_zero_flow_counts():
    # every source hits (T,L),(T,R),(B,L),(B,R) with R/B clamped at the border
    H, W = 6, 7
    out, count = cpu.flow_projection_forward(np.zeros((1, 2, H, W), np.float32), fillhole=0)
    c = count[0, 0]
    assert c[0, 0] == 1 and c[0, 3] == 2 and c[3, 0] == 2 and c[3, 3] == 4
    assert c[H - 1, 3] == 6 and c[3, W - 1] == 6 and c[H - 1, W - 1] == 9
    assert not out.any()
    assert count.sum() == 4 * H * W


def test_flow_projection_constant_flow_gives_minus_flow():
    H, W = 16, 20
    flow = np.zeros((1, 2, H, W), np.float32)
    flow[:, 0] = 2.25
    flow[:, 1] = -1.5
    out, count = cpu.flow_projection_forward(flow, fillhole=0)
    hit = count[0, 0] > 0
    assert np.allclose(out[0, 0][hit], -2.25) and np.allclose(out[0, 1][hit], 1.5)
    assert not out[0, :, ~hit].any()


def test_fill_hole_never_looks_down():
    # a single counted row at the bottom: holes above it see nothing left/right/up -> stay 0;
    # a single counted row at the top: holes below copy it (search "up")
    H, W = 8, 8
    flow = np.full((1, 2, H, W), np.nan, np.float32)      # NaN -> no splat at all
    flow[0, :, H - 1, :] = 0.0                            # bottom row splats onto itself
    flow[0, 0, H - 1, :] = 0.0
    out, count = cpu.flow_projection_forward(flow.copy(), fillhole=1)
    assert (count[0, 0, :H - 1] == 0).all() and not out[0, :, :H - 1].any()
    flow = np.full((1, 2, H, W), np.nan, np.float32)
    flow[0, 0, 0, :] = 0.25                               # top row splats to rows 0 and 1
    flow[0, 1, 0, :] = 0.0
    out, count = cpu.flow_projection_forward(flow.copy(), fillhole=1)
    assert (count[0, 0, 2:] == 0).all()
    assert np.allclose(out[0, 0, 2:, 1:-1], -0.25)        # filled from above


def test_fill_hole_means_left_right_up():
    H, W = 5, 9
    flow = np.full((1, 2, H, W), np.nan, np.float32)
    # counted cells: (2,1) value -1, (2,7) value -3, (0,4) value -5 ; hole at (2,4)
    for (y, x, v) in [(2, 1, 1.0), (2, 7, 3.0), (0, 4, 5.0)]:
        flow[0, 0, y, x] = 0.0
        flow[0, 1, y, x] = 0.0
    out0, count = cpu.flow_projection_forward(flow.copy(), fillhole=0)
    # put distinguishable values by using an integer shift instead: simpler -> check the mean rule
    out, _ = cpu.flow_projection_forward(flow.copy(), fillhole=1)
    assert count[0, 0, 2, 4] == 0
    # neighbours found: left (2,2) [splat of (2,1) covers x=1,2], right (2,7), up (1,4) [covers rows 0,1]
    exp = (out0[0, 0, 2, 2] + out0[0, 0, 2, 7] + out0[0, 0, 1, 4]) / 3.0
    assert out[0, 0, 2, 4] == np.float32(exp)


def test_filter_interpolation_autograd_matches_finite_differences():
    # fractional parts in [0.1, 0.9]: the op is smooth there (not at integer crossings)
    B, C, H, W = 1, 2, 9, 10
    rng = _rng(9)
    in1 = rng.random((B, C, H, W), dtype=np.float32)
    base = rng.integers(-2, 3, size=(B, 2, H, W)).astype(np.float32)
    flow = (base + rng.uniform(0.1, 0.9, size=(B, 2, H, W))).astype(np.float32)
    filt = rng.standard_normal((B, 16, H, W)).astype(np.float32)
    gout = rng.standard_normal((B, C, H, W)).astype(np.float32)
    g1, g2, g3 = cpu.filter_interpolation_backward(in1, flow, filt, gout, "f64")

    def loss(i1, fl, ft):
        return float((cpu.filter_interpolation_forward(i1, fl, ft, "f64") * gout).sum())

    eps = 1e-3
    for (arr, grad, idx) in [(in1, g1, (0, 1, 4, 5)), (filt, g3, (0, 6, 3, 3)), (flow, g2, (0, 0, 4, 4)),
                             (flow, g2, (0, 1, 5, 6))]:
        hi, lo = arr.copy(), arr.copy()
        hi[idx] += eps
        lo[idx] -= eps
        args_hi = [hi if a is arr else a for a in (in1, flow, filt)]
        args_lo = [lo if a is arr else a for a in (in1, flow, filt)]
        fd = (loss(*args_hi) - loss(*args_lo)) / (2 * eps)
        assert abs(fd - grad[idx]) < 5e-3 * max(1.0, abs(fd)), (idx, fd, grad[idx])


def test_python_loop_oracle_64x64():
    """BASELINE.json configs[0]: FilterInterpolation forward 64x64 RGB, 4x4 kernel, as a
    literal pure-Python loop, against the C restatement (bit-exact in fp32)."""
    from tests.cases import fi_case
    in1, flow, filt, _ = fi_case(1, 3, 64, 64, 4, 3.0, seed=0)
    assert np.array_equal(pyloop.filter_interpolation_forward(in1, flow, filt),
                          cpu.filter_interpolation_forward(in1, flow, filt))


def test_python_loop_flow_projection_small():
    from tests.cases import flow_case
    flow = flow_case(1, 12, 13, 3.0, seed=2)
    out, count = cpu.flow_projection_forward(flow, fillhole=1)
    pout, pcount = pyloop.flow_projection_forward(flow, fillhole=1)
    assert np.array_equal(count, pcount) and np.array_equal(out, pout)


def test_flow_projection_is_a_box_filter_of_the_corner_histogram():
    """The identity the fast FlowProjection kernels rest on (flow_projection_fast.cu): every valid source adds the
    SAME value to the cells (L..R) x (T..Bm), R = min(L+1, W-1), Bm = min(T+1, H-1) (my_lib.c:1491-1524), so the
    splat equals a 2x2 box filter of the corner histogram A[T][L] += value, with the last column / row of A counted
    twice where the clamp repeats a cell.  Checked against the oracle incl. targets on the right / bottom border."""
    B, H, W = 2, 19, 27
    rng = _rng(5)
    flow = (rng.standard_normal((B, 2, H, W)) * 4).astype(np.float32)
    flow[0, 0, :, -1] = 0.0                       # x2 == W-1 exactly: L == R
    flow[1, 1, -1, :] = 0.0                       # y2 == H-1 exactly: T == Bm
    out, count = cpu.flow_projection_forward(flow, 0, "f64")
    for b in range(B):
        ys, xs = np.mgrid[0:H, 0:W]
        x2 = (xs.astype(np.float32) + flow[b, 0]).astype(np.float32)
        y2 = (ys.astype(np.float32) + flow[b, 1]).astype(np.float32)
        valid = (x2 >= 0) & (y2 >= 0) & (x2 <= W - 1) & (y2 <= H - 1)
        L, T = x2[valid].astype(np.int64), y2[valid].astype(np.int64)
        A = np.zeros((3, H, W))
        np.add.at(A[0], (T, L), -flow[b, 0][valid].astype(np.float64))
        np.add.at(A[1], (T, L), -flow[b, 1][valid].astype(np.float64))
        np.add.at(A[2], (T, L), 1.0)
        wx = np.ones(W); wx[-1] = 2.0             # a corner in the last column hits its own cell twice
        wy = np.ones(H); wy[-1] = 2.0
        Aw = A * wx[None, None, :]
        hsum = Aw.copy(); hsum[:, :, 1:] += A[:, :, :-1]          # cell x <- A[x] (* wx) + A[x-1]
        cells = hsum * wy[None, :, None]
        cells[:, 1:, :] += hsum[:, :-1, :]                          # cell y <- h[y] (* wy) + h[y-1]
        assert np.array_equal(cells[2], count[b, 0].astype(np.float64))
        avg = np.where(cells[2] > 0, cells[:2] / np.maximum(cells[2], 1), 0.0)
        assert np.abs(avg - out[b]).max() <= 1e-9
