"""Size-independent properties at BASELINE.json's full sizes (1920x1080), where the CPU oracle
would take too long: identities, adjointness of forward/backward (the backward kernels must be
the exact transpose of the forward ones wherever the op is linear), splat conservation."""
import pytest
import torch

pytestmark = pytest.mark.gpu
H, W = 1080, 1920


@pytest.fixture(scope="module")
def L(built_lib):
    from memc_b200 import lib
    lib.load()
    return lib


def _valid_mask(flow):
    Hh, Ww = flow.shape[2:]
    xs = torch.arange(Ww, device=flow.device, dtype=torch.float32).view(1, 1, Ww)
    ys = torch.arange(Hh, device=flow.device, dtype=torch.float32).view(1, Hh, 1)
    x2, y2 = xs + flow[:, 0], ys + flow[:, 1]
    return ((x2 >= 0) & (y2 >= 0) & (x2 <= Ww - 1) & (y2 <= Hh - 1) &
            (flow[:, 0].abs() < Ww / 2.0) & (flow[:, 1].abs() < Hh / 2.0))


def test_zero_flow_centre_tap_is_identity_1080p(L):
    from my_package.modules.FilterInterpolationModule import FilterInterpolationModule
    from memc_b200 import synth
    x = synth.image(2, 3, H, W, seed=1, device="cuda")
    flow = torch.zeros(2, 2, H, W, device="cuda")
    filt = torch.zeros(2, 16, H, W, device="cuda")
    filt[:, 5] = 1.0
    assert torch.equal(FilterInterpolationModule()(x, flow, filt), x)


def test_bilinear_taps_equal_interpolation_op_1080p(L):
    """taps {5,6,9,10} = 1 turn the adaptive warp into the plain bilinear warp (interior targets)."""
    from my_package.modules.FilterInterpolationModule import FilterInterpolationModule
    from my_package.modules.InterpolationModule import InterpolationModule
    from memc_b200 import synth
    x = synth.image(1, 3, H, W, seed=2, device="cuda")
    flow = synth.smooth_flow(1, H, W, 6.0, seed=3, device="cuda")
    filt = torch.zeros(1, 16, H, W, device="cuda")
    filt[:, [5, 6, 9, 10]] = 1.0
    a = FilterInterpolationModule()(x, flow, filt)
    b = InterpolationModule()(x, flow)
    xs = torch.arange(W, device="cuda", dtype=torch.float32).view(1, 1, W)
    ys = torch.arange(H, device="cuda", dtype=torch.float32).view(1, H, 1)
    x2, y2 = xs + flow[:, 0], ys + flow[:, 1]
    inside = (x2 >= 0) & (y2 >= 0) & (x2 <= W - 2) & (y2 <= H - 2)
    assert float(inside.float().mean()) > 0.95
    assert float((a - b).abs()[inside[:, None].expand_as(a)].max()) < 1e-6


@pytest.mark.parametrize("C,B", [(3, 4), (64, 1)])
def test_filter_interpolation_adjoint_1080p(L, C, B):
    """<FI(x), g> = <x, gi1> + sum over invalid pixels of x*g  (FI is linear in x on valid pixels,
    copies x on invalid ones but gives those no gradient), and the same with (filter, gi3)."""
    from my_package.modules.FilterInterpolationModule import FilterInterpolationModule
    from memc_b200 import synth
    x, flow, filt, g = synth.filter_interpolation_case(B, C, H, W, seed=5, device="cuda")
    x.requires_grad_(), flow.requires_grad_(), filt.requires_grad_()
    out = FilterInterpolationModule()(x, flow, filt)
    gi1, gi2, gi3 = torch.autograd.grad(out, (x, flow, filt), g)
    inv = ~_valid_mask(flow.detach())
    lhs = (out.double() * g.double()).sum()
    leak = (x.detach().double() * g.double())[inv[:, None].expand_as(x)].sum()
    rhs1 = (x.detach().double() * gi1.double()).sum() + leak
    rhs3 = (filt.detach().double() * gi3.double()).sum() + leak
    scale = (out.double() * g.double()).abs().sum()
    assert abs(float((lhs - rhs1).detach())) <= 2e-6 * float(scale.detach())
    assert abs(float((lhs - rhs3).detach())) <= 2e-6 * float(scale.detach())
    assert float(gi2[inv[:, None].expand_as(gi2)].abs().max()) == 0.0 if bool(inv.any()) else True


def test_flow_projection_conservation_1080p(L):
    """count sums to 4 hits per valid source; a constant flow projects to minus itself."""
    from my_package.functions.FlowProjectionLayer import FlowProjectionLayer
    from memc_b200 import synth
    flow = synth.smooth_flow(4, H, W, 6.0, seed=7, device="cuda")
    layer = FlowProjectionLayer(requires_grad=True)
    layer(flow)
    xs = torch.arange(W, device="cuda", dtype=torch.float32).view(1, 1, W)
    ys = torch.arange(H, device="cuda", dtype=torch.float32).view(1, H, 1)
    x2, y2 = xs + flow[:, 0], ys + flow[:, 1]
    valid = (x2 >= 0) & (y2 >= 0) & (x2 <= W - 1) & (y2 <= H - 1)
    assert float(layer.count.double().sum()) == 4.0 * float(valid.sum())
    const = torch.zeros(2, 2, H, W, device="cuda")
    const[:, 0] = 3.25
    const[:, 1] = -2.5
    layer2 = FlowProjectionLayer(requires_grad=True)
    out = layer2(const)
    hit = layer2.count[:, 0] > 0
    assert float((out[:, 0][hit] + 3.25).abs().max()) < 1e-5 and float((out[:, 1][hit] - 2.5).abs().max()) < 1e-5
    assert float(out[:, 0][~hit].abs().max()) == 0.0


def test_flow_projection_fillhole_idempotent_and_complete_1080p(L):
    """After fill-hole every hole that has a counted pixel to its left/right/above is non-zero
    (for a field whose projected flow is nowhere 0) and counted pixels are untouched."""
    from my_package.functions.FlowProjectionLayer import FlowProjectionLayer
    from memc_b200 import synth
    flow = synth.tear_flow(2, H, W, 20.0, seed=9, device="cuda")
    plain = FlowProjectionLayer(requires_grad=True)
    o0 = plain(flow)
    filled = FlowProjectionLayer(requires_grad=False)
    o1 = filled(flow)
    hit = plain.count[:, 0] > 0
    # two runs: the cross-tile reduction order is not deterministic (as in the reference) -> tolerance
    assert float((o1[:, 0][hit] - o0[:, 0][hit]).abs().max()) <= 1e-4
    assert float((o1[:, 1][hit] - o0[:, 1][hit]).abs().max()) <= 1e-4
    holes = ~hit
    assert float(holes.float().mean()) > 0.02
    # rows that contain at least one counted pixel: every hole in them has a left or right neighbour
    row_has = hit.any(dim=2, keepdim=True).expand_as(hit)
    must_fill = holes & row_has
    assert float((o1[:, 0][must_fill].abs() + o1[:, 1][must_fill].abs()).min()) > 0.0


def test_interpolation_and_sepconv_adjoint_1080p(L):
    from my_package.modules.InterpolationModule import InterpolationModule
    from my_package.functions.SeparableConvLayer import SeparableConvLayer
    from memc_b200 import synth
    x, flow, _, g = synth.filter_interpolation_case(2, 3, H, W, seed=11, device="cuda")
    x.requires_grad_()
    out = InterpolationModule()(x, flow)
    (gi1,) = torch.autograd.grad(out, (x,), g)
    lhs, rhs = (out.double() * g.double()).sum(), (x.detach().double() * gi1.double()).sum()
    assert abs(float(lhs - rhs)) <= 2e-6 * float((out.double() * g.double()).abs().sum())
    fs = 4
    v = torch.randn(2, fs, H - fs + 1, W - fs + 1, device="cuda", requires_grad=True)
    hz = torch.randn(2, fs, H - fs + 1, W - fs + 1, device="cuda", requires_grad=True)
    x2 = synth.image(2, 3, H, W, seed=12, device="cuda").requires_grad_()
    o = SeparableConvLayer(fs)(x2, v, hz)
    go = torch.randn_like(o)
    g1, g2, g3 = torch.autograd.grad(o, (x2, v, hz), go)
    lhs = (o.double() * go.double()).sum()
    sc = float((o.double() * go.double()).abs().sum())
    for t, gt in ((x2, g1), (v, g2), (hz, g3)):   # the op is linear in each argument separately
        assert abs(float(lhs - (t.detach().double() * gt.double()).sum())) <= 2e-6 * sc
