"""GPU parity tests proper (run on the B200 box):  libmemc_b200.so, called through the C ABI
(reference-named launchers, extended entry points, and the autograd Functions / Modules that
sit on them) against

  * the CPU oracle (oracle/memc_oracle.c, pinned bit-exactly to the reference's my_lib.c),
    f32 and f64 builds, on the same seeded inputs, incl. the edge cases of tests/cases.py;
  * the reference's OWN CUDA kernels recompiled for sm_100a (oracle/_ref/libmemc_ref_gpu.so),
    when that binary travelled with the snapshot.

Tolerance (north_star): <= 1e-5 max-abs fp32 against the reference's kernels, written as
TOL below.  Outputs that are sums of float atomics (gradinput1, FlowProjection output) are
order-dependent even in the reference; they are compared with the f64 oracle using
TOL * max(1, max|expected|) and `count` (integer-valued) exactly.
"""
import numpy as np
import pytest
import torch

from oracle import cpu, ref
from tests.cases import fi_case, flow_case, sepconv_case

pytestmark = pytest.mark.gpu
TOL = 1e-5


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    torch.cuda.synchronize()
    return t.detach().cpu().numpy()


def close(got, exp, tol=TOL, what=""):
    got = host(got) if isinstance(got, torch.Tensor) else got
    exp = np.asarray(exp, dtype=np.float64)
    scale = max(1.0, float(np.abs(exp).max())) if exp.size else 1.0
    err = float(np.abs(got.astype(np.float64) - exp).max()) if exp.size else 0.0
    assert err <= tol * scale, "%s: max-abs err %.3e > %.1e * %.3g" % (what, err, tol, scale)
    return err


@pytest.fixture(scope="module")
def L(built_lib):
    from memc_b200 import lib
    lib.load()
    assert torch.cuda.is_available()
    return lib


FI_SHAPES = [  # B, C, H, W, fs, sigma
    (1, 3, 64, 64, 4, 3.0), (3, 3, 64, 64, 4, 3.0), (2, 3, 37, 53, 4, 8.0), (1, 5, 20, 31, 5, 2.0),
    (1, 2, 16, 16, 2, 1.0), (1, 3, 24, 24, 6, 40.0), (1, 64, 16, 24, 4, 2.0), (1, 1, 1, 1, 4, 0.0),
    (2, 3, 96, 128, 4, 4.0), (1, 3, 128, 256, 4, 20.0), (1, 64, 64, 128, 4, 3.0), (1, 3, 70, 260, 4, 1.0),
    (2, 3, 37, 100, 4, 2.0), (1, 4, 50, 196, 4, 6.0), (1, 1, 33, 96, 4, 1.5), (1, 2, 130, 132, 4, 60.0),
    (2, 6, 40, 100, 4, 3.0), (1, 9, 70, 132, 4, 12.0),   # C > 4: channel-chunked forward (last chunk ragged)
]


# ------------------------------------------------------------------- FilterInterpolation
@pytest.mark.parametrize("shape", FI_SHAPES)
@pytest.mark.parametrize("no_fast", [False, True])
def test_filter_interpolation_modules_vs_oracle(L, shape, no_fast, monkeypatch):
    """Through the public API (Module -> Function -> extended C ABI, OVERWRITE mode)."""
    from my_package.modules.FilterInterpolationModule import FilterInterpolationModule
    if no_fast:  # force the generic kernels to cross-check the fast path
        monkeypatch.setattr(L, "OVERWRITE", L.OVERWRITE | L.NO_FAST)
    B, C, H, W, fs, sigma = shape
    in1, flow, filt, gout = fi_case(B, C, H, W, fs, sigma, seed=sum(shape[:5]))
    t1, t2, t3 = dev(in1).requires_grad_(), dev(flow).requires_grad_(), dev(filt).requires_grad_()
    out = FilterInterpolationModule()(t1, t2, t3)
    close(out, cpu.filter_interpolation_forward(in1, flow, filt), what="out vs f32 oracle")
    close(out, cpu.filter_interpolation_forward(in1, flow, filt, "f64"), what="out vs f64 oracle")
    g1, g2, g3 = torch.autograd.grad(out, (t1, t2, t3), dev(gout))
    e1, e2, e3 = cpu.filter_interpolation_backward(in1, flow, filt, gout, "f64")
    close(g1, e1, what="gradinput1"), close(g2, e2, what="gradinput2"), close(g3, e3, what="gradinput3")


@pytest.mark.parametrize("shape", [(2, 3, 37, 53, 4, 8.0), (1, 5, 20, 31, 5, 2.0), (2, 3, 96, 128, 4, 4.0)])
def test_filter_interpolation_named_abi_reference_contract(L, shape):
    """Reference-named FFI functions: caller zero-fills; gi1/gi3 are ADDED into, gi2 is
    ASSIGNED for valid pixels only (my_lib_kernel.cu:1283-1286, 1424, 1495)."""
    import my_package._ext.my_lib as my_lib
    B, C, H, W, fs, sigma = shape
    in1, flow, filt, gout = fi_case(B, C, H, W, fs, sigma, seed=5)
    t1, t2, t3, tg = dev(in1), dev(flow), dev(filt), dev(gout)
    out = torch.zeros_like(t1)
    assert my_lib.FilterInterpolationLayer_gpu_forward(t1, t2, t3, out) == 0
    close(out, cpu.filter_interpolation_forward(in1, flow, filt, "f64"), what="out")
    e1, e2, e3 = cpu.filter_interpolation_backward(in1, flow, filt, gout, "f64")
    for prefill in (0.0, 1.0):
        g1, g2, g3 = (torch.full_like(t, prefill) for t in (t1, t2, t3))
        assert my_lib.FilterInterpolationLayer_gpu_backward(t1, t2, t3, tg, g1, g2, g3) == 0
        close(g1, e1 + prefill, what="gi1 (+=)")
        close(g3, e3 + prefill, what="gi3 (+=)")
        # invalid pixels keep the prefill in gi2: e2 is 0 there, valid ones are overwritten
        x2 = np.arange(W, dtype=np.float32)[None, None, :] + flow[:, 0]
        y2 = np.arange(H, dtype=np.float32)[None, :, None] + flow[:, 1]
        with np.errstate(invalid="ignore"):
            valid = ((x2 >= 0) & (y2 >= 0) & (x2 <= W - 1) & (y2 <= H - 1) &
                     (np.abs(flow[:, 0]) < np.float32(W) / 2) & (np.abs(flow[:, 1]) < np.float32(H) / 2))
        exp2 = np.where(valid[:, None], e2, prefill)
        close(g2, exp2, what="gi2 (= for valid only)")


def test_filter_interpolation_strided_views_and_stream(L):
    """b/c-strided views (the reference accepts them, my_lib_cuda.c:624-646) on a side stream."""
    import my_package._ext.my_lib as my_lib
    B, C, H, W = 2, 3, 40, 48
    in1, flow, filt, gout = fi_case(B, C, H, W, 4, 3.0, seed=17)
    big1 = torch.zeros(B + 1, C + 2, H, W, device="cuda")
    bigo = torch.zeros_like(big1)
    v1, vo = big1[1:, 1:1 + C], bigo[1:, 1:1 + C]
    v1.copy_(dev(in1))
    bigf = torch.zeros(B, 20, H, W, device="cuda")
    vf = bigf[:, 2:18]
    vf.copy_(dev(filt))
    t2 = dev(flow)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        assert my_lib.FilterInterpolationLayer_gpu_forward(v1, t2, vf, vo) == 0
    s.synchronize()
    close(vo, cpu.filter_interpolation_forward(in1, flow, filt, "f64"), what="strided out")
    assert float(bigo[0].abs().max()) == 0.0 and float(bigo[:, 0].abs().max()) == 0.0  # nothing spilled


def test_filter_interpolation_fast_path_on_strided_views(L):
    """Batch/channel-sliced views at a TMA-eligible size: the tensor maps must honour the real
    strides of every tensor (inputs, gradoutput, gradinput1) -- extended ABI, fwd + bwd."""
    B, C, H, W = 2, 3, 64, 128
    in1, flow, filt, gout = fi_case(B, C, H, W, 4, 3.0, seed=19)

    def view(a, pad_b, pad_c, off_c):
        big = torch.full((a.shape[0] + pad_b, a.shape[1] + pad_c, H, W), 7.0, device="cuda")
        v = big[pad_b:, off_c:off_c + a.shape[1]]
        v.copy_(dev(a))
        return big, v

    _, v1 = view(in1, 1, 2, 1)
    _, v2 = view(flow, 0, 3, 2)
    _, v3 = view(filt, 1, 4, 3)
    _, vg = view(gout, 2, 1, 0)
    bo, vo = view(np.zeros_like(in1), 1, 2, 2)
    b1, g1 = view(np.zeros_like(in1), 1, 1, 1)
    b2, g2 = view(np.zeros_like(flow), 1, 1, 0)
    b3, g3 = view(np.zeros_like(filt), 0, 2, 1)
    S, P = L.strides_of, L.ptr
    L.call("memc_b200_filter_interpolation_forward", L.stream_ptr(v1), B, C, H, W, 4, S(v1), S(v2), S(v3), S(vo),
           P(v1), P(v2), P(v3), P(vo), L.OVERWRITE)
    L.call("memc_b200_filter_interpolation_backward", L.stream_ptr(v1), B, C, H, W, 4, S(v1), S(v2), S(v3), S(vg),
           S(g1), S(g2), S(g3), P(v1), P(v2), P(v3), P(vg), P(g1), P(g2), P(g3), L.OVERWRITE)
    close(vo, cpu.filter_interpolation_forward(in1, flow, filt, "f64"), what="strided fast out")
    e1, e2, e3 = cpu.filter_interpolation_backward(in1, flow, filt, gout, "f64")
    close(g1, e1, what="strided fast gi1"), close(g2, e2, what="strided fast gi2"), close(g3, e3, what="strided fast gi3")
    for big, v in ((bo, vo), (b1, g1), (b2, g2), (b3, g3)):   # nothing written outside the views
        v.fill_(7.0)
        assert bool((big == 7.0).all())


@pytest.mark.skipif(not ref.available_gpu(), reason="oracle/_ref/libmemc_ref_gpu.so not present")
@pytest.mark.parametrize("shape", [(2, 3, 96, 128, 4, 4.0), (1, 64, 64, 128, 4, 3.0), (1, 5, 20, 31, 5, 2.0),
                                   (1, 3, 128, 256, 4, 20.0)])
def test_filter_interpolation_vs_reference_cuda_kernels(L, shape):
    """north_star bar: <= 1e-5 max-abs fp32 vs the reference's own CUDA kernels, same inputs."""
    import my_package._ext.my_lib as my_lib
    B, C, H, W, fs, sigma = shape
    in1, flow, filt, gout = fi_case(B, C, H, W, fs, sigma, seed=23)
    t1, t2, t3, tg = dev(in1), dev(flow), dev(filt), dev(gout)
    r_out = ref.gpu_filter_interpolation_forward(t1, t2, t3)
    out = torch.zeros_like(t1)
    assert my_lib.FilterInterpolationLayer_gpu_forward(t1, t2, t3, out) == 0
    close(out, host(r_out), what="out vs reference CUDA")
    r1, r2, r3 = ref.gpu_filter_interpolation_backward(t1, t2, t3, tg)
    g1, g2, g3 = torch.zeros_like(t1), torch.zeros_like(t2), torch.zeros_like(t3)
    assert my_lib.FilterInterpolationLayer_gpu_backward(t1, t2, t3, tg, g1, g2, g3) == 0
    close(g1, host(r1), what="gi1 vs reference CUDA")
    close(g2, host(r2), what="gi2 vs reference CUDA")
    close(g3, host(r3), what="gi3 vs reference CUDA")


@pytest.mark.parametrize("shape", [(2, 3, 96, 128, 3.0), (1, 3, 270, 480, 6.0), (1, 3, 200, 324, 30.0), (2, 4, 64, 96, 1.0),
                                   (1, 64, 96, 160, 4.0), (2, 7, 64, 128, 2.0), (1, 5, 40, 100, 12.0)])
def test_filter_interpolation_fast_path_equals_generic_bitwise(L, shape):
    """The TMA path stages data differently but performs the SAME fp32 operations in the same
    order as the generic kernel, so the two must agree bit for bit (forward)."""
    from memc_b200 import synth
    B, C, H, W, sigma = shape
    t1, t2, t3, _ = synth.filter_interpolation_case(B, C, H, W, sigma=sigma, seed=9, device="cuda")
    outs = []
    # production kernel, the row-segment / TMA-staged-output variants (MEMC_B200_VARIANT), the generic kernel
    for flags in (L.OVERWRITE, L.OVERWRITE | L.variant(1), L.OVERWRITE | L.variant(2), L.OVERWRITE | L.NO_FAST):
        o = torch.empty_like(t1)
        L.call("memc_b200_filter_interpolation_forward", L.stream_ptr(t1), B, C, H, W, 4, L.strides_of(t1),
               L.strides_of(t2), L.strides_of(t3), L.strides_of(o), L.ptr(t1), L.ptr(t2), L.ptr(t3), L.ptr(o), flags)
        outs.append(o)
    torch.cuda.synchronize()
    for o in (outs[1:-1] if C > 4 else outs[:-1]):
        assert torch.equal(o, outs[-1])
    # C > 4 production = (pixel, tap column) lanes: the partial sums of a pixel meet in shuffles and the taps are
    # pre-multiplied with the bilinear weights, so it differs from the generic kernel by rounding order only
    if C > 4:
        assert float((outs[0] - outs[-1]).abs().max()) <= 2e-6


@pytest.mark.parametrize("shape", [(2, 3, 96, 128, 3.0), (1, 3, 270, 480, 6.0), (1, 3, 200, 324, 30.0), (2, 4, 64, 96, 1.0),
                                   (1, 1, 33, 96, 1.5), (1, 2, 130, 132, 60.0), (3, 3, 64, 64, 3.0), (1, 3, 8, 16, 1.0),
                                   (1, 3, 1080, 1920, 6.0)])
@pytest.mark.parametrize("overwrite", [True, False])
def test_filter_interpolation_backward_kernels_agree(L, shape, overwrite):
    """The production backward ((pixel, tap row) lanes), the round-1 kernel (MEMC_B200_VARIANT(1)) and the generic
    kernel on the same input, OVERWRITE and the reference's += contract (prefilled gradients): gradinput2/3
    differ only by the summation order of <= 16 x C products, gradinput1 by the fixed-point rounding."""
    from memc_b200 import synth
    B, C, H, W, sigma = shape
    t1, t2, t3, tg = synth.filter_interpolation_case(B, C, H, W, sigma=sigma, seed=13, device="cuda")
    base = L.OVERWRITE if overwrite else 0
    res = []
    for flags in (base, base | L.variant(1), base | L.NO_FAST):
        if overwrite:
            g1, g2, g3 = torch.full_like(t1, 7.0), torch.full_like(t2, 7.0), torch.full_like(t3, 7.0)  # garbage
        else:
            g1, g2, g3 = torch.full_like(t1, 0.5), torch.full_like(t2, 0.5), torch.full_like(t3, 0.5)
        L.call("memc_b200_filter_interpolation_backward", L.stream_ptr(t1), B, C, H, W, 4, L.strides_of(t1),
               L.strides_of(t2), L.strides_of(t3), L.strides_of(tg), L.strides_of(g1), L.strides_of(g2),
               L.strides_of(g3), L.ptr(t1), L.ptr(t2), L.ptr(t3), L.ptr(tg), L.ptr(g1), L.ptr(g2), L.ptr(g3), flags)
        res.append((g1, g2, g3))
    torch.cuda.synchronize()
    for k, name in enumerate(("gradinput1", "gradinput2", "gradinput3")):
        gen = res[2][k]
        scale = max(1.0, float(gen.abs().max()))
        for which, r in (("rows", res[0]), ("round-1", res[1])):
            err = float((r[k] - gen).abs().max())
            assert err <= TOL * scale, "%s %s: %.3e vs generic (scale %.3g)" % (which, name, err, scale)


@pytest.mark.parametrize("shape", [(1, 8, 96, 128, 3.0), (2, 16, 64, 160, 6.0), (1, 64, 130, 200, 4.0), (1, 12, 200, 324, 30.0),
                                   (1, 8, 33, 96, 1.5), (1, 32, 8, 72, 1.0), (1, 8, 270, 480, 60.0)])
@pytest.mark.parametrize("overwrite", [True, False])
def test_filter_interpolation_backward_channel_chunks(L, shape, overwrite):
    """C > 4 (C % 4 == 0): the channel-chunked tap-row backward (filter_interpolation_bwd_chunked.cu) against the
    generic kernel and the fp64 oracle, OVERWRITE and the reference's += contract; large sigma = windows that leave
    the staged box (per-tap path), tiny frames = box clamped into the image."""
    from memc_b200 import synth
    B, C, H, W, sigma = shape
    t1, t2, t3, tg = synth.filter_interpolation_case(B, C, H, W, sigma=sigma, seed=29, device="cuda")
    base = L.OVERWRITE if overwrite else 0
    res = []
    n0 = L.launch_count()
    for flags in (base, base | L.NO_FAST):
        fill = 7.0 if overwrite else 0.5
        g1, g2, g3 = torch.full_like(t1, fill), torch.full_like(t2, fill), torch.full_like(t3, fill)
        L.call("memc_b200_filter_interpolation_backward", L.stream_ptr(t1), B, C, H, W, 4, L.strides_of(t1),
               L.strides_of(t2), L.strides_of(t3), L.strides_of(tg), L.strides_of(g1), L.strides_of(g2),
               L.strides_of(g3), L.ptr(t1), L.ptr(t2), L.ptr(t3), L.ptr(tg), L.ptr(g1), L.ptr(g2), L.ptr(g3), flags)
        res.append((g1, g2, g3))
    torch.cuda.synchronize()
    for k, name in enumerate(("gradinput1", "gradinput2", "gradinput3")):
        gen = res[1][k]
        scale = max(1.0, float(gen.abs().max()))
        err = float((res[0][k] - gen).abs().max())
        assert err <= TOL * scale, "chunked %s: %.3e vs generic (scale %.3g)" % (name, err, scale)
    if overwrite and H * W <= 130 * 200:
        e = cpu.filter_interpolation_backward(host(t1), host(t2), host(t3), host(tg), "f64")
        for got, exp, name in zip(res[0], e, ("gi1", "gi2", "gi3")):
            close(got, exp, what="chunked %s vs fp64 oracle" % name)


def test_filter_interpolation_backward_channel_chunks_nonfinite_and_float_accum(L):
    """A chunk with NaN / Inf gradients takes the fp32 path for that chunk only (same NaN / Inf masks as the generic
    kernel, which mirrors the reference arithmetic); MEMC_B200_FLOAT_ACCUM takes it for every chunk."""
    from memc_b200 import synth
    B, C, H, W = 1, 16, 96, 160
    t1, t2, t3, tg = synth.filter_interpolation_case(B, C, H, W, sigma=2.0, seed=31, device="cuda")
    tg[0, 5, 40, 70] = float("nan")
    tg[0, 9, 10, 20] = float("inf")
    res = []
    for flags in (L.OVERWRITE, L.OVERWRITE | L.FLOAT_ACCUM, L.OVERWRITE | L.NO_FAST):
        g1, g2, g3 = torch.empty_like(t1), torch.empty_like(t2), torch.empty_like(t3)
        L.call("memc_b200_filter_interpolation_backward", L.stream_ptr(t1), B, C, H, W, 4, L.strides_of(t1),
               L.strides_of(t2), L.strides_of(t3), L.strides_of(tg), L.strides_of(g1), L.strides_of(g2),
               L.strides_of(g3), L.ptr(t1), L.ptr(t2), L.ptr(t3), L.ptr(tg), L.ptr(g1), L.ptr(g2), L.ptr(g3), flags)
        res.append((g1, g2, g3))
    torch.cuda.synchronize()
    for r in res[:2]:
        for fast, gen in zip(r, res[2]):
            assert torch.equal(torch.isnan(fast), torch.isnan(gen))
            assert torch.equal(torch.isinf(fast), torch.isinf(gen))
            ok = torch.isfinite(gen)
            assert float((fast[ok] - gen[ok]).abs().max()) <= 1e-5 * max(1.0, float(gen[ok].abs().max()))
    assert bool(torch.isnan(res[0][0]).any()) and bool(torch.isinf(res[0][0]).any())


def test_filter_interpolation_backward_propagates_nonfinite_gradients(L):
    """The fast backward accumulates gradinput1 in per-tile fixed point; a tile whose gradients
    are not finite must fall back to float accumulation so NaN/Inf land where the reference
    puts them (same NaN mask as the generic kernel, which mirrors the reference arithmetic)."""
    from memc_b200 import synth
    B, C, H, W = 1, 3, 96, 160
    t1, t2, t3, tg = synth.filter_interpolation_case(B, C, H, W, sigma=2.0, seed=11, device="cuda")
    tg[0, 1, 40, 70] = float("nan")
    tg[0, 0, 10, 20] = float("inf")
    res = []
    for flags in (L.OVERWRITE, L.OVERWRITE | L.NO_FAST):
        g1, g2, g3 = torch.empty_like(t1), torch.empty_like(t2), torch.empty_like(t3)
        L.call("memc_b200_filter_interpolation_backward", L.stream_ptr(t1), B, C, H, W, 4, L.strides_of(t1),
               L.strides_of(t2), L.strides_of(t3), L.strides_of(tg), L.strides_of(g1), L.strides_of(g2),
               L.strides_of(g3), L.ptr(t1), L.ptr(t2), L.ptr(t3), L.ptr(tg), L.ptr(g1), L.ptr(g2), L.ptr(g3), flags)
        res.append((g1, g2, g3))
    torch.cuda.synchronize()
    for fast, gen in zip(*res):
        assert torch.equal(torch.isnan(fast), torch.isnan(gen))
        assert torch.equal(torch.isinf(fast), torch.isinf(gen))
        ok = torch.isfinite(gen)
        assert float((fast[ok] - gen[ok]).abs().max()) <= 1e-5
    assert bool(torch.isnan(res[0][0]).any()) and bool(torch.isinf(res[0][0]).any())


def test_host_pipeline_matches_direct_call(L):
    """memc_b200.host_pipeline (pinned host batch, frame-pipelined over streams) == direct Module call."""
    from memc_b200 import synth
    from memc_b200.host_pipeline import FilterInterpolationHostPipeline
    from my_package.modules.FilterInterpolationModule import FilterInterpolationModule
    B, C, H, W = 5, 3, 96, 160
    t = synth.filter_interpolation_case(B, C, H, W, seed=21, device="cuda")
    hs = [x.cpu().pin_memory() for x in t]
    pipe = FilterInterpolationHostPipeline("cuda:0", streams=3)
    h_out, (h1, h2, h3) = pipe.forward_backward(*hs)
    torch.cuda.synchronize()
    a, f, k = (x.clone().requires_grad_() for x in t[:3])
    o = FilterInterpolationModule()(a, f, k)
    g1, g2, g3 = torch.autograd.grad(o, (a, f, k), t[3])
    torch.cuda.synchronize()
    assert torch.equal(h_out, o.detach().cpu())
    assert float((h1 - g1.cpu()).abs().max()) <= 1e-5 and torch.equal(h2, g2.cpu()) and torch.equal(h3, g3.cpu())
    with pytest.raises(ValueError):
        pipe.forward_backward(t[0], hs[1], hs[2], hs[3])   # device tensor where a pinned host tensor is expected
    # back-to-back batches without joining in between (wait=False): distinct output buffers per batch
    t2 = synth.filter_interpolation_case(B, C, H, W, seed=22, device="cuda")
    hs2 = [x.cpu().pin_memory() for x in t2]
    o1 = pipe.alloc_outputs(hs[0], hs[1], hs[2])
    o2 = pipe.alloc_outputs(hs[0], hs[1], hs[2])
    pipe.forward_backward(*hs, outputs=o1, wait=False)
    pipe.forward_backward(*hs2, outputs=o2, wait=False)
    pipe.join()
    torch.cuda.synchronize()
    assert torch.equal(o1[0], h_out) and torch.equal(o1[2], h2) and torch.equal(o1[3], h3)
    o = FilterInterpolationModule()(t2[0], t2[1], t2[2])
    assert torch.equal(o2[0], o.cpu())


def test_filter_interpolation_backward_float_accum_flag(L):
    """MEMC_B200_FLOAT_ACCUM keeps fp32 relative precision when gradient magnitudes differ by
    many orders inside a tile (the default fixed point only bounds the ABSOLUTE error by
    ~2^-22 x the tile's largest contribution)."""
    from memc_b200 import synth
    B, C, H, W = 1, 3, 64, 96
    t1, t2, t3, tg = synth.filter_interpolation_case(B, C, H, W, sigma=1.0, seed=41, device="cuda")
    tg.mul_(1e-4)
    tg[0, :, 20, 30] = 1e4          # one huge gradient in a tile of tiny ones
    res = {}
    for name, flags in (("generic", L.OVERWRITE | L.NO_FAST), ("fixed", L.OVERWRITE), ("float", L.OVERWRITE | L.FLOAT_ACCUM)):
        g1, g2, g3 = torch.empty_like(t1), torch.empty_like(t2), torch.empty_like(t3)
        L.call("memc_b200_filter_interpolation_backward", L.stream_ptr(t1), B, C, H, W, 4, L.strides_of(t1),
               L.strides_of(t2), L.strides_of(t3), L.strides_of(tg), L.strides_of(g1), L.strides_of(g2),
               L.strides_of(g3), L.ptr(t1), L.ptr(t2), L.ptr(t3), L.ptr(tg), L.ptr(g1), L.ptr(g2), L.ptr(g3), flags)
        res[name] = g1
    torch.cuda.synchronize()
    ref = res["generic"]
    small = ref.abs() < 1e-3                      # cells that only received tiny contributions
    # relative error with a 1e-10 absolute floor: the tiny contributions are ~1e-5, so fp32 summation-order noise
    # is ~1e-12 while the fixed point's quantum here is ~1e-4 -- cells whose tiny terms cancel cannot flake the test
    rel = lambda a: float((((a - ref).abs() - 1e-10).clamp_min(0) / ref.abs().clamp_min(1e-12))[small & (ref != 0)].max())
    assert float((res["fixed"] - ref).abs().max()) <= 1e-5 * float(ref.abs().max())   # absolute bound holds
    assert rel(res["float"]) < 1e-3                                                     # relative precision kept
    assert float((res["float"] - ref).abs().max()) <= 1e-5 * float(ref.abs().max())


def test_ops_are_cuda_graph_capturable(L):
    """The C ABI only enqueues work on the caller's stream (tensor maps travel as kernel
    parameters, the FlowProjection scratch is stream-ordered), so whole steps can be captured
    in a CUDA graph and replayed -- the launch-bound small-frame regime wants exactly that."""
    from memc_b200 import synth
    B, C, H, W = 2, 3, 96, 160
    t1, t2, t3, tg = synth.filter_interpolation_case(B, C, H, W, seed=31, device="cuda")
    out, g1, g2, g3 = torch.empty_like(t1), torch.empty_like(t1), torch.empty_like(t2), torch.empty_like(t3)
    count, proj = torch.empty(B, 1, H, W, device="cuda"), torch.empty_like(t2)
    S, P = L.strides_of, L.ptr

    def step():
        st = L.stream_ptr(t1)
        L.call("memc_b200_filter_interpolation_forward", st, B, C, H, W, 4, S(t1), S(t2), S(t3), S(out),
               P(t1), P(t2), P(t3), P(out), L.OVERWRITE)
        L.call("memc_b200_filter_interpolation_backward", st, B, C, H, W, 4, S(t1), S(t2), S(t3), S(tg),
               S(g1), S(g2), S(g3), P(t1), P(t2), P(t3), P(tg), P(g1), P(g2), P(g3), L.OVERWRITE)
        L.call("memc_b200_flow_projection_forward", st, B, H, W, 1, S(t2), S(count), S(proj),
               P(t2), P(count), P(proj), L.OVERWRITE)

    step()
    torch.cuda.synchronize()
    eager = [x.clone() for x in (out, g1, g2, g3, count, proj)]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        step()                       # warm-up on the capture stream (function attributes, pool)
        side.synchronize()
        with torch.cuda.graph(graph, stream=side):
            step()
    for x in (out, g1, g2, g3, count, proj):
        x.fill_(float("nan"))
    graph.replay()
    torch.cuda.synchronize()
    for got, exp, name in zip((out, g1, g2, g3, count, proj), eager, ("out", "gi1", "gi2", "gi3", "count", "proj")):
        assert float((got - exp).abs().max()) <= 1e-5 * max(1.0, float(exp.abs().max())), name


def test_filter_interpolation_720p_vs_oracle(L):
    """One full 1280x720 frame (BASELINE.json configs[1] geometry) against the f64 oracle."""
    from my_package.modules.FilterInterpolationModule import FilterInterpolationModule
    from memc_b200 import synth
    B, C, H, W = 1, 3, 720, 1280
    t1, t2, t3, tg = synth.filter_interpolation_case(B, C, H, W, seed=3, device="cuda")
    t1.requires_grad_(), t2.requires_grad_(), t3.requires_grad_()
    out = FilterInterpolationModule()(t1, t2, t3)
    g1, g2, g3 = torch.autograd.grad(out, (t1, t2, t3), tg)
    in1, flow, filt, gout = host(t1), host(t2), host(t3), host(tg)
    close(out, cpu.filter_interpolation_forward(in1, flow, filt, "f64"), what="720p out")
    e1, e2, e3 = cpu.filter_interpolation_backward(in1, flow, filt, gout, "f64")
    close(g1, e1, what="720p gi1"), close(g2, e2, what="720p gi2"), close(g3, e3, what="720p gi3")


# ------------------------------------------------------ fused call site
@pytest.mark.parametrize("shape", [(2, 3, 96, 160, 4, 4.0), (1, 3, 270, 480, 4, 6.0), (1, 3, 37, 53, 4, 8.0),
                                   (1, 3, 70, 260, 4, 30.0), (1, 5, 20, 31, 5, 2.0), (1, 64, 64, 128, 4, 3.0)])
def test_fused_filter_interpolate_equals_composition(L, shape):
    """memc_b200.fused.FilterInterpolate (one kernel: two warps + occlusion blend) vs the reference's
    composition of two FilterInterpolationModule calls and the fp32 blend (networks/MEMC_Net.py:258-264):
    forward bit-identical; gradients of all eight inputs equal those of autograd through the composition."""
    from memc_b200 import fused
    from my_package.modules.FilterInterpolationModule import FilterInterpolationModule
    B, C, H, W, fs, sigma = shape
    g = torch.Generator().manual_seed(7)
    cases = [fi_case(B, C, H, W, fs, sigma, seed=70 + k) for k in range(2)]

    def leaves():
        refs = [dev(c[0]).requires_grad_() for c in cases]
        offs = [dev(c[1]).requires_grad_() for c in cases]
        filts = [dev(c[2]).requires_grad_() for c in cases]
        occs = [(0.5 + 0.3 * torch.randn(B, 1, H, W, generator=torch.Generator().manual_seed(9 + k))).cuda().requires_grad_()
                for k in range(2)]
        return refs, offs, filts, occs

    gout = torch.randn(B, C, H, W, generator=g).cuda()
    r, o, f, oc = leaves()
    fused_out = fused.FilterInterpolate(r[0], r[1], o, f, oc, fs * fs)
    fused_grads = torch.autograd.grad(fused_out, r + o + f + oc, gout)
    r2, o2, f2, oc2 = leaves()
    comp = oc2[0] * FilterInterpolationModule()(r2[0], o2[0], f2[0]) + oc2[1] * FilterInterpolationModule()(r2[1], o2[1], f2[1])
    comp_grads = torch.autograd.grad(comp, r2 + o2 + f2 + oc2, gout)
    assert torch.equal(fused_out, comp), "fused forward must be bit-identical to the composition"
    for a, b, name in zip(fused_grads, comp_grads, ["ref0", "ref2", "off0", "off1", "filt0", "filt1", "occ0", "occ1"]):
        close(a, host(b), tol=2e-5, what="fused grad " + name)


@pytest.mark.parametrize("shape", [(2, 3, 96, 128, 4, 4.0), (1, 3, 70, 132, 4, 10.0), (1, 5, 40, 100, 5, 3.0)])
@pytest.mark.parametrize("mean", [False, True])
def test_fused_call_sites_vs_oracle_composition(L, shape, mean):
    """The fused call sites against the ORACLE's composition (not this repo's own Modules): the C oracle (pinned to
    the reference's my_lib.c) warps both references, numpy blends them like networks/MEMC_Net.py:258-264
    (occlusion weights) or networks/MEMC_Net_s.py:258-264 (plain mean); gradients of the references / flows / filters
    against the oracle's backward fed with the blended gradient."""
    from memc_b200 import fused
    B, C, H, W, fs, sigma = shape
    cases = [fi_case(B, C, H, W, fs, sigma, seed=80 + k) for k in range(2)]
    rng = np.random.default_rng(5)
    occs = [(0.5 + 0.3 * rng.standard_normal((B, 1, H, W))).astype(np.float32) for _ in range(2)]
    gout = rng.standard_normal((B, C, H, W)).astype(np.float32)
    r = [dev(c[0]).requires_grad_() for c in cases]
    o = [dev(c[1]).requires_grad_() for c in cases]
    f = [dev(c[2]).requires_grad_() for c in cases]
    if mean:
        out = fused.FilterInterpolateMean(r[0], r[1], o, f, fs * fs)
        wts = [np.float32(0.5), np.float32(0.5)]
    else:
        out = fused.FilterInterpolate(r[0], r[1], o, f, [dev(x) for x in occs], fs * fs)
        wts = occs
    warps = [cpu.filter_interpolation_forward(c[0], c[1], c[2], "f64") for c in cases]
    close(out, wts[0] * warps[0] + wts[1] * warps[1], what="fused forward vs oracle composition")
    grads = torch.autograd.grad(out, r + o + f, dev(gout))
    for k in range(2):
        e1, e2, e3 = cpu.filter_interpolation_backward(cases[k][0], cases[k][1], cases[k][2], (gout * wts[k]).astype(np.float32), "f64")
        close(grads[k], e1, tol=2e-5, what="grad ref%d" % k)
        close(grads[2 + k], e2, tol=2e-5, what="grad offset%d" % k)
        close(grads[4 + k], e3, tol=2e-5, what="grad filter%d" % k)


@pytest.mark.parametrize("shape", [(2, 3, 64, 64, 128, 3.0), (1, 3, 16, 96, 160, 8.0), (1, 3, 5, 70, 132, 2.0), (1, 3, 64, 37, 100, 2.0)])
def test_fused_shared_flow_pair_vs_oracle(L, shape):
    """fused.FilterInterpolateShared (the RGB frame and its context features warped with ONE flow / filter in one kernel,
    networks/MEMC_Net_star.py:272-285) against the oracle's two separate warps, forward and backward; the last shape has
    an odd H: the library composes the two plain calls itself."""
    from memc_b200 import fused
    B, Ca, Cb, H, W, sigma = shape
    in_a, flow, filt, gout_a = fi_case(B, Ca, H, W, 4, sigma, seed=90)
    rng = np.random.default_rng(6)
    in_b = rng.standard_normal((B, Cb, H, W)).astype(np.float32)
    gout_b = rng.standard_normal((B, Cb, H, W)).astype(np.float32)
    ta, tb, tf, tk = (dev(x).requires_grad_() for x in (in_a, in_b, flow, filt))
    oa, ob = fused.FilterInterpolateShared(ta, tb, tf, tk)
    close(oa, cpu.filter_interpolation_forward(in_a, flow, filt, "f64"), what="shared pair: first image")
    close(ob, cpu.filter_interpolation_forward(in_b, flow, filt, "f64"), what="shared pair: second image")
    ga, gb, gf, gk = torch.autograd.grad((oa, ob), (ta, tb, tf, tk), (dev(gout_a), dev(gout_b)))
    a1, a2, a3 = cpu.filter_interpolation_backward(in_a, flow, filt, gout_a, "f64")
    b1, b2, b3 = cpu.filter_interpolation_backward(in_b, flow, filt, gout_b, "f64")
    close(ga, a1, tol=2e-5, what="grad first image"), close(gb, b1, tol=2e-5, what="grad second image")
    close(gf, a2 + b2, tol=2e-5, what="grad flow"), close(gk, a3 + b3, tol=2e-5, what="grad filter")


@pytest.mark.parametrize("fillhole", [0, 1])
def test_fused_flow_project_pair_vs_oracle(L, fillhole):
    """fused.FlowProjectPair against the oracle run on each direction separately (count exact, fill-hole incl.)."""
    from memc_b200 import fused, synth
    B, H, W = 2, 96, 160
    fa, fb = synth.smooth_flow(B, H, W, 5.0, seed=1, device="cuda"), synth.tear_flow(B, H, W, 8.0, seed=2, device="cuda")
    if fillhole:
        with torch.no_grad():
            pa, pb = fused.FlowProjectPair(fa, fb)
    else:
        pa, pb = fused.FlowProjectPair(fa.requires_grad_(), fb.requires_grad_())
    for got, src in ((pa, fa), (pb, fb)):
        eo, _ = cpu.flow_projection_forward(host(src), fillhole, "f64")
        close(got, eo, what="FlowProjectPair vs oracle")
    # mixed requires_grad: the reference decides fill-hole per direction
    fa2, fb2 = fa.detach().clone().requires_grad_(), fb.detach().clone()
    qa, qb = fused.FlowProjectPair(fa2, fb2)
    close(qa, cpu.flow_projection_forward(host(fa2), 0, "f64")[0], what="pair: direction with grad is not hole-filled")
    close(qb, cpu.flow_projection_forward(host(fb2), 1, "f64")[0], what="pair: direction without grad is hole-filled")


@pytest.mark.parametrize("B", [1, 3])
def test_fused_flow_project_pair_equals_two_calls(L, B):
    """memc_b200.fused.FlowProjectPair == two FlowProjectionModule calls (networks/MEMC_Net.py:109-113): counts
    equal, outputs equal up to the fp32 summation order, gradients flow to both inputs."""
    from memc_b200 import fused, synth
    from my_package.modules.FlowProjectionModule import FlowProjectionModule
    H, W = 96, 160
    fa, fb = synth.smooth_flow(B, H, W, 5.0, seed=1, device="cuda"), synth.tear_flow(B, H, W, 8.0, seed=2, device="cuda")
    with torch.no_grad():
        pa, pb = fused.FlowProjectPair(fa, fb)
        qa, qb = FlowProjectionModule(False)(fa), FlowProjectionModule(False)(fb)
    close(pa, host(qa), what="pair a"), close(pb, host(qb), what="pair b")
    ga, gb = fa.clone().requires_grad_(), fb.clone().requires_grad_()
    pa, pb = fused.FlowProjectPair(ga, gb)
    (pa.sum() + 2.0 * pb.sum()).backward()
    ha, hb = fa.clone().requires_grad_(), fb.clone().requires_grad_()
    (FlowProjectionModule(True)(ha).sum() + 2.0 * FlowProjectionModule(True)(hb).sum()).backward()
    close(ga.grad, host(ha.grad), what="pair grad a"), close(gb.grad, host(hb.grad), what="pair grad b")


# ------------------------------------------------------------------------ FlowProjection
FP_SHAPES = [(1, 64, 64, 3.0), (2, 37, 53, 8.0), (1, 24, 24, 40.0), (1, 1, 1, 0.0), (2, 96, 160, 6.0), (1, 70, 260, 1.0)]


@pytest.mark.parametrize("shape", FP_SHAPES)
@pytest.mark.parametrize("fillhole", [0, 1])
def test_flow_projection_vs_oracle(L, shape, fillhole):
    from my_package.functions.FlowProjectionLayer import FlowProjectionLayer
    B, H, W, sigma = shape
    flow = flow_case(B, H, W, sigma, seed=31)
    layer = FlowProjectionLayer(requires_grad=not fillhole)
    t = dev(flow).requires_grad_(not fillhole)
    out = layer(t)
    eo, ec = cpu.flow_projection_forward(flow, fillhole, "f64")
    assert np.array_equal(host(layer.count), ec), "count must be exact"
    close(out, eo, what="FlowProjection out")
    if not fillhole:
        gout = np.random.default_rng(3).standard_normal(flow.shape).astype(np.float32)
        (gi,) = torch.autograd.grad(out, (t,), dev(gout))
        close(gi, cpu.flow_projection_backward(flow, ec, gout, "f64"), what="FlowProjection gi")


def test_flow_projection_named_abi_reference_contract(L):
    import my_package._ext.my_lib as my_lib
    B, H, W = 2, 48, 64
    flow = flow_case(B, H, W, 5.0, seed=37)
    t = dev(flow)
    for fillhole in (0, 1):
        count, out = torch.zeros(B, 1, H, W, device="cuda"), torch.zeros_like(t)
        assert my_lib.FlowProjectionLayer_gpu_forward(t, count, out, fillhole) == 0
        eo, ec = cpu.flow_projection_forward(flow, fillhole, "f64")
        assert np.array_equal(host(count), ec)
        close(out, eo, what="named FlowProjection fwd fillhole=%d" % fillhole)
    gout = np.random.default_rng(4).standard_normal(flow.shape).astype(np.float32)
    for prefill in (0.0, 1.0):  # the reference kernel does `gradinput1[...] += ...` (my_lib_kernel.cu:1879)
        gi = torch.full_like(t, prefill)
        assert my_lib.FlowProjectionLayer_gpu_backward(t, count, dev(gout), gi) == 0
        e = cpu.flow_projection_backward(flow, ec, gout, "f64")
        x2 = np.arange(W, dtype=np.float32)[None, None, :] + flow[:, 0]
        y2 = np.arange(H, dtype=np.float32)[None, :, None] + flow[:, 1]
        with np.errstate(invalid="ignore"):
            valid = (x2 >= 0) & (y2 >= 0) & (x2 <= W - 1) & (y2 <= H - 1)
        close(gi, np.where(valid[:, None], e + prefill, prefill), what="named FlowProjection bwd")


@pytest.mark.parametrize("kind", ["contention", "divergent", "uniform", "tear"])
def test_flow_projection_regimes(L, kind):
    """The three BASELINE.json configs[2] regimes at a size the oracle finishes quickly."""
    from my_package.functions.FlowProjectionLayer import FlowProjectionLayer
    from memc_b200 import synth
    B, H, W = 2, 270, 480
    if kind == "contention":
        t = synth.radial_flow(B, H, W, 0.9, device="cuda")
    elif kind == "divergent":
        t = synth.radial_flow(B, H, W, -1.5, device="cuda")   # targets 2.5 px apart: lattice of holes
    elif kind == "tear":
        t = synth.tear_flow(B, H, W, 20.0, seed=4, device="cuda")
    else:
        t = synth.uniform_flow(B, H, W, 32.0, seed=2, device="cuda")
    layer = FlowProjectionLayer(requires_grad=False)
    out = layer(t)
    eo, ec = cpu.flow_projection_forward(host(t), 1, "f64")
    assert np.array_equal(host(layer.count), ec)
    # thousands of addends of magnitude ~100 px land on one cell in the contention case: the
    # fp32 atomic order noise scales with sum|addend| * eps, so bound by TOL * max|sum|/count...
    close(out, eo, tol=5e-5 if kind == "contention" else TOL, what=kind)


@pytest.mark.skipif(not ref.available_gpu(), reason="oracle/_ref/libmemc_ref_gpu.so not present")
@pytest.mark.parametrize("fillhole", [0, 1])
def test_flow_projection_vs_reference_cuda_kernels(L, fillhole):
    import my_package._ext.my_lib as my_lib
    from memc_b200 import synth
    B, H, W = 2, 180, 320
    for t in (synth.smooth_flow(B, H, W, 6.0, seed=1, device="cuda"), synth.radial_flow(B, H, W, -1.5, device="cuda"),
              synth.tear_flow(B, H, W, 12.0, seed=3, device="cuda")):
        r_out, r_count = ref.gpu_flow_projection_forward(t, fillhole)
        count, out = torch.zeros(B, 1, H, W, device="cuda"), torch.zeros_like(t)
        assert my_lib.FlowProjectionLayer_gpu_forward(t, count, out, fillhole) == 0
        assert torch.equal(count, r_count)
        close(out, host(r_out), what="FlowProjection vs reference CUDA (fillhole=%d)" % fillhole)
        gout = torch.randn_like(t)
        gi = torch.zeros_like(t)
        assert my_lib.FlowProjectionLayer_gpu_backward(t, count, gout, gi) == 0
        close(gi, host(ref.gpu_flow_projection_backward(t, r_count, gout)), what="FlowProjection bwd vs reference CUDA")


# ------------------------------------------------------------------- DepthFlowProjection
def _inverse_depth(shape, seed, kind="iid"):
    """iid: every pixel its own depth in [0.5, 20] (weights span 40:1 inside every tile: the splat's direct path);
    smooth: memc_b200.synth.inverse_depth (slowly varying + object edges: the fixed-point box for most tiles)."""
    if kind == "smooth":
        from memc_b200 import synth
        B, _, H, W = shape
        return synth.inverse_depth(B, H, W, seed=seed).numpy()
    rng = np.random.default_rng(seed)
    return (1e-6 + 1.0 / rng.uniform(0.5, 20.0, shape)).astype(np.float32)


@pytest.mark.parametrize("shape", FP_SHAPES + [(2, 270, 480, 6.0)])
@pytest.mark.parametrize("fillhole", [0, 1])
@pytest.mark.parametrize("no_fast", [False, True])
@pytest.mark.parametrize("weights", ["iid", "smooth"])
def test_depth_flow_projection_vs_oracle(L, shape, fillhole, no_fast, weights):
    """Function / Module surface (fast shared-memory splat when the frame is at least one box large) and the generic
    kernels (MEMC_B200_NO_FAST) against the fp64 oracle, forward (+ fill-hole) and backward for both inputs."""
    from my_package.modules.DepthFlowProjectionModule import DepthFlowProjectionModule
    B, H, W, sigma = shape
    flow = flow_case(B, H, W, sigma, seed=41)
    depth = _inverse_depth((B, 1, H, W), 43, weights)
    eo, ec = cpu.depth_flow_projection_forward(flow, depth, fillhole, "f64")
    if no_fast:
        S, P = L.strides_of, L.ptr
        t, d = dev(flow), dev(depth)
        count, out = torch.full((B, 1, H, W), 7.0, device="cuda"), torch.full_like(t, -3.0)  # OVERWRITE ignores it
        L.call("memc_b200_depth_flow_projection_forward", L.stream_ptr(t), B, H, W, fillhole, S(t), S(d), S(count), S(out),
               P(t), P(d), P(count), P(out), L.OVERWRITE | L.NO_FAST)
        close(count, ec, what="DepthFlowProjection count (generic)")
        close(out, eo, what="DepthFlowProjection out (generic)")
        return
    mod = DepthFlowProjectionModule(requires_grad=not fillhole)
    t, d = dev(flow).requires_grad_(not fillhole), dev(depth).requires_grad_(not fillhole)
    out = mod(t, d)
    close(mod.f.count, ec, what="DepthFlowProjection count")
    close(out, eo, what="DepthFlowProjection out")
    if not fillhole:
        gout = np.random.default_rng(3).standard_normal(flow.shape).astype(np.float32)
        g1, g2 = torch.autograd.grad(out, (t, d), dev(gout))
        e1, e2 = cpu.depth_flow_projection_backward(flow, depth, host(mod.f.count), host(out), gout, "f64")
        close(g1, e1, what="DepthFlowProjection gi1")
        close(g2, e2, what="DepthFlowProjection gi2")


def test_depth_flow_projection_named_abi_reference_contract(L):
    """my_lib_cuda.h:101-117: caller-zeroed count / output, gradients accumulated with += for valid pixels only."""
    import my_package._ext.my_lib as my_lib
    B, H, W = 2, 48, 128
    flow = flow_case(B, H, W, 5.0, seed=47)
    depth = _inverse_depth((B, 1, H, W), 49)
    t, d = dev(flow), dev(depth)
    for fillhole in (0, 1):
        count, out = torch.zeros(B, 1, H, W, device="cuda"), torch.zeros_like(t)
        assert my_lib.DepthFlowProjectionLayer_gpu_forward(t, d, count, out, fillhole) == 0
        eo, ec = cpu.depth_flow_projection_forward(flow, depth, fillhole, "f64")
        close(count, ec, what="named DepthFlowProjection count")
        close(out, eo, what="named DepthFlowProjection fwd fillhole=%d" % fillhole)
    gout = np.random.default_rng(4).standard_normal(flow.shape).astype(np.float32)
    x2 = np.arange(W, dtype=np.float32)[None, None, :] + flow[:, 0]
    y2 = np.arange(H, dtype=np.float32)[None, :, None] + flow[:, 1]
    with np.errstate(invalid="ignore"):
        valid = ((x2 >= 0) & (y2 >= 0) & (x2 <= W - 1) & (y2 <= H - 1))[:, None]
    e1, e2 = cpu.depth_flow_projection_backward(flow, depth, host(count), host(out), gout, "f64")
    for prefill in (0.0, 1.0):
        g1, g2 = torch.full_like(t, prefill), torch.full_like(d, prefill)
        assert my_lib.DepthFlowProjectionLayer_gpu_backward(t, d, count, out, dev(gout), g1, g2) == 0
        close(g1, np.where(valid, e1 + prefill, prefill), what="named DepthFlowProjection gi1")
        close(g2, np.where(valid, e2 + prefill, prefill), what="named DepthFlowProjection gi2")
    assert my_lib.DepthFlowProjectionLayer_gpu_forward(t, dev(np.concatenate([depth, depth], 1)), count, out, 0) == -1
    assert my_lib.DepthFlowProjectionLayer_gpu_forward(dev(np.concatenate([flow, flow], 1)), d, count, out, 0) == -1


@pytest.mark.parametrize("kind", ["contention", "divergent", "uniform", "tear", "wide_weights", "nonfinite_weight"])
def test_depth_flow_projection_regimes(L, kind):
    """The FlowProjection regimes with weights, plus the tiles the fixed point must hand to the direct path: weights
    spanning 2^30 inside a tile, and a NaN / Inf weight (which must land exactly where the generic kernel puts it)."""
    from memc_b200 import synth
    S, P = L.strides_of, L.ptr
    B, H, W = 2, 270, 480
    if kind == "contention":
        t = synth.radial_flow(B, H, W, 0.9, device="cuda")
    elif kind == "divergent":
        t = synth.radial_flow(B, H, W, -1.5, device="cuda")
    elif kind == "tear":
        t = synth.tear_flow(B, H, W, 20.0, seed=4, device="cuda")
    elif kind == "uniform":
        t = synth.uniform_flow(B, H, W, 32.0, seed=2, device="cuda")
    else:
        t = synth.smooth_flow(B, H, W, 5.0, seed=6, device="cuda")
    depth = _inverse_depth((B, 1, H, W), 53, "smooth")
    if kind == "wide_weights":
        depth[:, :, ::7, ::5] *= np.float32(2.0 ** -30)
    if kind == "nonfinite_weight":
        depth[0, 0, 100, 200] = np.nan
        depth[1, 0, 30, 40] = np.inf
    d = dev(depth)
    res = []
    for flags in (L.OVERWRITE, L.OVERWRITE | L.NO_FAST):
        count, out = torch.empty(B, 1, H, W, device="cuda"), torch.empty_like(t)
        L.call("memc_b200_depth_flow_projection_forward", L.stream_ptr(t), B, H, W, 1, S(t), S(d), S(count), S(out),
               P(t), P(d), P(count), P(out), flags)
        res.append((count, out))
    torch.cuda.synchronize()
    if kind == "nonfinite_weight":
        for fast, gen in zip(res[0], res[1]):
            assert torch.equal(torch.isnan(fast), torch.isnan(gen)) and torch.equal(torch.isinf(fast), torch.isinf(gen))
            ok = torch.isfinite(gen)
            assert float((fast[ok] - gen[ok]).abs().max()) <= 1e-5 * max(1.0, float(gen[ok].abs().max()))
        assert bool(torch.isnan(res[0][1]).any())
        return
    eo, ec = cpu.depth_flow_projection_forward(host(t), depth, 1, "f64")
    tol = 5e-5 if kind == "contention" else TOL
    for (count, out), name in zip(res, ("fast", "generic")):
        close(count, ec, tol=tol, what="%s %s count" % (kind, name))
        close(out, eo, tol=tol, what="%s %s out" % (kind, name))


@pytest.mark.skipif(not ref.available_gpu(), reason="oracle/_ref/libmemc_ref_gpu.so not present")
@pytest.mark.parametrize("fillhole", [0, 1])
def test_depth_flow_projection_vs_reference_cuda_kernels(L, fillhole):
    import my_package._ext.my_lib as my_lib
    from memc_b200 import synth
    B, H, W = 2, 180, 320
    d = dev(_inverse_depth((B, 1, H, W), 59, "smooth"))
    for t in (synth.smooth_flow(B, H, W, 6.0, seed=1, device="cuda"), synth.radial_flow(B, H, W, -1.5, device="cuda"),
              synth.tear_flow(B, H, W, 12.0, seed=3, device="cuda")):
        r_out, r_count = ref.gpu_depth_flow_projection_forward(t, d, fillhole)
        count, out = torch.zeros(B, 1, H, W, device="cuda"), torch.zeros_like(t)
        assert my_lib.DepthFlowProjectionLayer_gpu_forward(t, d, count, out, fillhole) == 0
        close(count, host(r_count), what="DepthFlowProjection count vs reference CUDA")
        close(out, host(r_out), what="DepthFlowProjection vs reference CUDA (fillhole=%d)" % fillhole)
        gout = torch.randn_like(t)
        g1, g2 = torch.zeros_like(t), torch.zeros_like(d)
        assert my_lib.DepthFlowProjectionLayer_gpu_backward(t, d, r_count, r_out, gout, g1, g2) == 0
        r1, r2 = ref.gpu_depth_flow_projection_backward(t, d, r_count, r_out, gout)
        close(g1, host(r1), what="DepthFlowProjection gi1 vs reference CUDA")
        close(g2, host(r2), what="DepthFlowProjection gi2 vs reference CUDA")


# ---------------------------------------------------------------- WeightedFlowProjection
def _frames_and_threshold(flow, seed, quantile=0.5):
    """Two random frames and a gate threshold that no source sits on: the middle of the widest gap between neighbouring
    brightness errors around the given quantile (the fp32 error of the CUDA source and of the C source can differ in the
    last bit, my_lib.c:1960 vs my_lib_kernel.cu:2575; the tests must not depend on which side of the gate that lands)."""
    B, _, H, W = flow.shape
    rng = np.random.default_rng(seed)
    im0, im1 = rng.random((B, 3, H, W), dtype=np.float32), rng.random((B, 3, H, W), dtype=np.float32)
    xs, ys = np.arange(W, dtype=np.float32)[None, None, :], np.arange(H, dtype=np.float32)[None, :, None]
    with np.errstate(invalid="ignore"):   # NaN / Inf flows (tests/cases.py edge cases) never reach the gate: any cell will do
        x3 = np.clip(np.nan_to_num(xs + np.float32(2) * flow[:, 0], nan=0.0), 0, W - 1).astype(np.int64)
        y3 = np.clip(np.nan_to_num(ys + np.float32(2) * flow[:, 1], nan=0.0), 0, H - 1).astype(np.int64)
    bi = np.arange(B)[:, None, None]
    err = sum(np.abs(im0[:, c].astype(np.float64) - im1[bi, c, y3, x3]) for c in range(3)) / 3.0
    e = np.sort(err[np.isfinite(err)].ravel())
    if len(e) < 2:
        return im0, im1, 1.0
    lo, hi = int(len(e) * max(0.0, quantile - 0.1)), min(len(e), max(int(len(e) * min(1.0, quantile + 0.1)), 2))
    lo = min(lo, hi - 2)
    k = lo + int(np.argmax(np.diff(e[lo:hi])))
    return im0, im1, float(0.5 * (e[k] + e[k + 1]))


@pytest.mark.parametrize("shape", FP_SHAPES + [(2, 270, 480, 6.0)])
@pytest.mark.parametrize("fillhole", [0, 1])
@pytest.mark.parametrize("no_fast", [False, True])
def test_weighted_flow_projection_vs_oracle(L, shape, fillhole, no_fast):
    from my_package.modules.WeightedFlowProjectionModule import WeightedFlowProjectionModule
    B, H, W, sigma = shape
    flow = flow_case(B, H, W, sigma, seed=61)
    im0, im1, thr = _frames_and_threshold(flow, 63)
    eo, ec, ew = cpu.weighted_flow_projection_forward(flow, im0, im1, fillhole, thr, "f64")
    t, a, b = dev(flow), dev(im0), dev(im1)
    if no_fast:
        S, P = L.strides_of, L.ptr
        count, weight = torch.full((B, 1, H, W), 7.0, device="cuda"), torch.full((B, 1, H, W), 5.0, device="cuda")
        out = torch.full_like(t, -3.0)  # OVERWRITE ignores what is there
        L.call("memc_b200_weighted_flow_projection_forward", L.stream_ptr(t), B, H, W, fillhole, thr, S(t), S(a), S(b),
               S(count), S(weight), S(out), P(t), P(a), P(b), P(count), P(weight), P(out), L.OVERWRITE | L.NO_FAST)
        assert np.array_equal(host(count), ec), "count must be exact"
        close(out, eo, what="WeightedFlowProjection out (generic)")
        close(weight, ew, what="WeightedFlowProjection weight (generic)")
        return
    mod = WeightedFlowProjectionModule(requires_grad=not fillhole, threshold=thr)
    t.requires_grad_(not fillhole)
    out = mod(t, a, b)
    assert np.array_equal(host(mod.f.count), ec), "count must be exact"
    assert 0 < ec.sum() < 4 * B * H * W or H * W == 1
    close(out, eo, what="WeightedFlowProjection out")
    close(mod.f.weight, ew, what="WeightedFlowProjection weight")
    if not fillhole:
        gout = np.random.default_rng(3).standard_normal(flow.shape).astype(np.float32)
        (gi,) = torch.autograd.grad(out, (t,), dev(gout))
        close(gi, cpu.weighted_flow_projection_backward(flow, im0, im1, ec, gout, thr, "f64"), what="WeightedFlowProjection gi")


def test_weighted_flow_projection_named_abi_reference_contract(L):
    """my_lib_cuda.h:119-139: caller-zeroed count / weight / output, gradinput accumulated with += for voting pixels only;
    thresholds that admit nobody and everybody."""
    import my_package._ext.my_lib as my_lib
    B, H, W = 2, 48, 128
    flow = flow_case(B, H, W, 5.0, seed=67)
    im0, im1, thr = _frames_and_threshold(flow, 69, 0.4)
    t, a, b = dev(flow), dev(im0), dev(im1)
    for fillhole in (0, 1):
        count, weight, out = torch.zeros(B, 1, H, W, device="cuda"), torch.zeros(B, 1, H, W, device="cuda"), torch.zeros_like(t)
        assert my_lib.WeightedFlowProjectionLayer_gpu_forward(t, a, b, count, weight, out, fillhole, thr) == 0
        eo, ec, ew = cpu.weighted_flow_projection_forward(flow, im0, im1, fillhole, thr, "f64")
        assert np.array_equal(host(count), ec)
        close(out, eo, what="named WeightedFlowProjection fwd fillhole=%d" % fillhole)
        close(weight, ew, what="named WeightedFlowProjection weight")
    gout = np.random.default_rng(4).standard_normal(flow.shape).astype(np.float32)
    e = cpu.weighted_flow_projection_backward(flow, im0, im1, ec, gout, thr, "f64")
    for prefill in (0.0, 1.0):
        gi = torch.full_like(t, prefill)
        assert my_lib.WeightedFlowProjectionLayer_gpu_backward(t, a, b, count, weight, dev(gout), gi, thr) == 0
        close(gi, e + prefill, what="named WeightedFlowProjection bwd")   # non-voting pixels: e == 0 there
    for thr2, want in ((0.0, 0.0), (10.0, None)):
        count, weight, out = torch.zeros(B, 1, H, W, device="cuda"), torch.zeros(B, 1, H, W, device="cuda"), torch.zeros_like(t)
        assert my_lib.WeightedFlowProjectionLayer_gpu_forward(t, a, b, count, weight, out, 0, thr2) == 0
        if want is not None:
            assert float(count.abs().max()) == want and float(out.abs().max()) == want
        else:   # everybody votes: FlowProjection's count and output
            fo, fc = cpu.flow_projection_forward(flow, 0, "f64")
            assert np.array_equal(host(count), fc)
            close(out, fo, what="threshold above every error = FlowProjection")
    assert my_lib.WeightedFlowProjectionLayer_gpu_forward(t, a[:, :2], b, count, weight, out, 0, thr) == -1


@pytest.mark.skipif(not ref.available_gpu(), reason="oracle/_ref/libmemc_ref_gpu.so not present")
@pytest.mark.parametrize("fillhole", [0, 1])
def test_weighted_flow_projection_vs_reference_cuda_kernels(L, fillhole):
    import my_package._ext.my_lib as my_lib
    from memc_b200 import synth
    B, H, W = 2, 180, 320
    for t in (synth.smooth_flow(B, H, W, 6.0, seed=1, device="cuda"), synth.radial_flow(B, H, W, -1.5, device="cuda"),
              synth.radial_flow(B, H, W, 0.9, device="cuda")):
        im0, im1, thr = _frames_and_threshold(host(t), 71)
        a, b = dev(im0), dev(im1)
        r_out, r_count, r_weight = ref.gpu_weighted_flow_projection_forward(t, a, b, fillhole, thr)
        count, weight, out = torch.zeros(B, 1, H, W, device="cuda"), torch.zeros(B, 1, H, W, device="cuda"), torch.zeros_like(t)
        assert my_lib.WeightedFlowProjectionLayer_gpu_forward(t, a, b, count, weight, out, fillhole, thr) == 0
        assert torch.equal(count, r_count)
        close(out, host(r_out), tol=5e-5, what="WeightedFlowProjection vs reference CUDA (fillhole=%d)" % fillhole)
        close(weight, host(r_weight), what="WeightedFlowProjection weight vs reference CUDA")
        gout = torch.randn_like(t)
        gi = torch.zeros_like(t)
        assert my_lib.WeightedFlowProjectionLayer_gpu_backward(t, a, b, count, weight, gout, gi, thr) == 0
        close(gi, host(ref.gpu_weighted_flow_projection_backward(t, a, b, r_count, r_weight, gout, thr)),
              what="WeightedFlowProjection bwd vs reference CUDA")


# ------------------------------------------------------------------------------------- WeightLayer
def _mostly_close(got, exp, what, frac=5e-4):
    """The backward takes 27 x C sign decisions per pixel on fp32 values whose last bit depends on FMA contraction; a near-tie
    flips one contribution.  All but a fraction `frac` of the elements must agree to TOL; none may be off by more than a
    few flipped contributions."""
    got = host(got) if isinstance(got, torch.Tensor) else got
    exp = np.asarray(exp, np.float64)
    scale = max(1.0, float(np.abs(exp).max()))
    d = np.abs(got.astype(np.float64) - exp)
    off = float((d > TOL * scale).mean())
    assert off <= frac, "%s: %.2e of the elements differ by more than %.1e" % (what, off, TOL * scale)
    return off


@pytest.mark.parametrize("shape", FP_SHAPES)
def test_weight_layer_vs_oracle_and_reference_cuda(L, shape):
    import my_package._ext.my_lib as my_lib
    S, P = L.strides_of, L.ptr
    B, H, W, sigma = shape
    flow = flow_case(B, H, W, sigma, seed=79)
    rng = np.random.default_rng(81)
    in1, in2 = rng.random((B, 3, H, W), dtype=np.float32), rng.random((B, 3, H, W), dtype=np.float32)
    lam = 0.9
    t1, t2, t3 = dev(in1), dev(in2), dev(flow)
    e = cpu.weight_layer_forward(in1, in2, flow, lam, 3.0, "f64")
    out = torch.full((B, 1, H, W), 7.0, device="cuda")
    assert my_lib.WeightLayer_gpu_forward(t1, t2, t3, out, lam, 0.0, 3.0) == 0
    close(out, e, what="WeightLayer forward")
    assert my_lib.WeightLayer_gpu_forward(t1, t2, t3, out, lam, 0.0, 5.0) == -1   # Nw must be 3
    o2 = torch.empty_like(out)
    L.call("memc_b200_weight_layer_forward", L.stream_ptr(t1), B, 3, H, W, lam, 3.0, S(t1), S(t2), S(t3), S(o2),
           P(t1), P(t2), P(t3), P(o2), L.OVERWRITE)
    assert torch.equal(o2, out)
    fout = host(out)
    gout = rng.standard_normal(fout.shape).astype(np.float32)
    tg = dev(gout)
    e1, e2, e3 = cpu.weight_layer_backward(in1, in2, flow, fout, gout, lam, 3.0, "f64")
    g1, g2, g3 = torch.full_like(t1, 7.0), torch.full_like(t2, 7.0), torch.full_like(t3, 7.0)
    L.call("memc_b200_weight_layer_backward", L.stream_ptr(t1), B, 3, H, W, lam, 3.0, S(t1), S(t2), S(t3), S(out),
           P(t1), P(t2), P(t3), P(out), P(tg), P(g1), P(g2), P(g3), L.OVERWRITE)
    _mostly_close(g1, e1, "WeightLayer gi1"), _mostly_close(g2, e2, "WeightLayer gi2"), _mostly_close(g3, e3, "WeightLayer gi3")
    h1, h2, h3 = torch.ones_like(t1), torch.ones_like(t2), torch.ones_like(t3)   # reference contract: accumulated into
    assert my_lib.WeightLayer_gpu_backward(t1, t2, t3, out, tg, h1, h2, h3, lam, 0.0, 3.0) == 0
    _mostly_close(h1, e1 + 1.0, "WeightLayer gi1 (named, +=)"), _mostly_close(h2, e2 + 1.0, "WeightLayer gi2 (named, +=)")
    _mostly_close(h3, e3 + 1.0, "WeightLayer gi3 (named, +=)")
    if ref.available_gpu():
        r = ref.gpu_weight_layer_forward(t1, t2, t3, lam)
        close(out, host(r), what="WeightLayer forward vs reference CUDA")
        r1, r2, r3 = ref.gpu_weight_layer_backward(t1, t2, t3, r, tg, lam)
        g1, g2, g3 = torch.zeros_like(t1), torch.zeros_like(t2), torch.zeros_like(t3)
        assert my_lib.WeightLayer_gpu_backward(t1, t2, t3, r, tg, g1, g2, g3, lam, 0.0, 3.0) == 0
        _mostly_close(g1, host(r1), "WeightLayer gi1 vs reference CUDA"), _mostly_close(g2, host(r2), "WeightLayer gi2 vs reference CUDA")
        _mostly_close(g3, host(r3), "WeightLayer gi3 vs reference CUDA")


# ------------------------------------------------------------------------------- SeparableConvFlow
@pytest.mark.parametrize("shape", [(1, 3, 32, 32, 4), (2, 3, 21, 35, 5), (1, 3, 9, 9, 3), (1, 3, 4, 4, 4), (2, 3, 64, 96, 4)])
def test_separable_conv_flow_vs_oracle_and_reference_cuda(L, shape):
    """Signed tap sums (the CUDA source's semantics, which the oracle follows), a zero-sum filter (-2000, no gradient),
    the reference contract (gradinput2 assigned, gradinput3 accumulated) and OVERWRITE; vs the reference kernels too."""
    import my_package._ext.my_lib as my_lib
    S, P = L.strides_of, L.ptr
    B, C, H, W, fs = shape
    in1, v, hz, _ = sepconv_case(B, C, H, W, fs, seed=39)
    v[0, :, 0, 0] = 0.0
    Ho, Wo = H - fs + 1, W - fs + 1
    t1, tv, th = dev(in1), dev(v), dev(hz)
    e = cpu.separable_conv_flow_forward(v, hz, "f64")
    flow = torch.full((B, 2, Ho, Wo), 7.0, device="cuda")
    assert my_lib.SeparableConvFlowLayer_gpu_forward(t1, tv, th, flow) == 0
    assert float(flow[0, 1, 0, 0]) == -2000.0
    # a centroid divides by the tap sum, and signed taps cancel: compare where the sum is well conditioned (|sum| > 0.5
    # against sum|tap| ~ 3: the fp32 rounding of the sum itself is amplified by sum|tap| / |sum| elsewhere)
    ok = np.concatenate([(np.abs(hz.sum(1)) > 0.5)[:, None], (np.abs(v.sum(1)) > 0.5)[:, None]], 1)
    got = host(flow)
    assert np.abs(np.where(ok, got - e, 0)).max() <= 1e-5 * max(1.0, np.abs(np.where(ok, e, 0)).max())
    f2 = torch.empty_like(flow)
    L.call("memc_b200_separable_conv_flow_forward", L.stream_ptr(tv), B, H, W, fs, S(tv), S(th), S(f2), P(tv), P(th), P(f2), L.OVERWRITE)
    assert torch.equal(f2, flow)
    gflow = np.random.default_rng(9).standard_normal((B, 2, Ho, Wo)).astype(np.float32)
    tg = dev(gflow)
    ev, eh = cpu.separable_conv_flow_backward(v, hz, gflow, "f64")
    okv, okh = np.broadcast_to(ok[:, 1:2], v.shape), np.broadcast_to(ok[:, 0:1], hz.shape)
    gv, gh = torch.full_like(tv, 7.0), torch.full_like(th, 7.0)
    L.call("memc_b200_separable_conv_flow_backward", L.stream_ptr(tv), B, H, W, fs, S(tv), S(th), S(tg), S(gv), S(gh),
           P(tv), P(th), P(tg), P(gv), P(gh), L.OVERWRITE)
    for got, exp, m in ((host(gv), ev, okv), (host(gh), eh, okh)):
        assert np.abs(np.where(m, got - exp, 0)).max() <= 1e-5 * max(1.0, np.abs(np.where(m, exp, 0)).max())
    assert float(gv[0, :, 0, 0].abs().max()) == 0.0   # zero-sum filter: no gradient (OVERWRITE writes the zeros)
    # reference contract: gradinput2 assigned, gradinput3 accumulated
    g1, hv, hh = torch.zeros_like(t1), torch.ones_like(tv), torch.ones_like(th)
    assert my_lib.SeparableConvFlowLayer_gpu_backward(t1, tv, th, tg, g1, hv, hh) == 0
    nz = dev(np.broadcast_to((v.sum(1) != 0)[:, None], v.shape).copy())
    assert torch.equal(hv[nz], gv[nz])                                      # assigned: the prefill is gone
    assert float((hh - 1.0 - gh).abs().max()) <= 1e-5 * max(1.0, float(gh.abs().max()))   # accumulated onto the prefill
    assert float(hv[0, :, 0, 0].min()) == 1.0 and float(g1.abs().max()) == 0.0   # untouched
    if ref.available_gpu():
        r = ref.gpu_separable_conv_flow_forward(t1, tv, th)
        assert np.abs(np.where(ok, host(flow) - host(r), 0)).max() <= 1e-5 * max(1.0, np.abs(np.where(ok, host(r), 0)).max())
        rv, rh = ref.gpu_separable_conv_flow_backward(t1, tv, th, tg)
        for a, b_, m in ((gv, rv, okv), (gh, rh, okh)):
            d = np.where(m, host(a) - host(b_), 0)
            assert np.abs(d).max() <= 1e-5 * max(1.0, np.abs(np.where(m, host(b_), 0)).max())


# ------------------------------------------- PixelValue / PixelWeight / ReliableWeight
def _px_case(mode, shape, seed):
    B, H, W, sigma = shape
    flow = flow_case(B, H, W, sigma, seed=seed)
    rng = np.random.default_rng(seed + 1)
    in1 = rng.random((B, 3, H, W), dtype=np.float32) if mode == "value" else None
    fw = rng.random((B, 1, H, W), dtype=np.float32) if mode != "reliable" else None
    return flow, in1, fw


@pytest.mark.parametrize("shape", FP_SHAPES)
@pytest.mark.parametrize("mode", ["value", "weight", "reliable"])
def test_pixel_splat_family_vs_oracle(L, shape, mode):
    """Extended entry points (OVERWRITE: garbage in the outputs) and the reference-named FFI entries (caller-zeroed
    outputs, gradients accumulated with +=) against the fp64 oracle, forward and backward."""
    import my_package._ext.my_lib as my_lib
    S, P = L.strides_of, L.ptr
    B, H, W, _ = shape
    flow, in1, fw = _px_case(mode, shape, 73)
    sd = 1.3
    eo = cpu.pixel_splat_forward(mode, flow, in1, fw, sd, "f64")
    t = dev(flow)
    a = dev(in1) if in1 is not None else None
    f = dev(fw) if fw is not None else None
    C = 3 if mode == "value" else 1
    out = torch.full((B, C, H, W), 7.0, device="cuda")
    st = L.stream_ptr(t)
    if mode == "value":
        L.call("memc_b200_pixel_value_forward", st, B, C, H, W, sd, S(a), S(t), S(f), S(out), P(a), P(t), P(f), P(out), L.OVERWRITE)
    elif mode == "weight":
        L.call("memc_b200_pixel_weight_forward", st, B, H, W, sd, S(t), S(f), S(out), P(t), P(f), P(out), L.OVERWRITE)
    else:
        L.call("memc_b200_reliable_weight_forward", st, B, H, W, sd, S(t), S(out), P(t), P(out), L.OVERWRITE)
    close(out, eo, what="%s forward" % mode)
    # named entry, caller-zeroed
    o2 = torch.zeros_like(out)
    if mode == "value":
        assert my_lib.PixelValueLayer_gpu_forward(a, t, f, o2, sd, 0.0, 2.0) == 0
        assert my_lib.PixelValueLayer_gpu_forward(a, t, f, o2, sd, 0.0, 3.0) == -1   # Prowindow must be 2
    elif mode == "weight":
        assert my_lib.PixelWeightLayer_gpu_forward(t, f, o2, sd, 0.0, 2.0) == 0
    else:
        assert my_lib.ReliableWeightLayer_gpu_forward(t, o2, sd, 0.0, 2.0) == 0
    close(o2, eo, what="%s forward (named)" % mode)
    # backward: the forward output handed to ours and to the oracle is the same tensor (threshold test)
    fout = host(out)
    rng = np.random.default_rng(5)
    gout = rng.standard_normal(fout.shape).astype(np.float32)
    thr = float(np.quantile(fout, 0.35)) if mode != "value" else 0.0
    e1, e3, ew = cpu.pixel_splat_backward(mode, flow, gout, in1, fw, fout, sd, thr, "f64")
    go = dev(gout)
    g3 = torch.full_like(t, 7.0)
    gw = torch.full((B, 1, H, W), 7.0, device="cuda")
    g1 = torch.full((B, C, H, W), 7.0, device="cuda")
    if mode == "value":
        L.call("memc_b200_pixel_value_backward", st, B, C, H, W, sd, S(a), S(t), S(f), S(go), S(g1), S(g3), S(gw),
               P(a), P(t), P(f), P(go), P(g1), P(g3), P(gw), L.OVERWRITE)
        close(g1, e1, what="value gi1"), close(gw, ew, what="value gfw")
    elif mode == "weight":
        L.call("memc_b200_pixel_weight_backward", st, B, H, W, sd, thr, S(t), S(f), S(out), S(go), S(g3), S(gw),
               P(t), P(f), P(out), P(go), P(g3), P(gw), L.OVERWRITE)
        close(gw, ew, what="weight gfw")
    else:
        L.call("memc_b200_reliable_weight_backward", st, B, H, W, sd, thr, S(t), S(out), S(go), S(g3),
               P(t), P(out), P(go), P(g3), L.OVERWRITE)
    close(g3, e3, what="%s gi3" % mode)
    # named entries accumulate into what the caller passes
    h3, hw, h1 = torch.ones_like(t), torch.ones(B, 1, H, W, device="cuda"), torch.ones(B, C, H, W, device="cuda")
    if mode == "value":
        assert my_lib.PixelValueLayer_gpu_backward(a, t, f, go, h1, h3, hw, sd, 0.0, 2.0) == 0
        close(h1, e1 + 1.0, what="value gi1 (named, +=)"), close(hw, ew + 1.0, what="value gfw (named, +=)")
    elif mode == "weight":
        assert my_lib.PixelWeightLayer_gpu_backward(t, f, out, go, h3, hw, thr, sd, 0.0, 2.0) == 0
        close(hw, ew + 1.0, what="weight gfw (named, +=)")
    else:
        assert my_lib.ReliableWeightLayer_gpu_backward(t, out, go, h3, thr, sd, 0.0, 2.0) == 0
    close(h3, e3 + 1.0, what="%s gi3 (named, +=)" % mode)


@pytest.mark.skipif(not ref.available_gpu(), reason="oracle/_ref/libmemc_ref_gpu.so not present")
@pytest.mark.parametrize("mode", ["value", "weight", "reliable"])
def test_pixel_splat_family_vs_reference_cuda_kernels(L, mode):
    import my_package._ext.my_lib as my_lib
    from memc_b200 import synth
    B, H, W = 2, 180, 320
    sd = 1.0
    g = torch.Generator(device="cuda").manual_seed(3)
    a = torch.rand(B, 3, H, W, device="cuda", generator=g) if mode == "value" else None
    f = torch.rand(B, 1, H, W, device="cuda", generator=g) if mode != "reliable" else None
    for t in (synth.smooth_flow(B, H, W, 6.0, seed=1, device="cuda"), synth.radial_flow(B, H, W, 1.6, device="cuda")):
        r_out = ref.gpu_pixel_splat_forward(mode, t, a, f, sd)
        out = torch.zeros_like(r_out)
        if mode == "value":
            assert my_lib.PixelValueLayer_gpu_forward(a, t, f, out, sd, 0.0, 2.0) == 0
        elif mode == "weight":
            assert my_lib.PixelWeightLayer_gpu_forward(t, f, out, sd, 0.0, 2.0) == 0
        else:
            assert my_lib.ReliableWeightLayer_gpu_forward(t, out, sd, 0.0, 2.0) == 0
        close(out, host(r_out), tol=3e-5, what="%s forward vs reference CUDA" % mode)
        gout = torch.randn_like(r_out)
        thr = float(torch.quantile(r_out.flatten()[:1000000], 0.35)) if mode != "value" else 0.0
        r1, r3, rw = ref.gpu_pixel_splat_backward(mode, t, gout, a, f, r_out, sd, thr)
        g3, gw, g1 = torch.zeros_like(t), torch.zeros(B, 1, H, W, device="cuda"), torch.zeros(B, 3, H, W, device="cuda")
        if mode == "value":
            assert my_lib.PixelValueLayer_gpu_backward(a, t, f, gout, g1, g3, gw, sd, 0.0, 2.0) == 0
            close(g1, host(r1), what="value gi1 vs reference CUDA"), close(gw, host(rw), what="value gfw vs reference CUDA")
        elif mode == "weight":
            assert my_lib.PixelWeightLayer_gpu_backward(t, f, r_out, gout, g3, gw, thr, sd, 0.0, 2.0) == 0
            close(gw, host(rw), what="weight gfw vs reference CUDA")
        else:
            assert my_lib.ReliableWeightLayer_gpu_backward(t, r_out, gout, g3, thr, sd, 0.0, 2.0) == 0
        close(g3, host(r3), what="%s gi3 vs reference CUDA" % mode)


@pytest.mark.parametrize("shape", [(7, 64, 96), (5, 70, 260), (3, 33, 3840), (1, 1100, 128), (4, 45, 2100)])
@pytest.mark.parametrize("fillhole", [0, 1])
def test_flow_projection_pipeline_many_frames(L, shape, fillhole):
    """The persistent pipeline (OVERWRITE calls): more frames than accumulator slots, ragged
    sizes, rows / columns longer than 32 mask words (two-level fill-hole search), vs the oracle."""
    from memc_b200 import synth
    S, P = L.strides_of, L.ptr
    B, H, W = shape
    for t in (synth.smooth_flow(B, H, W, 5.0, seed=11, device="cuda"), synth.tear_flow(B, H, W, 9.0, seed=5, device="cuda"),
              synth.radial_flow(B, H, W, 0.9, device="cuda")):
        count, out = torch.full((B, 1, H, W), 7.0, device="cuda"), torch.full_like(t, -3.0)  # garbage: OVERWRITE ignores it
        st = L.stream_ptr(t)
        assert L.call("memc_b200_flow_projection_forward", st, B, H, W, fillhole, S(t), S(count), S(out), P(t), P(count),
                      P(out), L.OVERWRITE) == 0
        eo, ec = cpu.flow_projection_forward(host(t), fillhole, "f64")
        assert np.array_equal(host(count), ec), "count must be exact"
        close(out, eo, tol=5e-5, what="pipeline %s fillhole=%d" % (shape, fillhole))


def test_flow_projection_pipeline_equals_per_frame_path(L):
    """Same call through the persistent pipeline and through the per-frame launches
    (MEMC_B200_VARIANT(1)): count bit-equal, output equal up to the fp32 summation order."""
    from memc_b200 import synth
    S, P = L.strides_of, L.ptr
    B, H, W = 6, 270, 480
    t = synth.smooth_flow(B, H, W, 6.0, seed=3, device="cuda")
    res = []
    for var in (0, 1):
        count, out = torch.empty(B, 1, H, W, device="cuda"), torch.empty_like(t)
        assert L.call("memc_b200_flow_projection_forward", L.stream_ptr(t), B, H, W, 1, S(t), S(count), S(out), P(t),
                      P(count), P(out), L.OVERWRITE | L.variant(var)) == 0
        res.append((count, out))
    assert torch.equal(res[0][0], res[1][0])
    assert float((res[0][1] - res[1][1]).abs().max()) <= TOL


def test_flow_projection_fillhole_follows_requires_grad(L):
    """FlowProjectionModule(input.requires_grad): holes are filled only when no grad is
    required (reference FlowProjectionLayer.py:15) -- callers use torch.no_grad()."""
    from my_package.modules.FlowProjectionModule import FlowProjectionModule
    from memc_b200 import synth
    t = synth.tear_flow(1, 64, 96, 6.0, device="cuda")
    filled = FlowProjectionModule(False)(t)
    unfilled = FlowProjectionModule(True)(t)
    eo1, _ = cpu.flow_projection_forward(host(t), 1, "f64")
    eo0, _ = cpu.flow_projection_forward(host(t), 0, "f64")
    close(filled, eo1), close(unfilled, eo0)
    assert np.abs(eo1 - eo0).max() > 0.1


# ------------------------------------------------------------------------- Interpolation
@pytest.mark.parametrize("shape", [(1, 3, 64, 64, 3.0), (2, 3, 37, 53, 8.0), (1, 7, 20, 31, 2.0), (1, 3, 1, 1, 0.0),
                                   (2, 3, 96, 160, 5.0)])
def test_interpolation_vs_oracle(L, shape):
    from my_package.modules.InterpolationModule import InterpolationModule
    B, C, H, W, sigma = shape
    in1, flow, _, gout = fi_case(B, C, H, W, 4, sigma, seed=41)
    t1, t2 = dev(in1).requires_grad_(), dev(flow).requires_grad_()
    out = InterpolationModule()(t1, t2)
    close(out, cpu.interpolation_forward(in1, flow, "f64"), what="Interpolation out")
    g1, g2 = torch.autograd.grad(out, (t1, t2), dev(gout))
    e1, e2 = cpu.interpolation_backward(in1, flow, gout, "f64")
    close(g1, e1, what="Interpolation gi1"), close(g2, e2, what="Interpolation gi2")


def test_interpolation_named_abi_and_reference_cuda(L):
    import my_package._ext.my_lib as my_lib
    B, C, H, W = 2, 3, 60, 84
    in1, flow, _, gout = fi_case(B, C, H, W, 4, 4.0, seed=43)
    t1, t2, tg = dev(in1), dev(flow), dev(gout)
    out = torch.zeros_like(t1)
    assert my_lib.InterpolationLayer_gpu_forward(t1, t2, out) == 0
    close(out, cpu.interpolation_forward(in1, flow, "f64"))
    g1, g2 = torch.zeros_like(t1), torch.zeros_like(t2)
    assert my_lib.InterpolationLayer_gpu_backward(t1, t2, tg, g1, g2) == 0
    e1, e2 = cpu.interpolation_backward(in1, flow, gout, "f64")
    close(g1, e1), close(g2, e2)
    # Ch variant: any channel count
    in5 = np.random.default_rng(1).random((B, 5, H, W), dtype=np.float32)
    o5 = torch.zeros(B, 5, H, W, device="cuda")
    assert my_lib.InterpolationLayer_gpu_forward(dev(in5), t2, o5) == -1  # C != 3 rejected (my_lib_cuda.c:373)
    assert my_lib.InterpolationChLayer_gpu_forward(dev(in5), t2, o5) == 0
    close(o5, cpu.interpolation_forward(in5, flow, "f64"))
    if ref.available_gpu():
        close(out, host(ref.gpu_interpolation_forward(t1, t2)), what="vs reference CUDA fwd")
        r1, r2 = ref.gpu_interpolation_backward(t1, t2, tg)
        close(g1, host(r1), what="vs reference CUDA gi1"), close(g2, host(r2), what="vs reference CUDA gi2")


# ------------------------------------------------------------------------- SeparableConv
@pytest.mark.parametrize("shape", [(1, 3, 32, 32, 4), (2, 3, 21, 35, 5), (1, 3, 9, 9, 3), (1, 3, 4, 4, 4), (2, 3, 64, 96, 4)])
def test_separable_conv_vs_oracle(L, shape):
    from my_package.functions.SeparableConvLayer import SeparableConvLayer
    B, C, H, W, fs = shape
    in1, v, hz, gout = sepconv_case(B, C, H, W, fs, seed=47)
    t1, t2, t3 = dev(in1).requires_grad_(), dev(v).requires_grad_(), dev(hz).requires_grad_()
    out = SeparableConvLayer(fs)(t1, t2, t3)
    close(out, cpu.separable_conv_forward(in1, v, hz, "f64"), what="SeparableConv out")
    g1, g2, g3 = torch.autograd.grad(out, (t1, t2, t3), dev(gout))
    e1, e2, e3 = cpu.separable_conv_backward(in1, v, hz, gout, "f64")
    close(g1, e1, what="SeparableConv gi1"), close(g2, e2, what="gi2"), close(g3, e3, what="gi3")


def test_separable_conv_named_abi_and_reference_cuda(L):
    import my_package._ext.my_lib as my_lib
    B, C, H, W, fs = 2, 3, 40, 56, 4
    in1, v, hz, gout = sepconv_case(B, C, H, W, fs, seed=53)
    t1, t2, t3, tg = dev(in1), dev(v), dev(hz), dev(gout)
    out = torch.zeros(B, C, H - fs + 1, W - fs + 1, device="cuda")
    assert my_lib.SeparableConvLayer_gpu_forward(t1, t2, t3, out) == 0
    close(out, cpu.separable_conv_forward(in1, v, hz, "f64"))
    e1, e2, e3 = cpu.separable_conv_backward(in1, v, hz, gout, "f64")
    for prefill in (0.0, 1.0):  # reference: atomicAdd into all three (my_lib_kernel.cu:376-381)
        g1, g2, g3 = (torch.full_like(t, prefill) for t in (t1, t2, t3))
        assert my_lib.SeparableConvLayer_gpu_backward(t1, t2, t3, tg, g1, g2, g3) == 0
        close(g1, e1 + prefill), close(g2, e2 + prefill), close(g3, e3 + prefill)
    if ref.available_gpu():
        close(out, host(ref.gpu_separable_conv_forward(t1, t2, t3)), what="vs reference CUDA fwd")
        r1, r2, r3 = ref.gpu_separable_conv_backward(t1, t2, t3, tg)
        g1, g2, g3 = (torch.zeros_like(t) for t in (t1, t2, t3))
        assert my_lib.SeparableConvLayer_gpu_backward(t1, t2, t3, tg, g1, g2, g3) == 0
        close(g1, host(r1)), close(g2, host(r2)), close(g3, host(r3))


# ------------------------------------------------------------------------------ goldens
def test_cuda_path_matches_golden_fixtures(L):
    """The committed reference-made vectors (tests/golden) through the CUDA path."""
    import glob
    import os
    import my_package._ext.my_lib as my_lib
    for path in sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))):
        z = np.load(path)
        op = str(z["op"])
        if op == "filter_interpolation":
            t1, t2, t3 = dev(z["in1"]), dev(z["flow"]), dev(z["filt"])
            out = torch.zeros_like(t1)
            assert my_lib.FilterInterpolationLayer_gpu_forward(t1, t2, t3, out) == 0
            close(out, z["out"], what=path)
            g1, g2, g3 = torch.zeros_like(t1), torch.zeros_like(t2), torch.zeros_like(t3)
            assert my_lib.FilterInterpolationLayer_gpu_backward(t1, t2, t3, dev(z["gout"]), g1, g2, g3) == 0
            close(g1, z["g1"], what=path), close(g2, z["g2"], what=path), close(g3, z["g3"], what=path)
        elif op == "flow_projection":
            t = dev(z["flow"])
            fh = int(z["fillhole"]) if "fillhole" in z.files else 0
            count, out = torch.zeros(t.shape[0], 1, *t.shape[2:], device="cuda"), torch.zeros_like(t)
            assert my_lib.FlowProjectionLayer_gpu_forward(t, count, out, fh) == 0
            assert np.array_equal(host(count), z["count"])
            close(out, z["out"], what=path)
        elif op == "depth_flow_projection":
            t, d = dev(z["flow"]), dev(z["depth"])
            count, out = torch.zeros(t.shape[0], 1, *t.shape[2:], device="cuda"), torch.zeros_like(t)
            assert my_lib.DepthFlowProjectionLayer_gpu_forward(t, d, count, out, 0) == 0
            close(count, z["count"], what=path), close(out, z["out"], what=path)
            g1, g2 = torch.zeros_like(t), torch.zeros_like(d)
            assert my_lib.DepthFlowProjectionLayer_gpu_backward(t, d, dev(z["count"]), dev(z["out"]), dev(z["gout"]), g1, g2) == 0
            ok = np.isfinite(z["g1"])   # a source whose cell has no accumulated weight divides by 0 in the reference too
            assert np.array_equal(np.isfinite(host(g1)), ok)
            close(np.where(ok, host(g1), 0), np.where(ok, z["g1"], 0), what=path)
            ok2 = np.isfinite(z["g2"])
            close(np.where(ok2, host(g2), 0), np.where(ok2, z["g2"], 0), what=path)
        elif op == "weighted_flow_projection":
            t, a, b, thr = dev(z["flow"]), dev(z["im0"]), dev(z["im1"]), float(z["threshold"])
            count, weight = torch.zeros(t.shape[0], 1, *t.shape[2:], device="cuda"), torch.zeros(t.shape[0], 1, *t.shape[2:], device="cuda")
            out = torch.zeros_like(t)
            assert my_lib.WeightedFlowProjectionLayer_gpu_forward(t, a, b, count, weight, out, 0, thr) == 0
            # (a source exactly on the gate could land on either side: fixtures were checked to have none within 1e-6)
            assert np.array_equal(host(count), z["count"])
            close(out, z["out"], what=path), close(weight, z["weight"], what=path)
            gi = torch.zeros_like(t)
            assert my_lib.WeightedFlowProjectionLayer_gpu_backward(t, a, b, count, weight, dev(z["gout"]), gi, thr) == 0
            close(gi, z["gi"], what=path)
        elif op == "pixel_splat":
            mode, sd, thr = str(z["mode"]), float(z["sigma_d"]), float(z["threshold"])
            t, out, go = dev(z["flow"]), torch.zeros(*z["out"].shape, device="cuda"), dev(z["gout"])
            g3 = torch.zeros_like(t)
            if mode == "value":
                a, f = dev(z["in1"]), dev(z["fw"])
                g1, gw = torch.zeros_like(a), torch.zeros_like(f)
                assert my_lib.PixelValueLayer_gpu_forward(a, t, f, out, sd, 0.0, 2.0) == 0
                assert my_lib.PixelValueLayer_gpu_backward(a, t, f, go, g1, g3, gw, sd, 0.0, 2.0) == 0
                close(g1, z["g1"], what=path), close(gw, z["gw"], what=path)
            elif mode == "weight":
                f = dev(z["fw"])
                gw = torch.zeros_like(f)
                assert my_lib.PixelWeightLayer_gpu_forward(t, f, out, sd, 0.0, 2.0) == 0
                assert my_lib.PixelWeightLayer_gpu_backward(t, f, dev(z["out"]), go, g3, gw, thr, sd, 0.0, 2.0) == 0
                close(gw, z["gw"], what=path)
            else:
                assert my_lib.ReliableWeightLayer_gpu_forward(t, out, sd, 0.0, 2.0) == 0
                assert my_lib.ReliableWeightLayer_gpu_backward(t, dev(z["out"]), go, g3, thr, sd, 0.0, 2.0) == 0
            close(out, z["out"], what=path), close(g3, z["g3"], what=path)
        elif op == "weight_layer":
            t1, t2, t3, lam = dev(z["in1"]), dev(z["in2"]), dev(z["flow"]), float(z["lambda_e"])
            out = torch.zeros(*z["out"].shape, device="cuda")
            assert my_lib.WeightLayer_gpu_forward(t1, t2, t3, out, lam, 0.0, 3.0) == 0
            close(out, z["out"], what=path)
            g1, g2, g3 = torch.zeros_like(t1), torch.zeros_like(t2), torch.zeros_like(t3)
            assert my_lib.WeightLayer_gpu_backward(t1, t2, t3, dev(z["out"]), dev(z["gout"]), g1, g2, g3, lam, 0.0, 3.0) == 0
            for got, k in ((g1, "g1"), (g2, "g2"), (g3, "g3")):   # a near-tie may flip a sign: all but a few elements
                d = np.abs(host(got) - z[k])
                assert float((d > 1e-5 * max(1.0, np.abs(z[k]).max())).mean()) <= 5e-3, (path, k)
        elif op == "separable_conv_flow":
            t1, tv, th = dev(z["in1"]), dev(z["vert"]), dev(z["horiz"])
            flow = torch.zeros(*z["flow"].shape, device="cuda")
            assert my_lib.SeparableConvFlowLayer_gpu_forward(t1, tv, th, flow) == 0
            close(flow, z["flow"], what=path)
            g1, gv, gh = torch.zeros_like(t1), torch.zeros_like(tv), torch.zeros_like(th)
            assert my_lib.SeparableConvFlowLayer_gpu_backward(t1, tv, th, dev(z["gflow"]), g1, gv, gh) == 0
            close(gv, z["gv"], what=path), close(gh, z["gh"], what=path)
        elif op == "interpolation":
            t1, t2 = dev(z["in1"]), dev(z["flow"])
            out = torch.zeros_like(t1)
            fn = my_lib.InterpolationLayer_gpu_forward if t1.shape[1] == 3 else my_lib.InterpolationChLayer_gpu_forward
            assert fn(t1, t2, out) == 0
            close(out, z["out"], what=path)
        elif op == "separable_conv":
            t1, t2, t3 = dev(z["in1"]), dev(z["vert"]), dev(z["horiz"])
            out = torch.zeros(*z["out"].shape, device="cuda")
            assert my_lib.SeparableConvLayer_gpu_forward(t1, t2, t3, out) == 0
            close(out, z["out"], what=path)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_ops_run_on_the_device_that_owns_the_tensors(L):
    """tensors on cuda:1 while cuda:0 is current -- the library makes the owning device current for the
    call (DeviceGuard in runtime.cu) and restores the caller's; results equal the same call made on cuda:0."""
    from memc_b200 import synth
    from my_package.modules.FilterInterpolationModule import FilterInterpolationModule
    from my_package.modules.FlowProjectionModule import FlowProjectionModule
    torch.cuda.set_device(0)
    in1, flow, filt, gout = synth.filter_interpolation_case(2, 3, 96, 128, seed=3, device="cuda:0")
    outs = []
    for d in ("cuda:0", "cuda:1"):
        t1, t2, t3 = (t.to(d).requires_grad_() for t in (in1, flow, filt))
        o = FilterInterpolationModule()(t1, t2, t3)
        g = torch.autograd.grad(o, (t1, t2, t3), gout.to(d))
        with torch.no_grad():
            pr = FlowProjectionModule(False)(torch.cat([flow.to(d)] * 2, 0))   # B = 4: persistent pipeline
        torch.cuda.synchronize(d)
        assert torch.cuda.current_device() == 0
        outs.append([x.cpu() for x in (o.detach(), *g, pr)])
    for a, b in zip(*outs):
        assert float((a - b).abs().max()) <= TOL
    with pytest.raises(L.MemcB200Error):
        FilterInterpolationModule()(in1, flow.to("cuda:1"), filt)
