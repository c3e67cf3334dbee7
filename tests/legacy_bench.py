"""Ours vs the reference's legacy kernels recompiled for sm_100a, same box, same inputs, same
harness (NOT a pytest module; test infrastructure because it drives oracle/_ref):

    python tests/legacy_bench.py [--iters 10] [--out gpurun_out/legacy_bench.json] [--only ops,net]

  ops   every hot-path op at the BASELINE.json sizes
  net   the op SEQUENCE one frame pair costs inside the reference's networks (call pattern of
        networks/MEMC_Net.py:109-125,252-264 and networks/MEMC_Net_star.py:127-140,272-285), on
        the frame sizes the demos pad to (demo_HD720p.py:90-108): 2 x FlowProjection (fill-hole on),
        2 x FilterInterpolation C=3 + the occlusion blend, and for MEMC_Net_star 2 more
        FilterInterpolation calls on the 64-channel context features.  The conv sub-networks
        between them are cuDNN work outside the hot path and are not part of the measurement.

CUDA-event median after 3 warm-ups, 256 MB L2 flush between iterations; the legacy arm includes the
zero fills its contract needs (functions/*.py allocate zeroed outputs).
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "memc-net_b200")):
    sys.path.insert(0, p)

from memc_b200 import lib, synth  # noqa: E402
from oracle import ref  # noqa: E402
from tools.kbench import P, S, timeit  # noqa: E402


def ours_fi(in1, flow, filt, out):
    B, C, H, W = in1.shape
    lib.call("memc_b200_filter_interpolation_forward", lib.stream_ptr(in1), B, C, H, W, 4, S(in1), S(flow), S(filt), S(out),
             P(in1), P(flow), P(filt), P(out), lib.OVERWRITE)


def ours_fp(flow, count, out):
    B, _, H, W = flow.shape
    lib.call("memc_b200_flow_projection_forward", lib.stream_ptr(flow), B, H, W, 1, S(flow), S(count), S(out), P(flow),
             P(count), P(out), lib.OVERWRITE)


def ours_blend(refs, flows, filts, occs, out):
    B, C, H, W = refs[0].shape
    lib.call("memc_b200_filter_interpolation_blend_forward", lib.stream_ptr(out), B, C, H, W, 4,
             S(refs[0]), S(flows[0]), S(filts[0]), S(refs[1]), S(flows[1]), S(filts[1]), S(occs[0]), S(occs[1]), S(out),
             P(refs[0]), P(flows[0]), P(filts[0]), P(refs[1]), P(flows[1]), P(filts[1]), P(occs[0]), P(occs[1]), P(out),
             lib.OVERWRITE)


def net_case(B, H, W, star, iters):
    """One frame pair of MEMC_Net (star=False) / MEMC_Net_star (star=True) per batch item."""
    dev = "cuda"
    flows = [synth.smooth_flow(B, H, W, 6.0, seed=s, device=dev) for s in (1, 2)]
    refs = [synth.image(B, 3, H, W, seed=s, device=dev) for s in (3, 4)]
    filts = [synth.softmax_filter(B, 4, H, W, seed=s, device=dev) for s in (5, 6)]
    occs = [torch.rand(B, 1, H, W, device=dev) for _ in range(2)]
    ctxs = [torch.randn(B, 64, H, W, device=dev) for _ in range(2)] if star else []
    proj = [torch.empty_like(f) for f in flows]
    cnt = [torch.empty(B, 1, H, W, device=dev) for _ in range(2)]
    warp = [torch.empty_like(r) for r in refs]
    cwarp = [torch.empty_like(c) for c in ctxs]

    def ours():
        for k in range(2):
            ours_fp(flows[k], cnt[k], proj[k])
        for k in range(2):
            ours_fi(refs[k], proj[k], filts[k], warp[k])
        out = occs[0] * warp[0] + occs[1] * warp[1]
        for k in range(len(ctxs)):
            ours_fi(ctxs[k], proj[k], filts[k], cwarp[k])
        return out

    blended = torch.empty_like(refs[0])

    def ours_fused():  # memc_b200.fused.FilterInterpolate in place of the two warps + blend
        for k in range(2):
            ours_fp(flows[k], cnt[k], proj[k])
        ours_blend(refs, proj, filts, occs, blended)
        for k in range(len(ctxs)):
            ours_fi(ctxs[k], proj[k], filts[k], cwarp[k])
        return blended

    def legacy():
        for k in range(2):
            cnt[k].zero_(); proj[k].zero_()
            ref.gpu_flow_projection_forward(flows[k], 1, (cnt[k], proj[k]))
        for k in range(2):
            warp[k].zero_()
            ref.gpu_filter_interpolation_forward(refs[k], proj[k], filts[k], warp[k])
        out = occs[0] * warp[0] + occs[1] * warp[1]
        for k in range(len(ctxs)):
            cwarp[k].zero_()
            ref.gpu_filter_interpolation_forward(ctxs[k], proj[k], filts[k], cwarp[k])
        return out

    # parity of the sequence: the projected flows agree to fp32 summation-order noise; the warps are
    # compared on the SAME projected flow (a 1e-6 flow difference can move a target across an
    # integer and legitimately change that pixel's 4x4 window)
    legacy()
    lproj = [t.clone() for t in proj]
    ours()
    torch.cuda.synchronize()
    err_fp = max(float((proj[k] - lproj[k]).abs().max()) for k in range(2))
    err = 0.0
    for k in range(2):
        a, b = torch.empty_like(refs[k]), torch.zeros_like(refs[k])
        ours_fi(refs[k], lproj[k], filts[k], a)
        ref.gpu_filter_interpolation_forward(refs[k], lproj[k], filts[k], b)
        err = max(err, float((a - b).abs().max()))
    for k in range(len(ctxs)):
        a, b = torch.empty_like(ctxs[k]), torch.zeros_like(ctxs[k])
        ours_fi(ctxs[k], lproj[k], filts[k], a)
        ref.gpu_filter_interpolation_forward(ctxs[k], lproj[k], filts[k], b)
        err = max(err, float((a - b).abs().max()))
    del a, b
    t_o, t_l = timeit(ours, iters), timeit(legacy, max(3, iters // 2))
    t_f = timeit(ours_fused, iters)
    return {"name": "%s op sequence, B=%d x %dx%d" % ("MEMC_Net_star" if star else "MEMC_Net", B, W, H),
            "ours_ms": t_o * 1e3, "ours_fused_blend_ms": t_f * 1e3, "legacy_ms": t_l * 1e3, "speedup": t_l / t_o,
            "speedup_fused": t_l / t_f, "max_abs_projected_flow_vs_legacy": err_fp,
            "max_abs_warps_vs_legacy_same_flow": err,
            "frame_pairs_per_s": B / t_o}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--out", default="gpurun_out/legacy_bench.json")
    ap.add_argument("--only", default="ops,net")
    args = ap.parse_args()
    only = set(args.only.split(","))
    lib.load()
    if not ref.available_gpu():
        print("oracle/_ref/libmemc_ref_gpu.so not present")
        return
    rows = []

    def emit(r):
        rows.append(r)
        print(json.dumps(r), flush=True)

    if "ops" in only:
        for (B, C, H, W) in [(4, 3, 1080, 1920), (4, 3, 720, 1280), (1, 64, 1080, 1920)]:
            in1, flow, filt, gout = synth.filter_interpolation_case(B, C, H, W, seed=0, device="cuda")
            out = torch.empty_like(in1)
            g = [torch.empty_like(in1), torch.empty_like(flow), torch.empty_like(filt)]
            st = lib.stream_ptr(in1)

            def bwd():
                lib.call("memc_b200_filter_interpolation_backward", st, B, C, H, W, 4, S(in1), S(flow), S(filt), S(gout),
                         S(g[0]), S(g[1]), S(g[2]), P(in1), P(flow), P(filt), P(gout), P(g[0]), P(g[1]), P(g[2]), lib.OVERWRITE)

            def lfwd():
                out.zero_()
                ref.gpu_filter_interpolation_forward(in1, flow, filt, out)

            def lbwd():
                for t in g:
                    t.zero_()
                ref.gpu_filter_interpolation_backward(in1, flow, filt, gout, g)

            t1, t2 = timeit(lambda: ours_fi(in1, flow, filt, out), args.iters), timeit(lfwd, max(3, args.iters // 2))
            emit({"name": "FI fwd B%d C%d %dx%d" % (B, C, W, H), "ours_ms": t1 * 1e3, "legacy_ms": t2 * 1e3, "speedup": t2 / t1})
            t1, t2 = timeit(bwd, args.iters), timeit(lbwd, max(3, args.iters // 2))
            emit({"name": "FI bwd B%d C%d %dx%d" % (B, C, W, H), "ours_ms": t1 * 1e3, "legacy_ms": t2 * 1e3, "speedup": t2 / t1})
            del in1, flow, filt, gout, out, g
            torch.cuda.empty_cache()
        B, H, W = 16, 1080, 1920
        for kind, flow in (("smooth", synth.smooth_flow(B, H, W, 6.0, seed=1, device="cuda")),
                           ("uniform", synth.uniform_flow(B, H, W, 32.0, seed=2, device="cuda")),
                           ("contention", synth.radial_flow(B, H, W, 0.9, device="cuda")),
                           ("tear", synth.tear_flow(B, H, W, 24.0, seed=3, device="cuda"))):
            count, out = torch.empty(B, 1, H, W, device="cuda"), torch.empty_like(flow)

            def lfp():
                count.zero_(); out.zero_()
                ref.gpu_flow_projection_forward(flow, 1, (count, out))

            t1, t2 = timeit(lambda: ours_fp(flow, count, out), args.iters), timeit(lfp, max(3, args.iters // 2))
            emit({"name": "FP fwd+fill %s B%d %dx%d" % (kind, B, W, H), "ours_ms": t1 * 1e3, "legacy_ms": t2 * 1e3,
                  "speedup": t2 / t1})
        torch.cuda.empty_cache()
    if "net" in only:
        for (B, H, W, star) in [(1, 768, 1344, False), (4, 768, 1344, False), (1, 1152, 1984, True), (2, 1152, 1984, True)]:
            emit(net_case(B, H, W, star, args.iters))
            torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
