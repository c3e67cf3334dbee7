"""Kernel micro-benchmarks on one GPU (development tool; bench.py is the contract bench).

    python tools/kbench.py [--out gpurun_out/kbench.json] [--iters 20] [--only fi,fp,...]

Times every op through the C ABI with CUDA events (median of --iters after 3 warm-ups, inputs
far larger than L2 or an L2 flush between iterations).  The comparison with the reference's
legacy kernels recompiled for sm_100a lives in tests/legacy_bench.py (only tests/ may touch
oracle/).
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


def _peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
    _flush.zero_()


def timeit(fn, iters, flush=True):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            flush_l2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    ts.sort()
    return ts[len(ts) // 2]


def S(t):
    return lib.strides_of(t)


def P(t):
    return lib.ptr(t)


def fi_calls(B, C, H, W, flags, grid=16):
    in1, flow, filt, gout = synth.filter_interpolation_case(B, C, H, W, seed=0, device="cuda", grid=grid)
    out = torch.empty_like(in1)
    g1, g2, g3 = torch.empty_like(in1), torch.empty_like(flow), torch.empty_like(filt)
    st = lib.stream_ptr(in1)

    def fwd():
        lib.call("memc_b200_filter_interpolation_forward", st, B, C, H, W, 4, S(in1), S(flow), S(filt), S(out),
                 P(in1), P(flow), P(filt), P(out), flags)

    def bwd():
        lib.call("memc_b200_filter_interpolation_backward", st, B, C, H, W, 4, S(in1), S(flow), S(filt), S(gout),
                 S(g1), S(g2), S(g3), P(in1), P(flow), P(filt), P(gout), P(g1), P(g2), P(g3), flags)

    return (in1, flow, filt, gout), fwd, bwd


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/kbench.json")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    only = set(args.only.split(",")) if args.only else None
    peak, peak_kind = _peak()
    lib.load()
    have_ref, ref = False, None  # legacy comparison: tests/legacy_bench.py
    res = {"peak_gbs": peak, "peak_kind": peak_kind, "gpu": torch.cuda.get_device_name(0), "rows": []}

    def row(name, px, bytes_per_px, t, t_ref=None, **kw):
        r = {"name": name, "ms": t * 1e3, "mpx_s": px / t / 1e6, "gbs": px * bytes_per_px / t / 1e9,
             "frac": px * bytes_per_px / t / 1e9 / peak}
        if t_ref:
            r.update(ref_ms=t_ref * 1e3, speedup_vs_legacy=t_ref / t)
        r.update(kw)
        res["rows"].append(r)
        print(json.dumps(r), flush=True)

    if only is None or "fi" in only:
        for (B, C, H, W) in [(4, 3, 1080, 1920), (4, 3, 720, 1280), (1, 3, 1080, 1920), (1, 64, 1080, 1920)]:
            px = B * H * W
            for flags, tag in ((lib.OVERWRITE, "fast"), (lib.OVERWRITE | lib.NO_FAST, "generic")):
                (in1, flow, filt, gout), fwd, bwd = fi_calls(B, C, H, W, flags)
                tf, tb = timeit(fwd, args.iters), timeit(bwd, args.iters)
                trf = trb = None
                if have_ref and tag == "fast":
                    o = torch.zeros_like(in1)
                    gs = (torch.zeros_like(in1), torch.zeros_like(flow), torch.zeros_like(filt))
                    trf = timeit(lambda: ref.gpu_filter_interpolation_forward(in1, flow, filt, o), max(5, args.iters // 4))

                    def rb():
                        for g in gs:
                            g.zero_()
                        ref.gpu_filter_interpolation_backward(in1, flow, filt, gout, gs)
                    trb = timeit(rb, max(5, args.iters // 4))
                row("FI fwd %s B%d C%d %dx%d" % (tag, B, C, W, H), px, (2 * C + 18) * 4, tf, trf)
                row("FI bwd %s B%d C%d %dx%d" % (tag, B, C, W, H), px, (3 * C + 36) * 4, tb, trb)
                del in1, flow, filt, gout
                torch.cuda.empty_cache()

    if only is None or "fismooth" in only:
        # the same op on smoother motion fields (flow gradient ~0.5 / 0.12 / 0.03 px per px)
        B, C, H, W = 4, 3, 1080, 1920
        for grid in (16, 64, 256):
            (in1, flow, filt, gout), fwd, bwd = fi_calls(B, C, H, W, lib.OVERWRITE, grid=grid)
            row("FI fwd fast B4 C3 1080p flow-grid %d" % grid, B * H * W, 96, timeit(fwd, args.iters))
            row("FI bwd fast B4 C3 1080p flow-grid %d" % grid, B * H * W, 180, timeit(bwd, args.iters))
            del in1, flow, filt, gout
            torch.cuda.empty_cache()

    if only is None or "fp" in only:
        B, H, W = 16, 1080, 1920
        px = B * H * W
        flows = {"smooth": synth.smooth_flow(B, H, W, 6.0, seed=1, device="cuda"),
                 "uniform": synth.uniform_flow(B, H, W, 32.0, seed=2, device="cuda"),
                 "contention": synth.radial_flow(B, H, W, 0.9, device="cuda"),
                 "tear": synth.tear_flow(B, H, W, 24.0, seed=3, device="cuda")}
        for kind, flow in flows.items():
            count = torch.empty(B, 1, H, W, device="cuda")
            out = torch.empty_like(flow)
            st = lib.stream_ptr(flow)
            for flags, tag in ((lib.OVERWRITE, "fast"), (lib.OVERWRITE | lib.NO_FAST, "generic")):
                def fwd():
                    lib.call("memc_b200_flow_projection_forward", st, B, H, W, 1, S(flow), S(count), S(out),
                             P(flow), P(count), P(out), flags)
                t = timeit(fwd, args.iters)
                tr = None
                if have_ref and tag == "fast":
                    def rf():
                        count.zero_()
                        out.zero_()
                        ref.gpu_flow_projection_forward(flow, 1, (count, out))
                    tr = timeit(rf, max(3, args.iters // 4))
                row("FP fwd+fill %s %s B%d %dx%d" % (tag, kind, B, W, H), px, 20, t, tr)
            gout, gi = torch.randn_like(flow), torch.empty_like(flow)
            fwd()

            def bwd():
                lib.call("memc_b200_flow_projection_backward", st, B, H, W, S(flow), S(count), S(gout), S(gi),
                         P(flow), P(count), P(gout), P(gi), lib.OVERWRITE)
            row("FP bwd %s B%d %dx%d" % (kind, B, W, H), px, 28, timeit(bwd, args.iters))
            del count, out, gout, gi
        del flows
        torch.cuda.empty_cache()

    if only is None or "ip" in only:
        B, C, H, W = 4, 3, 1080, 1920
        px = B * H * W
        in1, flow, _, gout = synth.filter_interpolation_case(B, C, H, W, seed=0, device="cuda")
        out, g1, g2 = torch.empty_like(in1), torch.empty_like(in1), torch.empty_like(flow)
        st = lib.stream_ptr(in1)
        row("IP fwd B4 C3 1080p", px, (2 * C + 2) * 4, timeit(lambda: lib.call(
            "memc_b200_interpolation_forward", st, B, C, H, W, S(in1), S(flow), S(out), P(in1), P(flow), P(out),
            lib.OVERWRITE), args.iters))
        row("IP bwd B4 C3 1080p", px, (3 * C + 4) * 4, timeit(lambda: lib.call(
            "memc_b200_interpolation_backward", st, B, C, H, W, S(in1), S(flow), S(gout), S(g1), S(g2), P(in1),
            P(flow), P(gout), P(g1), P(g2), lib.OVERWRITE), args.iters))

    if only is None or "sc" in only:
        B, C, H, W, fs = 4, 3, 1080, 1920, 4
        Ho, Wo = H - fs + 1, W - fs + 1
        px = B * Ho * Wo
        in1 = synth.image(B, C, H, W, device="cuda")
        v, hz = torch.randn(B, fs, Ho, Wo, device="cuda"), torch.randn(B, fs, Ho, Wo, device="cuda")
        out = torch.empty(B, C, Ho, Wo, device="cuda")
        gout = torch.randn_like(out)
        g1, g2, g3 = torch.empty_like(in1), torch.empty_like(v), torch.empty_like(hz)
        st = lib.stream_ptr(in1)
        row("SC fwd B4 1080p fs4", px, (2 * C + 2 * fs) * 4, timeit(lambda: lib.call(
            "memc_b200_separable_conv_forward", st, B, C, H, W, fs, S(in1), S(v), S(hz), S(out), P(in1), P(v), P(hz),
            P(out), lib.OVERWRITE), args.iters))
        row("SC bwd B4 1080p fs4", px, (3 * C + 4 * fs) * 4, timeit(lambda: lib.call(
            "memc_b200_separable_conv_backward", st, B, C, H, W, fs, S(in1), S(v), S(hz), S(gout), S(g1), S(g2),
            S(g3), P(in1), P(v), P(hz), P(gout), P(g1), P(g2), P(g3), lib.OVERWRITE), args.iters))

    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
