"""Sweep the FilterInterpolation TMA tile configurations (development tool).
    python tools/sweep_fi.py [--iters 20] [--out gpurun_out/sweep_fi.json]
Kernel variants are selected through the MEMC_B200_VARIANT field of the flags (include/memc_b200.h)."""
import argparse, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "memc-net_b200")):
    sys.path.insert(0, p)
from memc_b200 import lib
from tools.kbench import timeit, fi_calls, _peak

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--out", default="gpurun_out/sweep_fi.json")
ap.add_argument("--fwd", default="0,1,2")
ap.add_argument("--bwd", default="0,1")
args = ap.parse_args()
lib.load()
peak, _ = _peak()
B, C, H, W = 4, 3, 1080, 1920
px = B * H * W
rows = []
(in1, flow, filt, gout), fwd, bwd = fi_calls(B, C, H, W, lib.OVERWRITE)
# reference results from the generic kernels for a correctness check of every config
from tools.kbench import S, P
def run_fwd(flags):
    o = torch.empty_like(in1)
    lib.call("memc_b200_filter_interpolation_forward", lib.stream_ptr(in1), B, C, H, W, 4, S(in1), S(flow), S(filt), S(o),
             P(in1), P(flow), P(filt), P(o), flags)
    return o
def run_bwd(flags):
    g1, g2, g3 = torch.empty_like(in1), torch.empty_like(flow), torch.empty_like(filt)
    lib.call("memc_b200_filter_interpolation_backward", lib.stream_ptr(in1), B, C, H, W, 4, S(in1), S(flow), S(filt), S(gout),
             S(g1), S(g2), S(g3), P(in1), P(flow), P(filt), P(gout), P(g1), P(g2), P(g3), flags)
    return g1, g2, g3
ref_o = run_fwd(lib.OVERWRITE | lib.NO_FAST)
ref_g = run_bwd(lib.OVERWRITE | lib.NO_FAST)
torch.cuda.synchronize()
for cfg in [int(c) for c in args.fwd.split(",") if c != ""]:
    try:
        fl = lib.OVERWRITE | lib.variant(cfg)
        o = run_fwd(fl); torch.cuda.synchronize()
        ok = bool(torch.equal(o, ref_o))
        err = float((o - ref_o).abs().max())
        t = timeit(lambda: run_fwd(fl), args.iters)
        r = {"op": "fwd", "cfg": cfg, "ms": t * 1e3, "frac": px * 96 / t / 1e9 / peak, "bitwise_equal_generic": ok,
             "max_abs_vs_generic": err}
    except Exception as e:
        r = {"op": "fwd", "cfg": cfg, "error": str(e)[:200]}
    rows.append(r); print(json.dumps(r), flush=True)
for cfg in [int(c) for c in args.bwd.split(",") if c != ""]:
    try:
        fl = lib.OVERWRITE | lib.variant(cfg)
        g = run_bwd(fl); torch.cuda.synchronize()
        err = [float((a - b).abs().max()) for a, b in zip(g, ref_g)]
        t = timeit(lambda: run_bwd(fl), args.iters)
        r = {"op": "bwd", "cfg": cfg, "ms": t * 1e3, "frac": px * 180 / t / 1e9 / peak, "max_abs_vs_generic": err}
    except Exception as e:
        r = {"op": "bwd", "cfg": cfg, "error": str(e)[:200]}
    rows.append(r); print(json.dumps(r), flush=True)
os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
json.dump(rows, open(args.out, "w"), indent=1)
