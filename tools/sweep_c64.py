"""FilterInterpolation forward, C = 64: variants vs the generic kernel (development tool)."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "memc-net_b200")):
    sys.path.insert(0, p)
from memc_b200 import lib
from tools.kbench import timeit, fi_calls, _peak, S, P
lib.load()
peak, _ = _peak()
B, C, H, W = 1, int(os.environ.get("C", "64")), 1080, 1920
px = B * H * W
(in1, flow, filt, gout), fwd, bwd = fi_calls(B, C, H, W, lib.OVERWRITE)
def run(flags):
    o = torch.empty_like(in1)
    lib.call("memc_b200_filter_interpolation_forward", lib.stream_ptr(in1), B, C, H, W, 4, S(in1), S(flow), S(filt), S(o),
             P(in1), P(flow), P(filt), P(o), flags)
    torch.cuda.synchronize()
    return o
ref_o = run(lib.OVERWRITE | lib.NO_FAST)
for cfg in [int(c) for c in (sys.argv[1] if len(sys.argv) > 1 else "0,1,2").split(",")]:
    try:
        fl = lib.OVERWRITE | lib.variant(cfg)  # MEMC_B200_VARIANT: 0 production (tap-column lanes), 1 generic kernel, 2 round-1 patches
        ok = bool(torch.equal(run(fl), ref_o))
        t = timeit(lambda: run(fl), 10)
        print(json.dumps({"cfg": cfg, "ms": t * 1e3, "frac": px * (2 * C + 18) * 4 / t / 1e9 / peak, "bitwise_equal_generic": ok}), flush=True)
    except Exception as e:
        print(json.dumps({"cfg": cfg, "error": str(e)[:200]}), flush=True)
# backward, C = 64: production (channel-chunked tap-row lanes) vs the generic kernel (MEMC_B200_VARIANT(1))
for name, fl in (("bwd chunked", lib.OVERWRITE), ("bwd generic", lib.OVERWRITE | lib.variant(1))):
    _, _, b = fi_calls(B, C, H, W, fl)
    t = timeit(b, 10)
    print(json.dumps({"cfg": name, "ms": t * 1e3, "frac": px * (3 * C + 36) * 4 / t / 1e9 / peak}), flush=True)
