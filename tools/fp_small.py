"""FlowProjection forward at small batch: persistent pipeline vs per-frame launches (development)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "memc-net_b200")):
    sys.path.insert(0, p)
from memc_b200 import lib, synth
from tools.kbench import timeit, S, P
lib.load()
for (B, H, W) in [(1, 768, 1344), (2, 768, 1344), (4, 768, 1344), (1, 1152, 1984), (2, 1152, 1984), (8, 1152, 1984), (16, 1080, 1920)]:
    flow = synth.smooth_flow(B, H, W, 6.0, seed=1, device="cuda")
    count, out = torch.empty(B, 1, H, W, device="cuda"), torch.empty_like(flow)
    st = lib.stream_ptr(flow)
    def fwd():
        lib.call("memc_b200_flow_projection_forward", st, B, H, W, 1, S(flow), S(count), S(out), P(flow), P(count), P(out), lib.OVERWRITE)
    r = []
    for dbg in ("0", "128"):
        os.environ["MEMC_TMA_DBG"] = dbg
        r.append(timeit(fwd, 10) * 1e3)
    os.environ["MEMC_TMA_DBG"] = "0"
    print("B=%d %dx%d  pipeline %.4f ms  per-frame launches %.4f ms" % (B, W, H, r[0], r[1]), flush=True)
