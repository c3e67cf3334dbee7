"""FlowProjection pipeline phase timing (development): MEMC_FP_DBG skips phases (results invalid)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "memc-net_b200")):
    sys.path.insert(0, p)
from memc_b200 import lib, synth
from tools.kbench import timeit, S, P
lib.load()
B, H, W = 16, 1080, 1920
kinds = sys.argv[1].split(",") if len(sys.argv) > 1 else ["smooth"]
for kind in kinds:
    flow = {"smooth": lambda: synth.smooth_flow(B, H, W, 6.0, seed=1, device="cuda"),
            "uniform": lambda: synth.uniform_flow(B, H, W, 32.0, seed=2, device="cuda"),
            "contention": lambda: synth.radial_flow(B, H, W, 0.9, device="cuda"),
            "tear": lambda: synth.tear_flow(B, H, W, 24.0, seed=3, device="cuda")}[kind]()
    count, out = torch.empty(B, 1, H, W, device="cuda"), torch.empty_like(flow)
    st = lib.stream_ptr(flow)
    def fwd():
        lib.call("memc_b200_flow_projection_forward", st, B, H, W, 1, S(flow), S(count), S(out), P(flow), P(count), P(out), lib.OVERWRITE)
    for dbg in (0, 7, 6, 5, 3, 1, 2, 4):
        os.environ["MEMC_FP_DBG"] = str(dbg)
        print(kind, "skip-mask", dbg, "ms %.4f" % (timeit(fwd, 10) * 1e3), flush=True)
    os.environ["MEMC_FP_DBG"] = "0"
