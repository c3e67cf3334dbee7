"""DepthFlowProjection forward, B frames of 1920x1080: the shared-memory splat (MEMC_B200_VARIANT 0) vs the generic scatter
(1) inside the frame driver (development tool).   python tools/dfp_time.py [B]"""
import sys; sys.path.insert(0, "memc-net_b200"); sys.path.insert(0, ".")
import torch
from memc_b200 import lib, synth
from tools.kbench import timeit, S, P
lib.load()
B, H, W = int(sys.argv[1]) if len(sys.argv) > 1 else 16, 1080, 1920
dw = synth.inverse_depth(B, H, W, seed=7, device="cuda")
for kind, fl in (("smooth", synth.smooth_flow(B, H, W, 6.0, seed=1, device="cuda")), ("convergent", synth.radial_flow(B, H, W, 0.9, device="cuda")),
                 ("uniform", synth.uniform_flow(B, H, W, 32.0, seed=2, device="cuda"))):
    cnt, prj = torch.empty(B, 1, H, W, device="cuda"), torch.empty_like(fl)
    st = lib.stream_ptr(fl)
    for v in (0, 1, 0, 1):
        f = lambda: lib.call("memc_b200_depth_flow_projection_forward", st, B, H, W, 1, S(fl), S(dw), S(cnt), S(prj),
                             P(fl), P(dw), P(cnt), P(prj), lib.OVERWRITE | lib.variant(v))
        print(kind, "variant", v, "%.3f ms" % (timeit(f, 10, flush=False) * 1e3), flush=True)
