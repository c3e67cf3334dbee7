import sys; sys.path.insert(0, "memc-net_b200"); sys.path.insert(0, ".")
import torch
from memc_b200 import lib, synth
from tools.kbench import timeit, S, P
lib.load()
B, H, W = 16, 1080, 1920
for kind, fl in (("smooth", synth.smooth_flow(B, H, W, 6.0, seed=1, device="cuda")), ("uniform", synth.uniform_flow(B, H, W, 32.0, seed=2, device="cuda")),
                 ("convergent", synth.radial_flow(B, H, W, 0.9, device="cuda")), ("divergent", synth.radial_flow(B, H, W, -0.5, device="cuda"))):
    cnt, prj = torch.empty(B, 1, H, W, device="cuda"), torch.empty_like(fl)
    st = lib.stream_ptr(fl)
    for v in (0, 2, 0, 2):  # MEMC_B200_VARIANT(2): the pipeline without its L2 eviction hints
        f = lambda: lib.call("memc_b200_flow_projection_forward", st, B, H, W, 1, S(fl), S(cnt), S(prj), P(fl), P(cnt), P(prj),
                             lib.OVERWRITE | lib.variant(v))
        print(kind, "variant", v, "%.3f ms" % (timeit(f, 15, flush=False) * 1e3), flush=True)
