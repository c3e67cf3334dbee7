"""FlowProjectPair (one call on the concatenated bidirectional pair) vs two FlowProjection calls (development)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "memc-net_b200")):
    sys.path.insert(0, p)
from memc_b200 import fused, lib, synth
from my_package.modules.FlowProjectionModule import FlowProjectionModule
from tools.kbench import timeit
lib.load()
with torch.no_grad():
    for (B, H, W) in [(1, 768, 1344), (4, 768, 1344), (1, 1152, 1984), (8, 1152, 1984)]:
        fa, fb = synth.smooth_flow(B, H, W, 6.0, seed=1, device="cuda"), synth.smooth_flow(B, H, W, 6.0, seed=2, device="cuda")
        t2 = timeit(lambda: (FlowProjectionModule(False)(fa), FlowProjectionModule(False)(fb)), 10)
        t1 = timeit(lambda: fused.FlowProjectPair(fa, fb), 10)
        print("B=%d %dx%d  two calls %.4f ms  pair %.4f ms" % (B, W, H, t2 * 1e3, t1 * 1e3), flush=True)
