"""Debug helper: run FI forward fast vs generic on one shape in THIS process, print verdict."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "memc-net_b200")):
    sys.path.insert(0, p)
from memc_b200 import lib, synth
B, C, H, W = (int(v) for v in sys.argv[1:5])
sigma = float(sys.argv[5]) if len(sys.argv) > 5 else 3.0
lib.load()
t1, t2, t3, _ = synth.filter_interpolation_case(B, C, H, W, sigma=sigma, seed=9, device="cuda")
outs = []
for flags in (lib.OVERWRITE | lib.NO_FAST, lib.OVERWRITE):
    o = torch.full_like(t1, -7.0)
    lib.call("memc_b200_filter_interpolation_forward", lib.stream_ptr(t1), B, C, H, W, 4, lib.strides_of(t1),
             lib.strides_of(t2), lib.strides_of(t3), lib.strides_of(o), lib.ptr(t1), lib.ptr(t2), lib.ptr(t3), lib.ptr(o), flags)
    torch.cuda.synchronize()
    outs.append(o)
d = (outs[0] - outs[1]).abs()
print("shape", (B, C, H, W), "sigma", sigma, "max diff", float(d.max()), "equal", bool(torch.equal(outs[0], outs[1])), flush=True)
