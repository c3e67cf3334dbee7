"""Run each fast-path kernel once on small inputs (for compute-sanitizer memcheck / racecheck / synccheck)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "memc-net_b200")):
    sys.path.insert(0, p)
from memc_b200 import lib, synth
from tools.kbench import S, P
lib.load()
for (B, C, H, W, sigma) in [(1, 3, 96, 160, 3.0), (2, 3, 72, 132, 20.0), (1, 4, 64, 96, 1.0)]:
    in1, flow, filt, gout = synth.filter_interpolation_case(B, C, H, W, sigma=sigma, seed=1, device="cuda")
    out = torch.empty_like(in1)
    g1, g2, g3 = torch.empty_like(in1), torch.empty_like(flow), torch.empty_like(filt)
    st = lib.stream_ptr(in1)
    lib.call("memc_b200_filter_interpolation_forward", st, B, C, H, W, 4, S(in1), S(flow), S(filt), S(out),
             P(in1), P(flow), P(filt), P(out), lib.OVERWRITE)
    lib.call("memc_b200_filter_interpolation_backward", st, B, C, H, W, 4, S(in1), S(flow), S(filt), S(gout),
             S(g1), S(g2), S(g3), P(in1), P(flow), P(filt), P(gout), P(g1), P(g2), P(g3), lib.OVERWRITE)
    for fl in (flow, synth.radial_flow(B, H, W, 0.9, device="cuda"), synth.tear_flow(B, H, W, 10.0, device="cuda")):
        count, o2 = torch.empty(B, 1, H, W, device="cuda"), torch.empty_like(fl)
        lib.call("memc_b200_flow_projection_forward", lib.stream_ptr(fl), B, H, W, 1, S(fl), S(count), S(o2),
                 P(fl), P(count), P(o2), lib.OVERWRITE)
    torch.cuda.synchronize()
print("sanitize_fast: done")
