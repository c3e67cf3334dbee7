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
# persistent FlowProjection pipeline (3+ frames), fused pair + blend, channel-chunked C > 4 forward, SeparableConv fs = 4
B, H, W = 5, 70, 132
for fl in (synth.smooth_flow(B, H, W, 4.0, seed=2, device="cuda"), synth.radial_flow(B, H, W, 0.9, device="cuda"),
           synth.tear_flow(B, H, W, 10.0, device="cuda")):
    count, o2 = torch.empty(B, 1, H, W, device="cuda"), torch.empty_like(fl)
    lib.call("memc_b200_flow_projection_forward", lib.stream_ptr(fl), B, H, W, 1, S(fl), S(count), S(o2),
             P(fl), P(count), P(o2), lib.OVERWRITE)
cases = [synth.filter_interpolation_case(2, 3, 96, 160, sigma=3.0, seed=k, device="cuda") for k in (1, 2)]
occ = [torch.rand(2, 1, 96, 160, device="cuda") for _ in range(2)]
out = torch.empty_like(cases[0][0])
(r0, f0, w0, _), (r1, f1, w1, _) = cases
lib.call("memc_b200_filter_interpolation_blend_forward", lib.stream_ptr(out), 2, 3, 96, 160, 4, S(r0), S(f0), S(w0), S(r1), S(f1),
         S(w1), S(occ[0]), S(occ[1]), S(out), P(r0), P(f0), P(w0), P(r1), P(f1), P(w1), P(occ[0]), P(occ[1]), P(out), lib.OVERWRITE)
in1, flow, filt, _ = synth.filter_interpolation_case(1, 10, 64, 136, sigma=3.0, seed=3, device="cuda")
out = torch.empty_like(in1)
lib.call("memc_b200_filter_interpolation_forward", lib.stream_ptr(in1), 1, 10, 64, 136, 4, S(in1), S(flow), S(filt), S(out),
         P(in1), P(flow), P(filt), P(out), lib.OVERWRITE)
in1 = synth.image(1, 3, 40, 70, device="cuda")
v, hz = torch.randn(1, 4, 37, 67, device="cuda"), torch.randn(1, 4, 37, 67, device="cuda")
out = torch.empty(1, 3, 37, 67, device="cuda")
go = torch.randn_like(out)
g1, g2, g3 = torch.empty_like(in1), torch.empty_like(v), torch.empty_like(hz)
st = lib.stream_ptr(in1)
lib.call("memc_b200_separable_conv_forward", st, 1, 3, 40, 70, 4, S(in1), S(v), S(hz), S(out), P(in1), P(v), P(hz), P(out), lib.OVERWRITE)
lib.call("memc_b200_separable_conv_backward", st, 1, 3, 40, 70, 4, S(in1), S(v), S(hz), S(go), S(g1), S(g2), S(g3), P(in1), P(v),
         P(hz), P(go), P(g1), P(g2), P(g3), lib.OVERWRITE)
# round 2: channel-chunked backward (C = 8, borders + wild flow), shared flow / filter pair, DepthFlowProjection (shared-memory
# splat, signed masks), WeightedFlowProjection, fused mean
for (B, C, H, W, sigma) in [(1, 8, 64, 96, 2.0), (1, 12, 40, 132, 25.0)]:
    in1, flow, filt, gout = synth.filter_interpolation_case(B, C, H, W, sigma=sigma, seed=5, device="cuda")
    g1, g2, g3 = torch.empty_like(in1), torch.empty_like(flow), torch.empty_like(filt)
    lib.call("memc_b200_filter_interpolation_backward", lib.stream_ptr(in1), B, C, H, W, 4, S(in1), S(flow), S(filt), S(gout),
             S(g1), S(g2), S(g3), P(in1), P(flow), P(filt), P(gout), P(g1), P(g2), P(g3), lib.OVERWRITE)
r_in, r_flow, r_filt, _ = synth.filter_interpolation_case(1, 3, 64, 96, sigma=3.0, seed=6, device="cuda")
x_in = torch.randn(1, 8, 64, 96, device="cuda")
r_out, x_out = torch.empty_like(r_in), torch.empty_like(x_in)
lib.call("memc_b200_filter_interpolation_forward_pair", lib.stream_ptr(r_in), 1, 3, 8, 64, 96, 4, S(r_in), S(x_in), S(r_flow), S(r_filt),
         S(r_out), S(x_out), P(r_in), P(x_in), P(r_flow), P(r_filt), P(r_out), P(x_out), lib.OVERWRITE)
B, H, W = 2, 70, 132
dw = synth.inverse_depth(B, H, W, seed=7, device="cuda")
f0, f2 = torch.rand(B, 3, H, W, device="cuda"), torch.rand(B, 3, H, W, device="cuda")
for fl in (synth.smooth_flow(B, H, W, 4.0, seed=2, device="cuda"), synth.radial_flow(B, H, W, 0.9, device="cuda"),
           synth.tear_flow(B, H, W, 10.0, device="cuda")):
    cnt, wgt, prj = torch.empty(B, 1, H, W, device="cuda"), torch.empty(B, 1, H, W, device="cuda"), torch.empty_like(fl)
    st = lib.stream_ptr(fl)
    lib.call("memc_b200_depth_flow_projection_forward", st, B, H, W, 1, S(fl), S(dw), S(cnt), S(prj), P(fl), P(dw), P(cnt), P(prj),
             lib.OVERWRITE)
    go, g1, g2 = torch.randn_like(fl), torch.empty_like(fl), torch.empty_like(dw)
    lib.call("memc_b200_depth_flow_projection_backward", st, B, H, W, S(fl), S(dw), S(cnt), S(prj), S(go), S(g1), S(g2),
             P(fl), P(dw), P(cnt), P(prj), P(go), P(g1), P(g2), lib.OVERWRITE)
    lib.call("memc_b200_weighted_flow_projection_forward", st, B, H, W, 1, 0.3, S(fl), S(f0), S(f2), S(cnt), S(wgt), S(prj),
             P(fl), P(f0), P(f2), P(cnt), P(wgt), P(prj), lib.OVERWRITE)
    lib.call("memc_b200_weighted_flow_projection_backward", st, B, H, W, 0.3, S(fl), S(f0), S(f2), S(cnt), S(go), S(g1),
             P(fl), P(f0), P(f2), P(cnt), P(go), P(g1), lib.OVERWRITE)
torch.cuda.synchronize()
print("sanitize_fast: done")
