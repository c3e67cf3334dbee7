"""Launch one op a few times (for ncu captures).   python tools/run_one.py fi_fwd|fi_bwd|fp_fwd|... [B] [flags]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "memc-net_b200")):
    sys.path.insert(0, p)
from memc_b200 import lib, synth  # noqa: E402
from tools.kbench import fi_calls, S, P  # noqa: E402

op = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
flags = int(sys.argv[3]) if len(sys.argv) > 3 else lib.OVERWRITE
C = int(os.environ.get("C", "3"))
H, W = int(os.environ.get("H", "1080")), int(os.environ.get("W", "1920"))
lib.load()
if op in ("fi_fwd", "fi_bwd"):
    _, fwd, bwd = fi_calls(B, C, H, W, flags)
    for _ in range(3):
        (fwd if op == "fi_fwd" else bwd)()
elif op == "fp_fwd":
    kind = os.environ.get("FLOW", "smooth")
    flow = {"smooth": lambda: synth.smooth_flow(B, H, W, 6.0, seed=1, device="cuda"),
            "uniform": lambda: synth.uniform_flow(B, H, W, 32.0, seed=2, device="cuda"),
            "contention": lambda: synth.radial_flow(B, H, W, 0.9, device="cuda")}[kind]()
    count, out = torch.empty(B, 1, H, W, device="cuda"), torch.empty_like(flow)
    for _ in range(3):
        lib.call("memc_b200_flow_projection_forward", lib.stream_ptr(flow), B, H, W, 1, S(flow), S(count), S(out),
                 P(flow), P(count), P(out), flags)
torch.cuda.synchronize()
if op in ("ip_fwd", "fp_bwd"):
    H, W = 1080, 1920
    in1, flow, _, gout = synth.filter_interpolation_case(4, 3, H, W, seed=0, device="cuda")
    st = lib.stream_ptr(in1)
    if op == "ip_fwd":
        out = torch.empty_like(in1)
        for _ in range(3):
            lib.call("memc_b200_interpolation_forward", st, 4, 3, H, W, S(in1), S(flow), S(out), P(in1), P(flow), P(out), flags)
    else:
        fl = synth.smooth_flow(16, H, W, 6.0, seed=1, device="cuda")
        cnt, prj = torch.empty(16, 1, H, W, device="cuda"), torch.empty_like(fl)
        lib.call("memc_b200_flow_projection_forward", st, 16, H, W, 0, S(fl), S(cnt), S(prj), P(fl), P(cnt), P(prj), flags)
        go, gi = torch.randn_like(fl), torch.empty_like(fl)
        for _ in range(3):
            lib.call("memc_b200_flow_projection_backward", st, 16, H, W, S(fl), S(cnt), S(go), S(gi), P(fl), P(cnt), P(go),
                     P(gi), flags)
    torch.cuda.synchronize()
if op in ("ip_bwd", "sc_bwd", "sc_fwd"):
    H, W = 1080, 1920
    B = 4
    in1, flow, _, gout = synth.filter_interpolation_case(B, 3, H, W, seed=0, device="cuda")
    st = lib.stream_ptr(in1)
    if op == "ip_bwd":
        g1, g2 = torch.empty_like(in1), torch.empty_like(flow)
        for _ in range(3):
            lib.call("memc_b200_interpolation_backward", st, B, 3, H, W, S(in1), S(flow), S(gout), S(g1), S(g2), P(in1),
                     P(flow), P(gout), P(g1), P(g2), flags)
    else:
        fs = 4
        Ho, Wo = H - fs + 1, W - fs + 1
        v, hz = torch.randn(B, fs, Ho, Wo, device="cuda"), torch.randn(B, fs, Ho, Wo, device="cuda")
        out = torch.empty(B, 3, Ho, Wo, device="cuda")
        go = torch.randn_like(out)
        g1, g2, g3 = torch.empty_like(in1), torch.empty_like(v), torch.empty_like(hz)
        for _ in range(3):
            if op == "sc_fwd":
                lib.call("memc_b200_separable_conv_forward", st, B, 3, H, W, fs, S(in1), S(v), S(hz), S(out), P(in1), P(v),
                         P(hz), P(out), flags)
            else:
                lib.call("memc_b200_separable_conv_backward", st, B, 3, H, W, fs, S(in1), S(v), S(hz), S(go), S(g1), S(g2),
                         S(g3), P(in1), P(v), P(hz), P(go), P(g1), P(g2), P(g3), flags)
    torch.cuda.synchronize()
if op in ("dfp_fwd", "wfp_fwd"):
    H, W = 1080, 1920
    fl = synth.smooth_flow(B, H, W, 6.0, seed=1, device="cuda")
    cnt, prj = torch.empty(B, 1, H, W, device="cuda"), torch.empty_like(fl)
    st = lib.stream_ptr(fl)
    if op == "dfp_fwd":
        dw = synth.inverse_depth(B, H, W, seed=7, device="cuda")
        for _ in range(3):
            lib.call("memc_b200_depth_flow_projection_forward", st, B, H, W, 1, S(fl), S(dw), S(cnt), S(prj), P(fl), P(dw), P(cnt),
                     P(prj), flags)
    else:
        g = torch.Generator(device="cuda").manual_seed(9)
        f0 = torch.nn.functional.interpolate(torch.rand(B, 3, H // 30, W // 30, device="cuda", generator=g), size=(H, W),
                                             mode="bilinear", align_corners=True).contiguous()
        f2 = (f0 + 0.2 * torch.randn(B, 3, H, W, device="cuda", generator=g)).contiguous()
        wgt = torch.empty(B, 1, H, W, device="cuda")
        for _ in range(3):
            lib.call("memc_b200_weighted_flow_projection_forward", st, B, H, W, 1, 0.16, S(fl), S(f0), S(f2), S(cnt), S(wgt), S(prj),
                     P(fl), P(f0), P(f2), P(cnt), P(wgt), P(prj), flags)
    torch.cuda.synchronize()
