"""Generates tests/golden/*.npz from the REFERENCE'S OWN CPU code.

Run where /root/reference exists:   python tests/golden/make_golden.py
It builds oracle/_ref/libmemc_ref_cpu.so (my_package/src/my_lib.c compiled unchanged against
oracle/th_stub/TH.h, see oracle/Makefile), feeds it the seeded inputs of tests/cases.py and
stores inputs + reference outputs.  Fill-hole has no CPU implementation in the reference
(my_lib.c:1539-1543), so flow-projection fixtures are scatter+average only; the fill-hole
fixture is produced on the GPU box by tests/golden/make_golden_gpu.py from the reference
CUDA kernels.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
from tests.cases import fi_case, flow_case, sepconv_case  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    ref.build()
    assert ref.available_cpu(), "reference CPU library could not be built"
    for name, (B, C, H, W, fs, sigma, seed) in {
            "fi_rgb_fs4": (2, 3, 16, 20, 4, 3.0, 100),
            "fi_c5_fs5": (1, 5, 12, 14, 5, 2.0, 101),
            "fi_fs2": (1, 2, 9, 11, 2, 1.5, 102),
            "fi_fs6_wild": (1, 3, 14, 14, 6, 20.0, 103)}.items():
        in1, flow, filt, gout = fi_case(B, C, H, W, fs, sigma, seed)
        out = ref.cpu_filter_interpolation_forward(in1, flow, filt)
        g1, g2, g3 = ref.cpu_filter_interpolation_backward(in1, flow, filt, gout)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), op="filter_interpolation", in1=in1, flow=flow,
                            filt=filt, gout=gout, out=out, g1=g1, g2=g2, g3=g3)
    for name, (B, H, W, sigma, seed) in {"fp_small": (2, 12, 16, 3.0, 110), "fp_wild": (1, 10, 10, 15.0, 111)}.items():
        flow = flow_case(B, H, W, sigma, seed)
        out, count = ref.cpu_flow_projection_forward(flow)
        gout = np.random.default_rng(seed).standard_normal(flow.shape).astype(np.float32)
        gi = ref.cpu_flow_projection_backward(flow, count, gout)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), op="flow_projection", flow=flow, out=out,
                            count=count, gout=gout, gi=gi)
    for name, (B, H, W, sigma, seed) in {"dfp_small": (2, 12, 16, 3.0, 112), "dfp_wild": (1, 10, 10, 15.0, 113)}.items():
        flow = flow_case(B, H, W, sigma, seed)
        rng = np.random.default_rng(seed)
        depth = (1e-6 + 1.0 / rng.uniform(0.5, 20.0, (B, 1, H, W))).astype(np.float32)
        out, count = ref.cpu_depth_flow_projection_forward(flow, depth)
        gout = rng.standard_normal(flow.shape).astype(np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            g1, g2 = ref.cpu_depth_flow_projection_backward(flow, depth, count, out, gout)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), op="depth_flow_projection", flow=flow, depth=depth, out=out,
                            count=count, gout=gout, g1=g1, g2=g2)
    for name, (B, H, W, sigma, seed, thr) in {"wfp_small": (2, 12, 16, 3.0, 114, 0.3), "wfp_all": (1, 10, 10, 2.0, 115, 5.0)}.items():
        flow = flow_case(B, H, W, sigma, seed)
        rng = np.random.default_rng(seed)
        im0, im1 = rng.random((B, 3, H, W), dtype=np.float32), rng.random((B, 3, H, W), dtype=np.float32)
        out, count, weight = ref.cpu_weighted_flow_projection_forward(flow, im0, im1, thr)
        gout = rng.standard_normal(flow.shape).astype(np.float32)
        gi = ref.cpu_weighted_flow_projection_backward(flow, im0, im1, count, weight, gout, thr)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), op="weighted_flow_projection", flow=flow, im0=im0, im1=im1,
                            threshold=np.float32(thr), out=out, count=count, weight=weight, gout=gout, gi=gi)
    for name, (mode, B, H, W, sigma, seed) in {"px_value": ("value", 2, 12, 16, 3.0, 116), "px_weight": ("weight", 1, 10, 14, 6.0, 117),
                                               "px_reliable": ("reliable", 1, 9, 9, 2.0, 118)}.items():
        flow = flow_case(B, H, W, sigma, seed)
        rng = np.random.default_rng(seed)
        in1 = rng.random((B, 3, H, W), dtype=np.float32) if mode == "value" else np.zeros(0, np.float32)
        fw = rng.random((B, 1, H, W), dtype=np.float32) if mode != "reliable" else np.zeros(0, np.float32)
        sd = 1.3
        out = ref.cpu_pixel_splat_forward(mode, flow, in1, fw, sd)
        gout = rng.standard_normal(out.shape).astype(np.float32)
        thr = float(np.quantile(out, 0.35)) if mode != "value" else 0.0
        g1, g3, gw = ref.cpu_pixel_splat_backward(mode, flow, gout, in1, fw, out, sd, thr)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), op="pixel_splat", mode=mode, flow=flow, in1=in1, fw=fw,
                            sigma_d=np.float32(sd), threshold=np.float32(thr), out=out, gout=gout,
                            g1=g1 if g1 is not None else np.zeros(0, np.float32), g3=g3,
                            gw=gw if gw is not None else np.zeros(0, np.float32))
    for name, (B, H, W, sigma, seed) in {"wl_small": (2, 12, 16, 3.0, 119), "wl_wild": (1, 10, 10, 15.0, 120)}.items():
        flow = flow_case(B, H, W, sigma, seed)
        rng = np.random.default_rng(seed)
        in1, in2 = rng.random((B, 3, H, W), dtype=np.float32), rng.random((B, 3, H, W), dtype=np.float32)
        lam = 0.9
        out = ref.cpu_weight_layer_forward(in1, in2, flow, lam)
        gout = rng.standard_normal(out.shape).astype(np.float32)
        g1, g2, g3 = ref.cpu_weight_layer_backward(in1, in2, flow, out, gout, lam)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), op="weight_layer", in1=in1, in2=in2, flow=flow, lambda_e=np.float32(lam),
                            out=out, gout=gout, g1=g1, g2=g2, g3=g3)
    for name, (B, C, H, W, fs, seed) in {"scf_fs4": (2, 3, 12, 15, 4, 121)}.items():
        in1, v, hz, _ = sepconv_case(B, C, H, W, fs, seed)
        v, hz = np.abs(v) + np.float32(0.01), np.abs(hz) + np.float32(0.01)   # positive sums: the C and CUDA sources agree
        v[0, :, 0, 0] = 0.0
        flow = ref.cpu_separable_conv_flow_forward(in1, v, hz)
        gflow = np.random.default_rng(seed).standard_normal(flow.shape).astype(np.float32)
        gv, gh = ref.cpu_separable_conv_flow_backward(in1, v, hz, gflow)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), op="separable_conv_flow", in1=in1, vert=v, horiz=hz, flow=flow,
                            gflow=gflow, gv=gv, gh=gh)
    for name, (B, C, H, W, sigma, seed) in {"ip_rgb": (2, 3, 12, 16, 3.0, 120), "ip_c7": (1, 7, 9, 13, 2.0, 121)}.items():
        in1, flow, _, gout = fi_case(B, C, H, W, 4, sigma, seed)
        out = ref.cpu_interpolation_forward(in1, flow)
        g1, g2 = ref.cpu_interpolation_backward(in1, flow, gout)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), op="interpolation", in1=in1, flow=flow, gout=gout,
                            out=out, g1=g1, g2=g2)
    for name, (B, C, H, W, fs, seed) in {"sc_fs4": (2, 3, 12, 15, 4, 130), "sc_fs3": (1, 3, 8, 8, 3, 131)}.items():
        in1, v, hz, gout = sepconv_case(B, C, H, W, fs, seed)
        out = ref.cpu_separable_conv_forward(in1, v, hz)
        g1, g2, g3 = ref.cpu_separable_conv_backward(in1, v, hz, gout)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), op="separable_conv", in1=in1, vert=v, horiz=hz,
                            gout=gout, out=out, g1=g1, g2=g2, g3=g3)
    print("wrote", sorted(f for f in os.listdir(OUT) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
