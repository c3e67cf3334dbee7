"""Run the REFERENCE'S OWN networks (networks/MEMC_Net*.py, unmodified) on top of a chosen
implementation of the hot-path ops  (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

The networks look the ops up by module-global name at call time
(`FilterInterpolationModule()(ref0, offset[0], filter[0])`, networks/MEMC_Net.py:252-264,
MEMC_Net_s.py:255-264, MEMC_Net_star.py:272-285, MEMC_Net_VE.py:494-504), so ONE constructed
network can be run twice on identical weights and inputs:

  impl "ours"    the my_package of this repo (libmemc_b200.so) -- what `import networks` bound
  impl "ref"     my_package-shaped Modules bound to the reference's own kernels: the legacy CUDA
                 kernels recompiled for sm_100a (oracle/_ref/libmemc_ref_gpu.so) for CUDA tensors,
                 the reference's my_lib.c (oracle/_ref/libmemc_ref_cpu.so) for CPU tensors
                 (fill-hole has no CPU twin in the reference, my_lib.c:1539-1543: our restatement)
  impl "oracle"  my_package-shaped Modules bound to oracle/memc_oracle.c (CPU tensors only)

Where the networks come from: /root/reference when it exists (this container); on the GPU box the
copy staged by `make -C oracle ref_py` into oracle/_ref/reference_py/ (git-ignored, travels with
the gpurun snapshot exactly like the compiled oracle/_ref/*.so).
"""
import contextlib
import math
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_STAGED = os.path.join(_HERE, "_ref", "reference_py")
_PKG = os.path.join(os.path.dirname(_HERE), "memc-net_b200")


def reference_root():
    for root in ("/root/reference", _STAGED):
        if os.path.isfile(os.path.join(root, "networks", "MEMC_Net.py")):
            return root
    return None


def available():
    return reference_root() is not None


def install():
    """Make `import networks` work: our my_package first, then the reference's networks/ + Stack.py."""
    root = reference_root()
    if root is None:
        raise FileNotFoundError("reference networks not found (run `make -C oracle ref_py` where /root/reference exists)")
    if _PKG not in sys.path:
        sys.path.insert(0, _PKG)
    from memc_b200 import compat
    compat.install(None)
    # only networks/ and Stack.py may come from the reference tree: its my_package must stay shadowed
    if root not in sys.path:
        sys.path.append(root)
    import my_package
    assert os.path.abspath(my_package.__file__).startswith(_PKG), my_package.__file__
    return root


# ---------------------------------------------------------------------------------------------
# my_package-shaped Modules on the reference kernels / the oracle
# ---------------------------------------------------------------------------------------------
def _np(t):
    return np.ascontiguousarray(t.detach().cpu().numpy(), dtype=np.float32)


def _make_modules(impl):
    import torch
    from torch import nn
    from oracle import cpu, ref

    def like(a, t):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(t.device)

    class FilterInterpolationModule(nn.Module):
        def forward(self, input1, input2, input3):
            if input1.is_cuda:
                assert impl == "ref"
                return ref.gpu_filter_interpolation_forward(input1.contiguous(), input2.contiguous(), input3.contiguous())
            fn = ref.cpu_filter_interpolation_forward if impl == "ref" else cpu.filter_interpolation_forward
            return like(fn(_np(input1), _np(input2), _np(input3)), input1)

    class FlowProjectionModule(nn.Module):
        def __init__(self, requires_grad=True):
            super().__init__()
            self.fillhole = 0 if requires_grad else 1  # functions/FlowProjectionLayer.py:15

        def forward(self, input1):
            if input1.is_cuda:
                assert impl == "ref"
                out, _count = ref.gpu_flow_projection_forward(input1.contiguous(), self.fillhole)
                return out
            out, _count = cpu.flow_projection_forward(_np(input1), self.fillhole)
            return like(out, input1)

    class InterpolationModule(nn.Module):
        def forward(self, input1, input2):
            if input1.is_cuda:
                assert impl == "ref"
                return ref.gpu_interpolation_forward(input1.contiguous(), input2.contiguous())
            fn = ref.cpu_interpolation_forward if impl == "ref" else cpu.interpolation_forward
            return like(fn(_np(input1), _np(input2)), input1)

    return {"FilterInterpolationModule": FilterInterpolationModule, "FlowProjectionModule": FlowProjectionModule,
            "InterpolationModule": InterpolationModule}


@contextlib.contextmanager
def ops(net, impl):
    """Run `net` (an instance of a reference network class) with the hot-path ops of `impl`."""
    mod = sys.modules[type(net).__module__]
    if impl == "ours":
        yield
        return
    new = _make_modules(impl)
    old = {k: getattr(mod, k) for k in new if hasattr(mod, k)}
    try:
        for k in old:
            setattr(mod, k, new[k])
        yield
    finally:
        for k, v in old.items():
            setattr(mod, k, v)


# ---------------------------------------------------------------------------------------------
# building and running
# ---------------------------------------------------------------------------------------------
def build_network(name, seed=0, device="cpu", motion=0.0):
    """Random-init reference network in inference mode (torch.manual_seed(seed), training=False, .eval()).

    `motion` > 0 adds a smooth random displacement field of that many pixels (sigma, in final flow units) to
    the flow estimator's output through a forward hook: a random-init estimator predicts ~0 flow and no
    pretrained weights exist offline.  The hook belongs to the network object, so every arm sees it."""
    import torch
    install()
    import networks
    import warnings
    torch.manual_seed(seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        net = getattr(networks, name)(training=False)
    net.eval()
    if motion:
        # a forward hook on the flow estimator (both arms see it): + a smooth seeded displacement field
        scale = 0.5 if hasattr(net.flownets, "moduleBasic") else net.div_flow / 2.0  # what the caller multiplies by

        def add_motion(_m, _inp, out):
            g = torch.Generator().manual_seed(seed + 77)
            B, _, h, w = out.shape
            low = torch.randn(B, 2, max(2, h // 8), max(2, w // 8), generator=g) * (motion / scale)
            field = torch.nn.functional.interpolate(low, size=(h, w), mode="bilinear", align_corners=True)
            return out + field.to(out.device, out.dtype)

        net.flownets.register_forward_hook(add_motion)
    return net.to(device)


def synthetic_frames(B, H, W, seed=0, device="cpu"):
    """[2, B, 3, H, W] in [0, 1]: smooth random texture and a shifted/perturbed second frame."""
    import torch
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(seed)
    low = torch.rand(B, 3, max(2, H // 8) + 2, max(2, W // 8) + 2, generator=g)
    f0 = F.interpolate(low, size=(H + 16, W + 16), mode="bicubic", align_corners=False).clamp(0, 1)
    fine = torch.rand(B, 3, H + 16, W + 16, generator=g) * 0.15
    f0 = (0.85 * f0 + fine).clamp(0, 1)
    a, b = f0[:, :, 8:8 + H, 8:8 + W], f0[:, :, 5:5 + H, 12:12 + W]  # second frame = shifted view
    return torch.stack([a, b], 0).contiguous().to(device)


def run(net, frames, impl):
    """Forward in inference mode -> dict of the tensors the network returns."""
    import torch
    import warnings
    with torch.no_grad(), warnings.catch_warnings(), ops(net, impl):
        warnings.simplefilter("ignore")
        outs, offsets, filters, occlusions = net(frames)
    return {"output": outs[0], "rectified": outs[1], "offset0": offsets[0], "offset1": offsets[1],
            "filter0": filters[0], "filter1": filters[1]}


def compare(a, b):
    """max-abs, max-abs relative to the dynamic range of b, and PSNR (peak = range of b) of a vs b."""
    import torch
    a, b = a.double(), b.double()
    finite = bool(torch.isfinite(a).all() and torch.isfinite(b).all())
    rng = float((b.max() - b.min()).abs()) or 1.0
    err = float((a - b).abs().max())
    mse = float(((a - b) ** 2).mean())
    psnr = float("inf") if mse == 0 else 10.0 * math.log10(rng * rng / mse)
    return {"finite": finite, "max_abs": err, "range": rng, "max_rel": err / rng, "psnr_db": psnr}


# ---------------------------------------------------------------------------------------------
# MEMC_Net_VE (video enhancement) on the reference's in-tree Vimeo fixtures
# ---------------------------------------------------------------------------------------------
def vimeo_septuplet(index=0, kind="input", device="cpu"):
    """The 7 frames of one septuplet of vimeo_video_enhancement_test/ as a list of [1,3,H,W] tensors in
    [0,1], replication-padded like demo_Vimeo_VE.py:113-133 (to multiples of 128; 32 px where already one)."""
    import torch
    from PIL import Image
    root = reference_root()
    base = os.path.join(root or "", "vimeo_video_enhancement_test")
    lst = os.path.join(base, "sep_testlist.txt")
    if not os.path.isfile(lst):
        return None
    seqs = [l.strip() for l in open(lst) if l.strip()]
    d = os.path.join(base, kind, seqs[index % len(seqs)])
    frames = []
    for k in range(1, 8):
        im = np.asarray(Image.open(os.path.join(d, "im%d.png" % k)).convert("RGB"), dtype=np.float32) / 255.0
        frames.append(torch.from_numpy(im.transpose(2, 0, 1).copy())[None])
    h, w = frames[0].shape[2:]

    def pad(n):
        if n != ((n >> 7) << 7):
            tot = (((n >> 7) + 1) << 7) - n
            return tot // 2, tot - tot // 2
        return 32, 32

    (pl, pr), (pt, pb) = pad(w), pad(h)
    pader = torch.nn.ReplicationPad2d([pl, pr, pt, pb])
    return [pader(f).contiguous().to(device) for f in frames]


def build_network_ve(seed=0, device="cpu", motion=0.0):
    """Random-init MEMC_Net_VE (its constructor asks model_zoo for ResNet-18 weights: no network here, so the
    download is replaced by "no pretrained entries" and conv1 keeps its random init)."""
    install()
    import networks.ResNet.Resnet_conv1 as rc
    old = rc.model_zoo.load_url
    rc.model_zoo.load_url = lambda *a, **k: {}
    try:
        return build_network("MEMC_Net_VE", seed=seed, device=device, motion=motion)
    finally:
        rc.model_zoo.load_url = old


def run_ve(net, frames, impl):
    import torch
    import warnings
    with torch.no_grad(), warnings.catch_warnings(), ops(net, impl):
        warnings.simplefilter("ignore")
        return net(frames)
