"""Seeded synthetic inputs of the shapes BASELINE.json names.

All generators take a `torch.Generator` seed and a device and are deterministic per
(seed, shape).  They are used by bench.py and by the tests so that both run the same data.
"""
import torch
import torch.nn.functional as F


def _gen(seed, device):
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return g


def smooth_flow(B, H, W, sigma, seed=0, device="cpu", jitter=0.25, grid=16):
    """Low-res N(0, sigma^2) field at (H/grid, W/grid) bilinearly upsampled, plus per-pixel
    N(0, jitter^2): smooth motion with fractional positions everywhere.  grid=16 is the
    SURVEY's benchmark field (flow gradient ~0.5 px/px: rather rough); larger grids are smoother."""
    g = _gen(seed, device)
    lh, lw = max(2, H // grid), max(2, W // grid)
    low = torch.randn(B, 2, lh, lw, generator=g, device=device) * sigma
    flow = F.interpolate(low, size=(H, W), mode="bilinear", align_corners=True)
    flow = flow + torch.randn(B, 2, H, W, generator=g, device=device) * jitter
    return flow.contiguous()


def inverse_depth(B, H, W, seed=0, device="cpu", grid=64, spread=0.7, objects=True):
    """A weight map for DepthFlowProjection shaped like what a depth network delivers: 1e-6 + 1 / depth with
    depth = 4 * exp(spread * smooth N(0,1) field on a (H/grid, W/grid) lattice) -- slowly varying over a 64 x 16 source tile --
    and, with `objects`, a few rectangles 3x nearer than their surroundings (depth edges: weight jumps inside a tile)."""
    g = _gen(seed, device)
    lh, lw = max(2, H // grid), max(2, W // grid)
    low = torch.randn(B, 1, lh, lw, generator=g, device=device)
    depth = 4.0 * torch.exp(spread * F.interpolate(low, size=(H, W), mode="bilinear", align_corners=True))
    if objects and H >= 64 and W >= 64:
        for k in range(6):
            y0 = int(torch.randint(0, H - H // 4, (1,), generator=g, device=device))
            x0 = int(torch.randint(0, W - W // 4, (1,), generator=g, device=device))
            depth[:, :, y0:y0 + H // 6 + 8 * k, x0:x0 + W // 8 + 8 * k] *= 1.0 / 3.0
    return (1e-6 + 1.0 / depth).contiguous()


def uniform_flow(B, H, W, amplitude, seed=0, device="cpu"):
    g = _gen(seed, device)
    return ((torch.rand(B, 2, H, W, generator=g, device=device) * 2 - 1) * amplitude).contiguous()


def radial_flow(B, H, W, gain, device="cpu"):
    """flow = gain * (centre - p): gain ~0.9 makes nearly every pixel splat into a few
    hundred cells (atomic contention); a negative gain diverges and tears large holes."""
    ys = torch.arange(H, device=device, dtype=torch.float32).view(1, 1, H, 1)
    xs = torch.arange(W, device=device, dtype=torch.float32).view(1, 1, 1, W)
    fx = gain * ((W - 1) / 2.0 - xs).expand(B, 1, H, W)
    fy = gain * ((H - 1) / 2.0 - ys).expand(B, 1, H, W)
    return torch.cat([fx, fy], dim=1).contiguous()


def tear_flow(B, H, W, amplitude, seed=0, device="cpu", jitter=0.25):
    """Two motion layers pulling apart: the left half moves left and the right half moves right
    by `amplitude`, the top half up and the bottom half down -> a cross of holes 2*amplitude wide
    in the projected flow (the fill-hole stress case: long searches, whole columns empty)."""
    g = _gen(seed, device)
    xs = torch.arange(W, device=device, dtype=torch.float32).view(1, 1, 1, W)
    ys = torch.arange(H, device=device, dtype=torch.float32).view(1, 1, H, 1)
    fx = torch.where(xs < W / 2.0, -float(amplitude), float(amplitude)).expand(B, 1, H, W)
    fy = torch.where(ys < H / 2.0, -float(amplitude), float(amplitude)).expand(B, 1, H, W)
    flow = torch.cat([fx, fy], dim=1) + torch.randn(B, 2, H, W, generator=g, device=device) * jitter
    return flow.contiguous()


def image(B, C, H, W, seed=0, device="cpu"):
    return torch.rand(B, C, H, W, generator=_gen(seed, device), device=device)


def softmax_filter(B, fs, H, W, seed=0, device="cpu"):
    """Per-pixel kernels that are positive and sum to 1 over the fs*fs taps."""
    return torch.softmax(torch.randn(B, fs * fs, H, W, generator=_gen(seed, device), device=device), dim=1)


def gaussian_filter(B, fs, H, W, seed=0, device="cpu", scale=0.25):
    return torch.randn(B, fs * fs, H, W, generator=_gen(seed, device), device=device) * scale


def grad_like(t, seed=0):
    return torch.randn(t.shape, generator=_gen(seed, t.device), device=t.device)


def filter_interpolation_case(B, C, H, W, fs=4, sigma=None, seed=0, device="cpu", grid=16):
    """(input1, flow, filter, gradoutput) for FilterInterpolation; sigma defaults to the
    SURVEY's 4 px @720p / 6 px @1080p scaled by height."""
    if sigma is None:
        sigma = 6.0 * H / 1080.0 if H >= 900 else 4.0 * max(H, 64) / 720.0
    in1 = image(B, C, H, W, seed, device)
    flow = smooth_flow(B, H, W, sigma, seed + 1, device, grid=grid)
    filt = softmax_filter(B, fs, H, W, seed + 2, device)
    gout = grad_like(in1, seed + 3)
    return in1, flow, filt, gout
