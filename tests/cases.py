"""Shared seeded test inputs (numpy, CPU).  Edge cases follow SURVEY.md section 8(d):
exact-integer flows, x2 == W-1, |f| >= W/2, negative targets, NaN flow, fs in {2,4,5,6}."""
import numpy as np


def fi_case(B, C, H, W, fs, sigma, seed, edge=True):
    rng = np.random.default_rng(seed)
    in1 = rng.random((B, C, H, W), dtype=np.float32)
    flow = (rng.standard_normal((B, 2, H, W)) * sigma).astype(np.float32)
    filt = (rng.standard_normal((B, fs * fs, H, W)) * 0.25).astype(np.float32)
    gout = rng.standard_normal((B, C, H, W)).astype(np.float32)
    if edge and H >= 8 and W >= 8:
        flow[:, :, 0, :] = np.round(flow[:, :, 0, :])          # exact integer targets
        flow[:, 0, 1, :] = (W - 1) - np.arange(W)              # x2 == W-1 exactly
        flow[:, 1, 1, :] = 0.0
        flow[:, 0, 2, 0] = W / 2.0                             # |fx| >= W/2  -> invalid
        flow[:, 0, 2, 1] = -3.5                                # negative target -> invalid
        flow[:, 1, 2, 2] = np.nan                              # NaN -> invalid branch
        flow[:, 0, 3, :] = 0.0                                 # zero flow row
        flow[:, 1, 3, :] = 0.0
        flow[:, 1, 4, :] = (H - 1) - 4.0                       # y2 == H-1 exactly
    return in1, flow, filt, gout


def flow_case(B, H, W, sigma, seed, edge=True):
    rng = np.random.default_rng(seed)
    flow = (rng.standard_normal((B, 2, H, W)) * sigma).astype(np.float32)
    if edge and H >= 8 and W >= 8:
        flow[:, :, 0, :] = np.round(flow[:, :, 0, :])
        flow[:, 0, 1, :] = (W - 1) - np.arange(W)
        flow[:, 1, 1, :] = 0.0
        flow[:, 0, 2, 1] = -30.5
        flow[:, 1, 2, 2] = np.nan
    return flow


def sepconv_case(B, C, H, W, fs, seed):
    rng = np.random.default_rng(seed)
    in1 = rng.random((B, C, H, W), dtype=np.float32)
    v = rng.standard_normal((B, fs, H - fs + 1, W - fs + 1)).astype(np.float32)
    hz = rng.standard_normal((B, fs, H - fs + 1, W - fs + 1)).astype(np.float32)
    gout = rng.standard_normal((B, C, H - fs + 1, W - fs + 1)).astype(np.float32)
    return in1, v, hz, gout
