#!/usr/bin/env python
"""Offline model of the shared-memory bank conflicts of the adaptive-warp gather / scatter.

For one 1080p frame of the benchmark motion field (memc_b200.synth.smooth_flow) every warp-wide
access of the staged image box is replayed: the number of shared-memory wavefronts of one warp
instruction is  max over the 32 banks of the number of DISTINCT words addressed in that bank
(loads broadcast equal addresses; for atomics equal addresses serialise too, reported as `atom`).
Box origin and channel plane only shift all addresses of an instruction by a constant, so the
address of tap (i, j) of a pixel with integer source (ix, iy) is  (iy + j) * pitch + ix + i.

Mappings (lane -> pixel) are parameterised as  lanes_w x lanes_h  lanes, each thread owning `ppt`
pixels spaced `xstep` apart in x (xstep == 1: adjacent), optionally staggered by lane row so that
at one instruction the lane rows work on different sub-pixels.

  python tools/bank_model.py            # table for the candidate mappings / pitches
  python tools/bank_model.py --roles    # round 2: lanes = (pixel, tap row / tap column) maps
No GPU needed.  Results are summarised in profiles/r01_bank_conflict_model.md.
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "memc-net_b200"))
from memc_b200 import synth  # noqa: E402


def wavefronts(addr, active):
    """addr [N,32] int64 word addresses, active [N,32] bool -> (load wavefronts [N], atomic wavefronts [N])."""
    n = addr.shape[0]
    bank = addr & 31
    key = np.where(active, bank * (1 << 40) + (addr - addr.min() + 1), -1)
    order = np.argsort(key, axis=1)
    ks = np.take_along_axis(key, order, axis=1)
    bs = np.take_along_axis(bank, order, axis=1)
    act = ks >= 0
    uniq = act.copy()
    uniq[:, 1:] &= ks[:, 1:] != ks[:, :-1]
    rows = np.repeat(np.arange(n)[:, None], 32, axis=1)
    cnt_u = np.zeros((n, 32), np.int32)
    cnt_a = np.zeros((n, 32), np.int32)
    np.add.at(cnt_u, (rows[uniq], bs[uniq]), 1)
    np.add.at(cnt_a, (rows[act], bs[act]), 1)
    return cnt_u.max(axis=1), cnt_a.max(axis=1)


def model(ix, iy, valid, lw, lh, ppt, xstep, stagger, pitch, sample_taps=(0, 5, 10, 15)):
    """Average wavefronts per warp instruction over the frame."""
    H, W = ix.shape
    # warp region: lw*ppt (if xstep == lw... see below) -- two layouts:
    #   xstep == 1 : thread owns ppt pixels at x = lx + s*lw   (s-th instruction covers an lw-wide strip)
    #   xstep == ppt: thread owns ppt adjacent pixels x = ppt*lx + s
    rw = lw * ppt
    rh = lh
    Hc, Wc = (H // rh) * rh, (W // rw) * rw
    ys = np.arange(0, Hc, rh)
    xs = np.arange(0, Wc, rw)
    lane = np.arange(32)
    lx, ly = lane % lw, lane // lw
    tot_l = tot_a = 0.0
    cnt = 0
    for s in range(ppt):
        sub = (s + (ly if stagger else 0)) % ppt
        px = (ppt * lx + sub) if xstep != 1 else (lx + sub * lw)
        X = (xs[None, :, None] + px[None, None, :]).reshape(1, len(xs), 32)
        Y = (ys[:, None, None] + ly[None, None, :])
        X = np.broadcast_to(X, (len(ys), len(xs), 32)).reshape(-1, 32)
        Y = np.broadcast_to(Y, (len(ys), len(xs), 32)).reshape(-1, 32)
        sx, sy, v = ix[Y, X], iy[Y, X], valid[Y, X]
        for t in sample_taps:
            j, i = t // 4, t % 4
            a = (sy + j).astype(np.int64) * pitch + sx + i
            wl, wa = wavefronts(a, v)
            tot_l += wl.sum()
            tot_a += wa.sum()
            cnt += len(wl)
    return tot_l / cnt, tot_a / cnt


def role_model(ix, iy, valid, pix, colsets, rowsets, pitch):
    """Lanes = (pixel of a small group, tap subset).  pix: list of (dx, dy) of the group's pixels; the 32 / len(pix)
    lanes of a pixel are indexed by role r; role r handles the taps (j, i) for j in rowsets[r], i in colsets[r], one
    warp instruction per (j-index, i-index) pair.  Returns (load, atomic) wavefronts per 32 pixels and channel
    (ideal: 16)."""
    H, W = ix.shape
    pix = np.array(pix)
    npx = len(pix)
    lane = np.arange(32)
    p, r = lane % npx, lane // npx
    gw, gh = int(pix[:, 0].max()) + 1, int(pix[:, 1].max()) + 1
    ys = np.arange(0, H - gh, gh)[::3]
    xs = np.arange(0, W - gw, gw)
    X = np.broadcast_to(xs[None, :, None] + pix[p, 0][None, None, :], (len(ys), len(xs), 32)).reshape(-1, 32)
    Y = np.broadcast_to(ys[:, None, None] + pix[p, 1][None, None, :], (len(ys), len(xs), 32)).reshape(-1, 32)
    sx, sy, v = ix[Y, X], iy[Y, X], valid[Y, X]
    tl = ta = 0.0
    n = steps = 0
    for a in range(len(rowsets[0])):
        for b in range(len(colsets[0])):
            jj = np.array([rowsets[rr][a] for rr in r])
            ii = np.array([colsets[rr][b] for rr in r])
            addr = (sy + jj[None, :]).astype(np.int64) * pitch + sx + ii[None, :]
            wl, wa = wavefronts(addr, v)
            tl += wl.sum(); ta += wa.sum(); n += len(wl); steps += 1
    per32 = (32.0 / npx) * steps
    return tl / n * per32, ta / n * per32


def roles_table(ix, iy, valid):
    row8 = [(i, 0) for i in range(8)]
    g42 = [(i % 4, i // 4) for i in range(8)]
    g82 = [(i % 8, i // 8) for i in range(16)]
    px84 = [(i % 8, i // 8) for i in range(32)]
    px321 = [(i, 0) for i in range(32)]
    all4 = [[0, 1, 2, 3]]
    cases = [
        ("one pixel per lane, 32x1 row segments (round-1 backward)", px321, all4, all4),
        ("one pixel per lane, 8x4 patches (forward C <= 4)", px84, all4, all4),
        ("(pixel, tap ROW) lanes, 8x1 groups (backward C <= 4)", row8, [[0, 1, 2, 3]] * 4, [[0], [1], [2], [3]]),
        ("(pixel, tap ROW) lanes, 4x2 groups", g42, [[0, 1, 2, 3]] * 4, [[0], [1], [2], [3]]),
        ("(pixel, tap COLUMN) lanes, 4 per pixel, 8x1 groups", row8, [[0], [1], [2], [3]], [[0, 1, 2, 3]] * 4),
        ("(pixel, tap COLUMN) lanes, 4 per pixel, 4x2 groups", g42, [[0], [1], [2], [3]], [[0, 1, 2, 3]] * 4),
        ("(pixel, tap COLUMNS r, r+2) lanes, 2 per pixel, 8x2 groups (forward C > 4)", g82, [[0, 2], [1, 3]], [[0, 1, 2, 3]] * 2),
    ]
    pitches = [64, 72, 80]
    print("wavefronts per 32 pixels and channel (ideal 16), load / atomic")
    print("| lane map | " + " | ".join("pitch %d" % p for p in pitches) + " |")
    print("|---|" + "---|" * len(pitches))
    for name, pix, cols, rows in cases:
        cells = []
        for pitch in pitches:
            l, a = role_model(ix, iy, valid, pix, cols, rows, pitch)
            cells.append("%.1f / %.1f" % (l, a))
        print("| %s | " % name + " | ".join(cells) + " |", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--roles", action="store_true", help="lanes = (pixel, tap subset) maps of round 2")
    ap.add_argument("--grid", type=int, default=16)
    ap.add_argument("--sigma", type=float, default=6.0)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--width", type=int, default=1920)
    args = ap.parse_args()
    H, W = args.height, args.width
    flow = synth.smooth_flow(1, H, W, args.sigma, seed=1, grid=args.grid)[0].numpy()
    xx, yy = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32))
    x2, y2 = xx + flow[0], yy + flow[1]
    valid = (x2 >= 0) & (y2 >= 0) & (x2 <= W - 1) & (y2 <= H - 1)
    ix = np.where(valid, x2, 0).astype(np.int32)
    iy = np.where(valid, y2, 0).astype(np.int32)
    if args.roles:
        roles_table(ix, iy, valid)
        return

    maps = [
        ("32x1 (production)", 32, 1, 1, 1, False),
        ("16x2", 16, 2, 1, 1, False),
        ("8x4", 8, 4, 1, 1, False),
        ("4x8", 4, 8, 1, 1, False),
        ("16x2 lanes, 2 adjacent px/thread", 16, 2, 2, 2, False),
        ("16x2 lanes, 2 adjacent px/thread, staggered", 16, 2, 2, 2, True),
        ("8x4 lanes, 4 adjacent px/thread", 8, 4, 4, 4, False),
        ("8x4 lanes, 4 adjacent px/thread, staggered", 8, 4, 4, 4, True),
    ]
    pitches = [64, 68, 72, 76, 80, 88]
    print(f"flow: sigma {args.sigma} grid {args.grid}, {W}x{H}; wavefronts per warp access, load / atomic")
    print("| mapping | " + " | ".join(f"pitch {p}" for p in pitches) + " |")
    print("|---|" + "---|" * len(pitches))
    for name, lw, lh, ppt, xstep, stag in maps:
        cells = []
        for p in pitches:
            l, a = model(ix, iy, valid, lw, lh, ppt, xstep, stag, p)
            cells.append(f"{l:.2f} / {a:.2f}")
        print(f"| {name} | " + " | ".join(cells) + " |", flush=True)


if __name__ == "__main__":
    main()
