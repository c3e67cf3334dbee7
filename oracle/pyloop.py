"""Literal pure-Python loops (numpy float32 scalars) for tiny cases -- BASELINE.json
configs[0] "FilterInterpolation forward 64x64 RGB, 4x4 kernel -- pure-PyTorch CPU loop
oracle".  TEST INFRASTRUCTURE ONLY.  Each loop follows the same reference lines as the C
restatement (my_lib.c:963-1073; my_lib.c:1491-1536 + my_lib_kernel.cu:1776-1833) and uses
np.float32 arithmetic in the reference's evaluation order, so it is bit-identical to it.
"""
import numpy as np

f32 = np.float32


def filter_interpolation_forward(in1, flow, filt):
    B, C, H, W = in1.shape
    fs = int(np.sqrt(f32(filt.shape[1])))
    out = np.zeros_like(in1)
    one = f32(1)
    for b in range(B):
        for h in range(H):
            for w in range(W):
                fx, fy = flow[b, 0, h, w], flow[b, 1, h, w]
                x2, y2 = f32(w) + fx, f32(h) + fy
                valid = (x2 >= 0 and y2 >= 0 and x2 <= f32(W - 1) and y2 <= f32(H - 1)
                         and abs(fx) < f32(W) / f32(2) and abs(fy) < f32(H) / f32(2))
                if not valid:
                    out[b, :, h, w] = in1[b, :, h, w]
                    continue
                ix, iy = int(x2), int(y2)
                L, T = ix + 1 - fs // 2, iy + 1 - fs // 2
                a, bt = x2 - f32(ix), y2 - f32(iy)
                for c in range(C):
                    q = [f32(0)] * 4
                    for qi, (rows, cols) in enumerate([
                            (range(T, iy + 1), range(L, ix + 1)), (range(T, iy + 1), range(ix + 1, L + fs)),
                            (range(iy + 1, T + fs), range(L, ix + 1)),
                            (range(iy + 1, T + fs), range(ix + 1, L + fs))]):
                        acc = f32(0)
                        for j in rows:
                            jj = min(max(0, j), H - 1)
                            for i in cols:
                                ii = min(max(0, i), W - 1)
                                acc = f32(acc + f32(in1[b, c, jj, ii] * filt[b, (j - T) * fs + (i - L), h, w]))
                        q[qi] = acc
                    v = f32(f32(f32(one - a) * f32(one - bt)) * q[0])
                    v = f32(v + f32(f32(a * f32(one - bt)) * q[1]))
                    v = f32(v + f32(f32(f32(one - a) * bt) * q[2]))
                    v = f32(v + f32(f32(a * bt) * q[3]))
                    out[b, c, h, w] = v
    return out


def flow_projection_forward(flow, fillhole):
    B, _, H, W = flow.shape
    out = np.zeros_like(flow)
    count = np.zeros((B, 1, H, W), np.float32)
    for b in range(B):
        for h in range(H):
            for w in range(W):
                fx, fy = flow[b, 0, h, w], flow[b, 1, h, w]
                x2, y2 = f32(w) + fx, f32(h) + fy
                if not (x2 >= 0 and y2 >= 0 and x2 <= f32(W - 1) and y2 <= f32(H - 1)):
                    continue
                L, T = int(x2), int(y2)
                R, Bm = min(L + 1, W - 1), min(T + 1, H - 1)
                for (y, x) in [(T, L), (T, R), (Bm, L), (Bm, R)]:
                    out[b, 0, y, x] = f32(out[b, 0, y, x] + (-fx))
                    out[b, 1, y, x] = f32(out[b, 1, y, x] + (-fy))
                    count[b, 0, y, x] += f32(1)
        hit = count[b, 0] > 0
        out[b, 0][hit] = out[b, 0][hit] / count[b, 0][hit]
        out[b, 1][hit] = out[b, 1][hit] / count[b, 0][hit]
        if not fillhole:
            continue
        src = out[b].copy()
        for h in range(H):
            for w in range(W):
                if count[b, 0, h, w] > 0:
                    continue
                found = []
                for step in [(0, -1), (0, 1), (-1, 0)]:     # left, right, up; never down
                    y, x = h + step[0], w + step[1]
                    while 0 <= y < H and 0 <= x < W and count[b, 0, y, x] == 0:
                        y, x = y + step[0], x + step[1]
                    if 0 <= y < H and 0 <= x < W:
                        found.append((y, x))
                if not found:
                    continue
                for ch in range(2):
                    s = f32(0)
                    for (y, x) in found:
                        s = f32(s + src[ch, y, x])
                    out[b, ch, h, w] = f32(s / f32(len(found)))
    return out, count
