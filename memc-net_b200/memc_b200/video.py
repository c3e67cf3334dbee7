"""Either side of the hot path in the reference's HD demo: frame IO,
the padding the networks need, and a frame-pair driver that shards a sequence over the GPUs of a box.
Host-side Python only; the model is whatever callable the caller brings (the reference's
`networks.MEMC_Net*` with this package's `my_package` under it, or any stand-in).

    network_padding(h, w)            (left, right, top, bottom)       demo_HD720p.py:90-108, demo_MiddleBury.py:98-116
    YUV420Reader / YUV420Writer      planar I420, 8 bit                yuv_frame_io.py:32-103, :106-187
    rgb2yuv / yuv2rgb                the colour matrices the reference takes from scikit-image
    frame_pairs, shard_pairs         which (i, i + step) pairs a rank interpolates
    interpolate_pairs                pad -> model -> crop -> uint8, under torch.no_grad()

THIRD-PARTY ARITHMETIC.  The reference converts colours with `skimage.color.rgb2yuv / yuv2rgb`
(scikit-image 0.13.0, environment.yml:209) and upsamples chroma with `scipy.misc.imresize(...,
interp='nearest')` (SciPy 1.1.0, environment.yml:24).  Neither is installed here, so both are restated
from their published definitions: the YUV matrix of skimage/color/colorconv.py (`yuv_from_rgb`, its
inverse for the way back; `np.dot(arr, M.T)`), and PIL-style nearest resampling, which for the exact
2x chroma upsampling is `out[i, j] = in[i // 2, j // 2]`.  PARITY UNPINNED for these two (no run of the
originals is possible in this image); everything else here follows the reference's own lines.
"""
import os

import numpy as np
import torch

# skimage/color/colorconv.py (0.13.0): yuv_from_rgb; rgb_from_yuv = inv(yuv_from_rgb)
YUV_FROM_RGB = np.array([[0.299, 0.587, 0.114],
                         [-0.14714119, -0.28886916, 0.43601035],
                         [0.61497538, -0.51496512, -0.10001026]], dtype=np.float64)
RGB_FROM_YUV = np.linalg.inv(YUV_FROM_RGB)


def rgb2yuv(rgb):
    """float RGB in [0, 1], HxWx3 -> YUV (Y in [0, 1], U / V centred on 0)."""
    return np.dot(np.asarray(rgb, dtype=np.float64), YUV_FROM_RGB.T.copy())


def yuv2rgb(yuv):
    return np.dot(np.asarray(yuv, dtype=np.float64), RGB_FROM_YUV.T.copy())


def network_padding(height, width):
    """Padding the demos apply before the network (demo_HD720p.py:90-108): each extent goes up to the
    next multiple of 128, split evenly (the odd pixel to the right / bottom); an extent that already IS a
    multiple of 128 still gets 32 + 32.  Returns (left, right, top, bottom) for ReplicationPad2d."""
    def one(n):
        if n != ((n >> 7) << 7):
            padded = ((n >> 7) + 1) << 7
            first = int((padded - n) / 2)
            return first, padded - n - first
        return 32, 32
    left, right = one(int(width))
    top, bottom = one(int(height))
    return left, right, top, bottom


class YUV420Reader(object):
    """Planar 8-bit I420 file -> RGB uint8 frames (reference `YUV_Read`, yuv_frame_io.py:32-103)."""

    def __init__(self, path, height, width, to_rgb=True):
        self.h, self.w = int(height), int(width)
        if self.h % 2 or self.w % 2:
            raise ValueError("YUV 4:2:0 needs even frame extents")
        self.fp = open(path, "rb")
        self.y_len = self.h * self.w
        self.c_len = self.y_len // 4
        self.frame_len = self.y_len + 2 * self.c_len          # int(1.5 * h * w), :42
        self.to_rgb = to_rgb

    def __len__(self):
        return os.fstat(self.fp.fileno()).st_size // self.frame_len

    def read(self, index=None):
        """(frame, True), or (None, False) past the end (:54-57).  index: frame number to seek to (:49-50)."""
        if index is not None:
            self.fp.seek(int(index) * self.frame_len, 0)
        buf = np.frombuffer(self.fp.read(self.frame_len), dtype=np.uint8)
        if buf.size < self.frame_len:
            return None, False
        y = buf[:self.y_len].reshape(self.h, self.w)                               # :59-60 (column-major + transpose)
        u = buf[self.y_len:self.y_len + self.c_len].reshape(self.h // 2, self.w // 2)
        v = buf[self.y_len + self.c_len:].reshape(self.h // 2, self.w // 2)
        u = np.repeat(np.repeat(u, 2, axis=0), 2, axis=1)                          # imresize(..., 'nearest'), :68-69
        v = np.repeat(np.repeat(v, 2, axis=0), 2, axis=1)
        if not self.to_rgb:
            return np.stack((y, u, v), axis=-1), True                              # :97-99
        yuv = np.stack((y / 255.0, u / 255.0 - 0.5, v / 255.0 - 0.5), axis=-1)     # :86-89
        rgb = (255.0 * np.clip(yuv2rgb(yuv), 0.0, 1.0)).astype("uint8")            # :90 (truncating cast)
        return rgb, True

    def close(self):
        self.fp.close()


class YUV420Writer(object):
    """RGB uint8 frames -> planar 8-bit I420 file (reference `YUV_Write`, yuv_frame_io.py:106-187)."""

    def __init__(self, path, from_rgb=True):
        self.fp = open(path, "wb")  # no appending (:116)
        self.from_rgb = from_rgb

    def write(self, frame):
        frame = np.asarray(frame)
        if frame.ndim != 3 or frame.shape[2] != 3:
            raise ValueError("expected an HxWx3 frame")
        if self.from_rgb:
            yuv = rgb2yuv(frame / 255.0)                                           # :131-132
            y = yuv[:, :, 0]
            u = np.clip(yuv[:, :, 1] + 0.5, 0.0, 1.0)[::2, ::2]                    # :140-144
            v = np.clip(yuv[:, :, 2] + 0.5, 0.0, 1.0)[::2, ::2]
            y, u, v = ((255.0 * p).astype("uint8") for p in (y, u, v))             # :145-147
        else:
            y, u, v = frame[:, :, 0], frame[::2, ::2, 1], frame[::2, ::2, 2]       # :149-152
        for plane in (y, u, v):
            np.ascontiguousarray(plane).tofile(self.fp)                            # :170-176
        return True

    def close(self):
        self.fp.close()


def frame_pairs(n_frames, step=2):
    """(first, second) indices the HD demo interpolates between (demo_HD720p.py:69-72: index, index + 2)."""
    return [(i, i + step) for i in range(0, int(n_frames) - step, step)]


def shard_pairs(pairs, rank, world):
    """Contiguous split of the pair list over `world` ranks (frames are independent units: SURVEY 8e)."""
    from .shard import frame_range
    lo, hi = frame_range(len(pairs), rank, world)
    return pairs[lo:hi]


def interpolate_pairs(model, read_frame, pairs, device, save_which=0):
    """For each (i, j) in `pairs`: read both frames (HxWx3 uint8), pad like the demos, run
    `model(torch.stack((X0, X1), 0))` under torch.no_grad() -- callers get the reference's inference
    behaviour (fill-hole on) exactly this way, SURVEY 3.1 -- crop, and yield (i, j, uint8 HxWx3).
    `model` may return the frame tensor itself or the reference networks' tuple
    (y_s, offset, filter, occlusion), of which y_s[save_which] is the frame (demo_HD720p.py:115-116)."""
    device = torch.device(device)
    for i, j in pairs:
        f0, ok0 = read_frame(i)
        f1, ok1 = read_frame(j)
        if not ok0 or not ok1:                                                     # demo_HD720p.py:73-74
            break
        h, w = f0.shape[:2]
        left, right, top, bottom = network_padding(h, w)
        pad = torch.nn.ReplicationPad2d([left, right, top, bottom])
        x0 = torch.from_numpy(np.transpose(f0, (2, 0, 1)).astype("float32") / 255.0).unsqueeze(0)   # :76-77
        x1 = torch.from_numpy(np.transpose(f1, (2, 0, 1)).astype("float32") / 255.0).unsqueeze(0)
        with torch.no_grad():
            y = model(torch.stack((pad(x0).to(device), pad(x1).to(device)), dim=0))
        if isinstance(y, (tuple, list)):
            y = y[0][save_which] if isinstance(y[0], (tuple, list)) else y[0]
        y = y.detach().float().cpu().numpy()
        y = np.transpose(255.0 * y.clip(0, 1.0)[0, :, top:top + h, left:left + w], (1, 2, 0))        # :136
        yield i, j, np.round(y).astype(np.uint8)                                                     # :150
