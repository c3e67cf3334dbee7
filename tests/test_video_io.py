"""Host-side IO / padding / driver around the hot path (memc_b200.video).  CPU only.
The colour matrices and the nearest chroma upsampling restate third-party code the reference calls
(scikit-image 0.13 / SciPy 1.1, absent here): those are checked for self-consistency and against the
published BT.601-analog values, not against the originals (parity unpinned, see the module header)."""
import numpy as np
import pytest
import torch

from memc_b200 import video


@pytest.mark.parametrize("h,w,exp", [
    (720, 1280, (32, 32, 24, 24)),     # demo_HD720p.py: 1280 is a multiple of 128 -> +32+32 = 1344; 720 -> 768
    (480, 640, (32, 32, 16, 16)),      # 640 = 5 * 128 -> 704; 480 -> 512
    (1080, 1920, (32, 32, 36, 36)),    # 1920 = 15 * 128 -> 1984; 1080 -> 1152 (1984 x 1152)
    (256, 448, (32, 32, 32, 32)),      # the Vimeo fixtures: 448 -> 512, 256 is a multiple -> 320
    (100, 129, (63, 64, 14, 14)),      # odd remainder: the extra pixel goes right / bottom
    (1, 1, (63, 64, 63, 64)),
])
def test_network_padding_matches_the_demo_rule(h, w, exp):
    left, right, top, bottom = video.network_padding(h, w)
    assert (left, right, top, bottom) == exp
    assert (w + left + right) % 32 == 0 and (h + top + bottom) % 32 == 0
    if w % 128:
        assert (w + left + right) % 128 == 0
    if h % 128:
        assert (h + top + bottom) % 128 == 0


def test_colour_matrices():
    # luma row is BT.601; white -> (1, 0, 0); the inverse really inverts
    assert np.allclose(video.YUV_FROM_RGB[0], [0.299, 0.587, 0.114])
    assert np.allclose(video.rgb2yuv(np.ones((1, 1, 3))), [[[1.0, 0.0, 0.0]]], atol=1e-7)
    rgb = np.random.default_rng(0).random((5, 7, 3))
    assert np.abs(video.yuv2rgb(video.rgb2yuv(rgb)) - rgb).max() < 1e-12


def test_yuv420_write_read_round_trip(tmp_path):
    """Writer then Reader: luma survives to 1 level (two truncating casts), frames with constant chroma per 2x2
    block survive entirely; frame count / seek / end of file behave like the reference reader."""
    h, w, n = 16, 24, 3
    rng = np.random.default_rng(1)
    frames = []
    for _ in range(n):
        blocks = rng.integers(64, 192, size=(h // 2, w // 2, 3), dtype=np.uint8)      # unsaturated: the writer clips U, V to [0, 1]
        frames.append(np.repeat(np.repeat(blocks, 2, axis=0), 2, axis=1))         # chroma constant per 2x2 block
    path = str(tmp_path / "clip.yuv")
    wr = video.YUV420Writer(path)
    for f in frames:
        assert wr.write(f)
    wr.close()
    import os
    assert os.path.getsize(path) == n * (h * w * 3 // 2)
    rd = video.YUV420Reader(path, h, w)
    assert len(rd) == n
    for k in (2, 0, 1):                                                             # random access by frame index
        got, ok = rd.read(k)
        assert ok and got.shape == (h, w, 3) and got.dtype == np.uint8
        assert np.abs(got.astype(int) - frames[k].astype(int)).max() <= 6           # truncating 8-bit casts on the way in and out (chroma gain up to 2.03)
    got, ok = rd.read(n)
    assert got is None and not ok
    yuv, ok = video.YUV420Reader(path, h, w, to_rgb=False).read(0)
    assert ok and yuv.shape == (h, w, 3) and np.array_equal(yuv[0::2, 0::2, 1], yuv[1::2, 1::2, 1])   # nearest 2x chroma
    rd.close()


def test_frame_pairs_and_sharding():
    pairs = video.frame_pairs(100, 2)
    assert pairs[0] == (0, 2) and pairs[-1] == (96, 98) and len(pairs) == 49      # demo_HD720p.py:69-72
    got = []
    for r in range(8):
        got += video.shard_pairs(pairs, r, 8)
    assert got == pairs                                                            # every pair exactly once, in order
    sizes = [len(video.shard_pairs(pairs, r, 8)) for r in range(8)]
    assert max(sizes) - min(sizes) <= 1


def test_interpolate_pairs_pads_runs_and_crops():
    """Driver logic with a stand-in model (mean of the two frames): the model sees the padded size the demos use, the
    result is cropped back and rounded; a tuple-returning model (the reference networks' convention) works too."""
    h, w = 70, 200
    rng = np.random.default_rng(2)
    clip = [rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8) for _ in range(5)]
    seen = []

    def model(x):                          # x: [2, 1, 3, Hp, Wp]
        seen.append(tuple(x.shape))
        assert not torch.is_grad_enabled()
        return 0.5 * (x[0] + x[1])

    def read(i):
        return (clip[i], True) if i < len(clip) else (None, False)

    out = list(video.interpolate_pairs(model, read, video.frame_pairs(len(clip), 2) + [(4, 6)], "cpu"))
    assert [(i, j) for i, j, _ in out] == [(0, 2), (2, 4)]                          # (4, 6) runs off the clip: stops
    assert seen[0] == (2, 1, 3, 128, 256)
    for i, j, mid in out:
        exp = np.round(0.5 * (clip[i].astype(np.float32) / 255.0 + clip[j].astype(np.float32) / 255.0) * 255.0)
        assert mid.shape == (h, w, 3) and np.abs(mid.astype(int) - exp.astype(int)).max() <= 1
    ref_style = lambda x: ([0.5 * (x[0] + x[1]), x[0]], None, None, None)           # (y_s, offset, filter, occlusion)
    (_, _, mid2), = list(video.interpolate_pairs(ref_style, read, [(0, 2)], "cpu", save_which=0))
    assert np.array_equal(mid2, out[0][2])
