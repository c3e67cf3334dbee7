"""Frame sharding (memc_b200.shard) on CPU with the gloo backend, world_size 2: the host-side
logic of the N>1 path (the CUDA ops themselves are covered by the gpu tests)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from memc_b200 import shard


def test_frame_range_partitions_exactly():
    for batch in (0, 1, 4, 7, 16, 33, 64):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = shard.frame_range(batch, r, world)
                assert 0 <= lo <= hi <= batch
                seen.extend(range(lo, hi))
            assert seen == list(range(batch))
            sizes = [shard.frame_range(batch, r, world)[1] - shard.frame_range(batch, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.frame_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, batch, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        full = torch.rand(batch, 3, 5, 7, generator=g)      # identical on every rank (same seed)
        flow = torch.rand(batch, 2, 5, 7, generator=g)
        a, f = shard.shard_frames((full, flow))
        lo, hi = shard.frame_range(batch, rank, world)
        assert a.shape[0] == hi - lo and torch.equal(a, full[lo:hi]) and torch.equal(f, flow[lo:hi])
        # a stand-in per-frame op (frames independent): result must equal the unsharded op
        op = lambda x, y: x * 2 + y.sum(1, keepdim=True)
        got = shard.run_sharded(op, (full, flow), gather=True)
        assert torch.equal(got, op(full, flow))
        part = shard.run_sharded(op, (full, flow), gather=False)
        assert torch.equal(part, op(full, flow)[lo:hi])
        ret[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch", [4, 5])
def test_shard_and_gather_world2(batch):
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, batch, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert dict(ret) == {0: 1, 1: 1}


def _video_worker(rank, world, port, path, h, w, n, ret):
    """Each rank interpolates ITS share of the frame pairs of one YUV clip (memc_b200.video); the mid frames
    are gathered as tensors and rank 0 checks them against the single-process run."""
    import numpy as np
    from memc_b200 import video
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = lambda x: 0.25 * x[0] + 0.75 * x[1]               # stand-in for the network (frames independent)
        rd = video.YUV420Reader(path, h, w)
        pairs = video.frame_pairs(n, 2)
        mine = video.shard_pairs(pairs, rank, world)
        mids = [m for _, _, m in video.interpolate_pairs(model, rd.read, mine, "cpu")]
        local = torch.from_numpy(np.stack(mids)) if mids else torch.empty(0, h, w, 3, dtype=torch.uint8)
        full = shard.gather_frames(local, len(pairs))
        if rank == 0:
            ref = [m for _, _, m in video.interpolate_pairs(model, rd.read, pairs, "cpu")]
            assert torch.equal(full, torch.from_numpy(np.stack(ref)))
        rd.close()
        ret[rank] = 1
    finally:
        dist.destroy_process_group()


def test_sharded_video_driver_world2(tmp_path):
    import numpy as np
    from memc_b200 import video
    h, w, n = 32, 48, 9
    rng = np.random.default_rng(3)
    path = str(tmp_path / "clip.yuv")
    wr = video.YUV420Writer(path)
    for _ in range(n):
        wr.write(rng.integers(64, 192, size=(h, w, 3), dtype=np.uint8))
    wr.close()
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_video_worker, args=(r, world, port, path, h, w, n, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert dict(ret) == {0: 1, 1: 1}


def test_compat_shim_makes_reference_networks_importable(built_lib):
    """With the shim, the reference's own networks/MEMC_Net*.py import OUR my_package unchanged
    (skipped where the reference tree is absent, e.g. on the GPU box)."""
    ref = "/root/reference"
    if not os.path.isdir(os.path.join(ref, "networks")):
        pytest.skip("reference tree not present")
    import subprocess
    import sys
    from tests.conftest import PKG
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from memc_b200 import compat; compat.install(%r)\n"
        "import networks, my_package\n"
        "from networks.MEMC_Net import MEMC_Net\n"
        "import my_package.modules.FilterInterpolationModule as m\n"
        "assert my_package.__file__.startswith(%r), my_package.__file__\n"
        "net = MEMC_Net(training=False)\n"
        "print('ok', sum(p.numel() for p in net.parameters()))\n" % (PKG, ref, PKG))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp", timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip().startswith("ok")
