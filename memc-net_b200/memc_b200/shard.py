"""Frame sharding of the motion-compensation ops across the GPUs of one box.

Every kernel of the path indexes the batch by blockIdx.z and never crosses batch items
(reference my_lib_kernel.cu:1114-1115, 1658-1659), so frames (or frame pairs) are independent
units: rank r of G owns the contiguous slice [r*B/G, (r+1)*B/G) (uneven batches: the first
B % G ranks get one extra frame).  There is NO data-path collective; the only exchange is the
optional gather of the output batch (an all-gather over NCCL / NVLink on GPUs, gloo on CPU for
the host-logic tests).  One process per GPU (torchrun), backend chosen by the caller.
"""
import torch
import torch.distributed as dist


def frame_range(batch, rank, world):
    """[start, stop) of the frames owned by `rank` (contiguous split, remainder to low ranks)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world %r/%r" % (rank, world))
    base, rem = divmod(int(batch), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_frames(tensors, rank=None, world=None):
    """Slice every [B, ...] tensor in `tensors` down to this rank's frames (views, no copy)."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    single = isinstance(tensors, torch.Tensor)
    ts = (tensors,) if single else tuple(tensors)
    out = []
    for t in ts:
        lo, hi = frame_range(t.size(0), rank, world)
        out.append(t[lo:hi])
    return out[0] if single else tuple(out)


def gather_frames(local, batch, group=None):
    """All-gather the per-rank output slices back into the full [batch, ...] tensor, in rank
    order.  Uneven slices are padded to the largest one for the collective and trimmed after."""
    world = dist.get_world_size(group)
    sizes = [frame_range(batch, r, world)[1] - frame_range(batch, r, world)[0] for r in range(world)]
    mx = max(sizes)
    pad = local
    if local.size(0) < mx:
        pad = torch.cat([local, local.new_zeros((mx - local.size(0),) + tuple(local.shape[1:]))], 0)
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad.contiguous(), group=group)
    return torch.cat([b[:n] for b, n in zip(bufs, sizes)], 0)


def run_sharded(fn, tensors, gather=True, group=None):
    """Apply `fn(*local_tensors) -> Tensor` to this rank's frames; optionally gather the result."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    batch = tensors[0].size(0)
    local = shard_frames(tensors, rank, world)
    out = fn(*local)
    return gather_frames(out, batch, group) if gather else out
