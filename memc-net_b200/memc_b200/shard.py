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


class PeerGather(object):
    """The optional output exchange over NVLink peer memory instead of a NCCL collective.

    Every rank owns one SYMMETRIC buffer [world, B, ...] (torch.distributed._symmetric_memory: the same allocation on every
    rank, each mapped into every other rank's address space over NVLink / NVSwitch).  `push(frame, index)` orders a side
    stream after the caller's current stream and copies the frame into slot [rank, index] of EVERY rank's buffer with
    device-to-device copies -- copy engines, no SMs, so the exchange of frame i runs under the compute of frame i + 1 without
    taking SMs from it the way a NCCL kernel does (measured on 8 x B200: profiles/r02_scaling.md).  `finish()` adds a
    device-side barrier over all ranks (signal pads in the same symmetric memory) behind the pushes and makes the caller's
    stream wait for it: after that the local buffer holds every rank's frames.

    One process per GPU, NCCL process group initialised (the rendezvous exchanges the memory handles through it).  Raises
    RuntimeError where symmetric memory is not available (callers fall back to `gather_frames`)."""

    def __init__(self, frames_per_rank_shape, dtype=torch.float32, device=None, group=None, streams=None):
        import torch.distributed._symmetric_memory as symm
        group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        try:
            symm.enable_symm_mem_for_group(group.group_name)
        except Exception:  # noqa: BLE001  (newer torch enables it on rendezvous)
            pass
        shape = (self.world,) + tuple(frames_per_rank_shape)
        self.buf = symm.empty(*shape, dtype=dtype, device=device)
        self.hdl = symm.rendezvous(self.buf, group)
        self.peers = [self.hdl.get_buffer(r, shape, dtype) for r in range(self.world)]
        # one copy stream per peer (capped): a single stream of peer copies keeps ONE copy engine busy (~380 GB/s
        # measured on 8 x B200); several streams drive several engines and NVLink ports at once
        n = max(1, min(self.world - 1, 7) if streams is None else int(streams))
        self.streams = [torch.cuda.Stream(device=device) for _ in range(n)]
        self.stream = self.streams[0]
        self._read_done = {}
        self.hdl.barrier()

    def push(self, frame, index):
        """frame: this rank's frame `index` ([...] = frames_per_rank_shape[1:]), produced on the current stream.  The source
        may be overwritten again once `wait_reusable(index)` has been ordered before the overwrite."""
        cur = torch.cuda.current_stream(self.buf.device)
        evs = self._read_done.setdefault(index, [torch.cuda.Event() for _ in self.streams])
        for s in self.streams:
            s.wait_stream(cur)
        for k in range(self.world):  # the own slot first, then the peers in ring order, round robin over the copy streams
            r = (self.rank + k) % self.world
            with torch.cuda.stream(self.streams[k % len(self.streams)]):
                self.peers[r][self.rank, index].copy_(frame, non_blocking=True)
        for s, ev in zip(self.streams, evs):
            ev.record(s)

    def wait_reusable(self, index):
        """Order the current stream after the last push of frame `index` has finished READING its source."""
        cur = torch.cuda.current_stream(self.buf.device)
        for ev in self._read_done.get(index, ()):
            cur.wait_event(ev)

    def barrier_async(self):
        """Device-side barrier over all ranks behind the pushes issued so far, on the first copy stream (nobody waits here)."""
        for s in self.streams[1:]:
            self.stream.wait_stream(s)
        with torch.cuda.stream(self.stream):
            self.hdl.barrier()

    def finish(self):
        """barrier_async() + make the caller's current stream wait: afterwards the local buffer holds everybody's frames."""
        self.barrier_async()
        torch.cuda.current_stream(self.buf.device).wait_stream(self.stream)
        return self.buf
