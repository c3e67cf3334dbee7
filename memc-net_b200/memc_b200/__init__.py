"""memc_b200 -- host-side plumbing of the B200-native MEMC-Net motion-compensation ops.

    memc_b200.lib       ctypes binding of libmemc_b200.so (the C ABI of include/memc_b200.h)
    memc_b200.shard     frame sharding across the GPUs of one box (torch.distributed / NCCL)
    memc_b200.compat    import shims so the reference's networks/ import on a modern stack
    memc_b200.fused     the networks' call sites around the ops as single calls (FilterInterpolate, FlowProjectPair)
    memc_b200.host_pipeline  pinned-host batches streamed through the GPU (bench.py's e2e path)
    memc_b200.video     YUV420 IO, the demos' padding rule, frame-pair sharding and driver (host side)

The drop-in package the reference's networks import is the sibling `my_package/`.
"""
from . import lib  # noqa: F401

__all__ = ["lib"]
__version__ = "0.1.0"
