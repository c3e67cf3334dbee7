"""Host-resident batches through the adaptive-warp op: frame-pipelined H2D -> fwd (-> bwd) -> D2H.

Frames are independent units of the path, so a batch that lives in pinned
HOST memory is streamed through the GPU one frame at a time on a small ring of CUDA streams:
while frame i computes, frame i+1 uploads and frame i-1 downloads (both PCIe directions and the
SMs busy at once).  This is the public API bench.py's `e2e` number goes through.

    pipe = FilterInterpolationHostPipeline(device, streams=3)
    out, (gi1, gi2, gi3) = pipe.forward_backward(h_in1, h_flow, h_filt, h_gout, outputs=...)

Back-to-back batches: `forward_backward(..., wait=False)` does not join the ring with the caller's
stream, so the first upload of the next batch overlaps the last download of this one (a batch of B
frames otherwise costs B + 1 transfer slots: nothing to download during the first upload, nothing
to upload during the last download); call `pipe.join()` before reading the host results.
"""
import os

import torch

from my_package.modules.FilterInterpolationModule import FilterInterpolationModule


def bind_to_gpu_numa_node(device_index):
    """Pin this process (and so its first-touch page placement, incl. the pinned staging buffers allocated
    afterwards) to the CPUs of the NUMA node the GPU hangs off.  One process per GPU is the deployment model
    (torchrun); without this every rank's pinned buffers can land on one socket and all H2D / D2H traffic of an
    8-GPU box crosses the inter-socket link (profiles/r02_scaling.md).  Returns a dict describing what was done;
    never raises (a container may hide /sys or forbid sched_setaffinity)."""
    info = {"device": int(device_index), "numa_node": None, "cpus": None, "bound": False}
    try:
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        bdf = "%04x:%02x:%02x.0" % (dom, bus, dev)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            node = int(f.read().strip())
        info["pci"] = bdf
        info["numa_node"] = node
        if node < 0:
            return info
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["cpus"] = len(allowed)
            info["bound"] = True
    except Exception as e:  # noqa: BLE001
        info["error"] = repr(e)[:120]
    return info


class FilterInterpolationHostPipeline(object):
    def __init__(self, device, streams=3):
        self.device = torch.device(device)
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(max(1, int(streams)))]
        self.module = FilterInterpolationModule()
        self._primed = False  # the ring has been ordered after the caller's stream at least once

    @staticmethod
    def alloc_outputs(h_in1, h_flow, h_filt):
        """Pinned host buffers for (output, gradinput1, gradinput2, gradinput3)."""
        return tuple(torch.empty_like(t, device="cpu").pin_memory() for t in (h_in1, h_in1, h_flow, h_filt))

    def join(self):
        """Make the caller's current stream wait for everything the ring has been given."""
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            cur.wait_stream(s)

    def forward_backward(self, h_in1, h_flow, h_filt, h_gout, outputs=None, wait=True):
        """All arguments are pinned CPU tensors [B, ...]; returns pinned CPU results.  Every frame's
        inputs cross H2D and every frame's output + three gradients cross D2H on every call.
        wait=False: the caller must not touch the host buffers of this batch until `join()` +
        a synchronisation of its stream (frames of consecutive batches keep their stream order)."""
        for t in (h_in1, h_flow, h_filt, h_gout):
            if t.is_cuda or not t.is_pinned():
                raise ValueError("host pipeline expects pinned CPU tensors")
        if outputs is None:
            outputs = self.alloc_outputs(h_in1, h_flow, h_filt)
        h_out, h_g1, h_g2, h_g3 = outputs
        B = h_in1.size(0)
        cur = torch.cuda.current_stream(self.device)
        if wait or not self._primed:
            for s in self.streams:
                s.wait_stream(cur)
            self._primed = True
        for b in range(B):
            s = self.streams[b % len(self.streams)]
            with torch.cuda.stream(s):
                a = h_in1[b:b + 1].to(self.device, non_blocking=True).requires_grad_()
                f = h_flow[b:b + 1].to(self.device, non_blocking=True).requires_grad_()
                k = h_filt[b:b + 1].to(self.device, non_blocking=True).requires_grad_()
                g = h_gout[b:b + 1].to(self.device, non_blocking=True)
                o = self.module(a, f, k)
                g1, g2, g3 = torch.autograd.grad(o, (a, f, k), g)
                h_out[b:b + 1].copy_(o.detach(), non_blocking=True)
                h_g1[b:b + 1].copy_(g1, non_blocking=True)
                h_g2[b:b + 1].copy_(g2, non_blocking=True)
                h_g3[b:b + 1].copy_(g3, non_blocking=True)
        if wait:
            self.join()
        return h_out, (h_g1, h_g2, h_g3)
