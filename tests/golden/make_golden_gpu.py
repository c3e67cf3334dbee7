"""GPU-side golden vectors: FlowProjection WITH fill-hole, produced by the reference's own
CUDA kernels (oracle/_ref/libmemc_ref_gpu.so = my_lib_kernel.cu recompiled for sm_100a),
because the reference has no CPU fill-hole (my_lib.c:1539-1543).

Run on the GPU box:  python tests/golden/make_golden_gpu.py gpurun_out/golden
then copy the .npz files into tests/golden/.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    H, W = 20, 28
    ys = torch.arange(H, dtype=torch.float32).view(1, 1, H, 1)
    xs = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W)
    g = torch.Generator().manual_seed(7)
    div = torch.cat([0.6 * (xs - (W - 1) / 2).expand(1, 1, H, W), 0.6 * (ys - (H - 1) / 2).expand(1, 1, H, W)], 1)
    div = div + 0.3 * torch.randn(1, 2, H, W, generator=g)
    rnd = 4.0 * torch.randn(1, 2, H, W, generator=g)
    for name, flow in (("fp_fillhole_divergent_gpu", div), ("fp_fillhole_random_gpu", rnd)):
        t = flow.contiguous().cuda()
        out, count = ref.gpu_flow_projection_forward(t, 1)
        torch.cuda.synchronize()
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), op="flow_projection", fillhole=1,
                            flow=t.cpu().numpy(), out=out.cpu().numpy(), count=count.cpu().numpy())
        print("wrote", name, "holes:", float((count == 0).float().mean()))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
