"""`FilterInterpolationModule` -- nn.Module face of the adaptive-warp op.

Import path, class name, constructor and `forward(input1, input2, input3)` signature are the
contract with the reference's networks (`networks/MEMC_Net.py:258-264` builds a fresh
`FilterInterpolationModule()` per call; reference class: my_package/modules/
FilterInterpolationModule.py:7-13).  The attribute `f` holds the op object, as there.

    input1  [B, C, H, W]      frame or feature map to warp
    input2  [B, 2, H, W]      flow (x, y) from the target pixel to its source position
    input3  [B, fs*fs, H, W]  per-pixel fs x fs kernel (fs = 4 in the shipped models)
    ->      [B, C, H, W]
"""
from torch import nn

from my_package.functions.FilterInterpolationLayer import FilterInterpolationLayer


class FilterInterpolationModule(nn.Module):
    def __init__(self):
        super().__init__()
        self.f = FilterInterpolationLayer()

    def forward(self, input1, input2, input3):
        warped = self.f(input1, input2, input3)
        return warped

    def extra_repr(self):
        return "adaptive warp, libmemc_b200 (sm_100a)"
