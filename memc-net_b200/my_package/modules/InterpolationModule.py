"""`InterpolationModule` -- nn.Module face of the plain bilinear backward warp.

Contract with the reference (`networks/MEMC_Net_VE.py:494-504`; reference class:
my_package/modules/InterpolationModule.py:5-11): no constructor arguments,
`forward(input1 [B,C,H,W], input2 = flow [B,2,H,W]) -> [B,C,H,W]`, zero outside the frame.
"""
from torch import nn

from my_package.functions.InterpolationLayer import InterpolationLayer


class InterpolationModule(nn.Module):
    def __init__(self):
        super().__init__()
        self.f = InterpolationLayer()

    def forward(self, input1, input2):
        warped = self.f(input1, input2)
        return warped

    def extra_repr(self):
        return "bilinear warp, libmemc_b200 (sm_100a)"
