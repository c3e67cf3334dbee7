"""`FlowProjectionModule` -- nn.Module face of the flow forward-splat.

Contract with the reference's networks (`networks/MEMC_Net.py:252-256`:
`FlowProjectionModule(input.requires_grad)(input)`; reference class: my_package/modules/
FlowProjectionModule.py:5-12): constructor argument `requires_grad` (default True) is handed to
the op, which fills holes only when it is False.

    input1  [B, 2, H, W]  flow frame0 -> frame2
    ->      [B, 2, H, W]  flow seen from the intermediate frame (mean of the splatted -flow)
"""
from torch import nn

from my_package.functions.FlowProjectionLayer import FlowProjectionLayer


class FlowProjectionModule(nn.Module):
    def __init__(self, requires_grad=True):
        super().__init__()
        self.f = FlowProjectionLayer(requires_grad)

    def forward(self, input1):
        projected = self.f(input1)
        return projected

    def extra_repr(self):
        return "fillhole=%d, libmemc_b200 (sm_100a)" % self.f.fillhole
