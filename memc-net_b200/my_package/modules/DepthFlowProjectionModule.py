"""`DepthFlowProjectionModule` -- nn.Module face of the depth-weighted flow forward-splat.

Same shape as the reference's FlowProjectionModule (my_package/modules/FlowProjectionModule.py:5-12) with the weight map
as a second input: `DepthFlowProjectionModule(flow.requires_grad)(flow, inverse_depth)`.
"""
from torch import nn

from my_package.functions.DepthFlowProjectionLayer import DepthFlowProjectionLayer


class DepthFlowProjectionModule(nn.Module):
    def __init__(self, requires_grad=True):
        super().__init__()
        self.f = DepthFlowProjectionLayer(requires_grad)

    def forward(self, input1, input2):
        return self.f(input1, input2)

    def extra_repr(self):
        return "fillhole=%d, libmemc_b200 (sm_100a)" % self.f.fillhole
