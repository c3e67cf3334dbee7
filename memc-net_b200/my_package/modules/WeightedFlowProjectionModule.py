"""`WeightedFlowProjectionModule` -- nn.Module face of the brightness-gated flow forward-splat.

Same shape as the reference's FlowProjectionModule (my_package/modules/FlowProjectionModule.py:5-12) with the two frames as
further inputs: `WeightedFlowProjectionModule(flow.requires_grad, threshold)(flow, frame0, frame2)`.
"""
from torch import nn

from my_package.functions.WeightedFlowProjectionLayer import WeightedFlowProjectionLayer


class WeightedFlowProjectionModule(nn.Module):
    def __init__(self, requires_grad=True, threshold=2.0):
        super().__init__()
        self.f = WeightedFlowProjectionLayer(requires_grad, threshold)

    def forward(self, input1, input2, input3):
        return self.f(input1, input2, input3)

    def extra_repr(self):
        return "fillhole=%d, threshold=%g, libmemc_b200 (sm_100a)" % (self.f.fillhole, self.f.threshold)
