"""Shared plumbing of the re-authored autograd Functions.

The reference's Functions are PyTorch-0.2 legacy (instance-style) Functions
(my_package/functions/*.py); torch >= 1.3 refuses to run those.  Each op is therefore a
new-style static `torch.autograd.Function` (`_XxxFunction`) plus a thin class with the
REFERENCE'S NAME AND CONSTRUCTOR that is callable like the legacy instance was:

    FilterInterpolationLayer()(input1, input2, input3)        # reference style
    FilterInterpolationLayer.apply(input1, input2, input3)    # modern style
"""
from memc_b200 import lib as _lib


def fast_call(name, *args):
    return _lib.call(name, *args)


def prep(t, name):
    """The reference does `.contiguous()` on every input (FilterInterpolationLayer.py:14-16)."""
    return _lib.check_tensor(t, name).contiguous()
