"""Drop-in replacement of baowenbo/MEMC-Net's `my_package` (reference: /my_package).

Same import paths and class names, so the reference's networks/MEMC_Net*.py import it
unchanged:

    from my_package.modules.FilterInterpolationModule import FilterInterpolationModule
    from my_package.modules.FlowProjectionModule import FlowProjectionModule
    from my_package.modules.InterpolationModule import InterpolationModule

Compute is libmemc_b200.so (hand-written sm_100a CUDA behind the C ABI of
include/memc_b200.h); there is no CPU fallback.
"""
