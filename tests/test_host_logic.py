"""Host-side mirror of the reference interface, checked on CPU: import paths / class names /
constructor signatures the reference's networks rely on; the -1 error convention of the FFI
namespace (my_lib_cuda.c); loud failure (no CPU fallback) for CPU tensors."""
import inspect

import pytest
import torch


def test_reference_import_paths_and_signatures(built_lib):
    from my_package.modules.FilterInterpolationModule import FilterInterpolationModule
    from my_package.modules.FlowProjectionModule import FlowProjectionModule
    from my_package.modules.InterpolationModule import InterpolationModule
    from my_package.functions.FilterInterpolationLayer import FilterInterpolationLayer
    from my_package.functions.FlowProjectionLayer import FlowProjectionLayer
    from my_package.functions.InterpolationLayer import InterpolationLayer
    from my_package.functions.SeparableConvLayer import SeparableConvLayer
    assert list(inspect.signature(FilterInterpolationModule.forward).parameters) == ["self", "input1", "input2", "input3"]
    assert list(inspect.signature(FlowProjectionModule.__init__).parameters) == ["self", "requires_grad"]
    assert inspect.signature(FlowProjectionModule.__init__).parameters["requires_grad"].default is True
    assert list(inspect.signature(FlowProjectionModule.forward).parameters) == ["self", "input1"]
    assert list(inspect.signature(InterpolationModule.forward).parameters) == ["self", "input1", "input2"]
    assert isinstance(FilterInterpolationModule().f, FilterInterpolationLayer)
    assert isinstance(InterpolationModule().f, InterpolationLayer)
    # fillhole = 1 iff the input did not require grad (FlowProjectionLayer.py:15)
    assert FlowProjectionLayer(False).fillhole == 1 and FlowProjectionLayer(True).fillhole == 0
    assert FlowProjectionModule().f.fillhole == 0
    assert SeparableConvLayer(4).filtersize == 4


def test_ffi_namespace_names(built_lib):
    import my_package._ext.my_lib as my_lib
    for op in ["FilterInterpolationLayer", "FlowProjectionLayer", "InterpolationLayer", "InterpolationChLayer",
               "SeparableConvLayer"]:
        assert callable(getattr(my_lib, op + "_gpu_forward"))
        assert callable(getattr(my_lib, op + "_gpu_backward"))


def test_ffi_returns_minus_one_on_shape_or_stride_violation(built_lib):
    import my_package._ext.my_lib as my_lib
    z = torch.zeros
    # flow with 3 channels (my_lib_cuda.c:612)
    assert my_lib.FilterInterpolationLayer_gpu_forward(z(1, 3, 8, 8), z(1, 3, 8, 8), z(1, 16, 8, 8), z(1, 3, 8, 8)) == -1
    # flow spatial mismatch (:616-617)
    assert my_lib.FilterInterpolationLayer_gpu_forward(z(1, 3, 8, 8), z(1, 2, 8, 9), z(1, 16, 8, 8), z(1, 3, 8, 8)) == -1
    # w-stride != 1 (:642)
    assert my_lib.FilterInterpolationLayer_gpu_forward(z(1, 3, 8, 16)[..., ::2], z(1, 2, 8, 8), z(1, 16, 8, 8), z(1, 3, 8, 8)) == -1
    # output batch stride differs (:645)
    assert my_lib.FilterInterpolationLayer_gpu_forward(z(1, 3, 8, 8), z(1, 2, 8, 8), z(1, 16, 8, 8), z(2, 4, 8, 8)[:1, :3]) == -1
    # FlowProjection needs 2 channels (:763)
    assert my_lib.FlowProjectionLayer_gpu_forward(z(1, 3, 8, 8), z(1, 1, 8, 8), z(1, 3, 8, 8), 0) == -1
    assert my_lib.FlowProjectionLayer_gpu_backward(z(1, 2, 8, 8), z(1, 2, 8, 8), z(1, 2, 8, 8), z(1, 2, 8, 8)) == -1
    # Interpolation (non-Ch) needs 3 channels (:373)
    assert my_lib.InterpolationLayer_gpu_forward(z(1, 4, 8, 8), z(1, 2, 8, 8), z(1, 4, 8, 8)) == -1
    # SeparableConv filter extent (:218-219)
    assert my_lib.SeparableConvLayer_gpu_forward(z(1, 3, 8, 8), z(1, 4, 6, 5), z(1, 4, 5, 5), z(1, 3, 5, 5)) == -1


def test_no_cpu_fallback(built_lib):
    """Well-formed CPU tensors must fail LOUDLY: there is no CPU path in the product."""
    import my_package._ext.my_lib as my_lib
    from memc_b200.lib import MemcB200Error
    from my_package.modules.FilterInterpolationModule import FilterInterpolationModule
    from my_package.modules.FlowProjectionModule import FlowProjectionModule
    z = torch.zeros
    with pytest.raises(MemcB200Error):
        my_lib.FilterInterpolationLayer_gpu_forward(z(1, 3, 8, 8), z(1, 2, 8, 8), z(1, 16, 8, 8), z(1, 3, 8, 8))
    with pytest.raises(MemcB200Error):
        FilterInterpolationModule()(z(1, 3, 8, 8), z(1, 2, 8, 8), z(1, 16, 8, 8))
    with pytest.raises(MemcB200Error):
        FlowProjectionModule(False)(z(1, 2, 8, 8))
    from memc_b200 import fused
    with pytest.raises(MemcB200Error):  # the fused call site has no CPU path either
        fused.FilterInterpolate(z(1, 3, 8, 8), z(1, 3, 8, 8), [z(1, 2, 8, 8)] * 2, [z(1, 16, 8, 8)] * 2, [z(1, 1, 8, 8)] * 2, 16)


def test_tools_do_not_touch_the_oracle():
    """Only tests/, smoke() and bench.py's CPU legs may execute anything under oracle/: the development
    tools must not (the legacy-kernel comparison lives in tests/legacy_bench.py)."""
    import os
    from tests.conftest import ROOT
    tdir = os.path.join(ROOT, "tools")
    for f in os.listdir(tdir):
        if f.endswith((".py", ".sh")):
            src = open(os.path.join(tdir, f)).read()
            assert "import oracle" not in src and "from oracle" not in src and "oracle/_ref" not in src, f


def test_missing_library_fails_loudly(monkeypatch, built_lib):
    from memc_b200 import lib
    monkeypatch.setattr(lib, "_lib", None)
    monkeypatch.setattr(lib, "LIB_PATH", "/nonexistent/libmemc_b200.so")
    with pytest.raises(lib.MemcB200Error):
        lib.load()


def test_product_does_not_import_oracle():
    """The product tree must never reference the oracle (test infrastructure)."""
    import os
    from tests.conftest import PKG
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, os.path.join(dirpath, f)
                assert "liboracle" not in src and "libmemc_ref" not in src, os.path.join(dirpath, f)


def test_fi_accumulation_mode_is_exposed(built_lib):
    """training code can choose fp32 accumulation of gradinput1 (run time or MEMC_B200_FI_ACCUM)."""
    from memc_b200 import lib
    assert lib.get_fi_accumulation() in ("fixed", "float")
    old = lib.get_fi_accumulation()
    try:
        lib.set_fi_accumulation("float")
        assert lib.fi_backward_flags() == lib.OVERWRITE | lib.FLOAT_ACCUM
        lib.set_fi_accumulation("fixed")
        assert lib.fi_backward_flags() == lib.OVERWRITE
        with pytest.raises(ValueError):
            lib.set_fi_accumulation("double")
    finally:
        lib.set_fi_accumulation(old)


def test_operands_must_share_a_device():
    from memc_b200 import lib
    a, b = torch.zeros(1, device="cpu"), torch.zeros(1, device="meta")
    with pytest.raises(lib.MemcB200Error):
        lib.check_same_device(a, b)
    lib.check_same_device(a, a)
