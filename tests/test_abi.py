"""The C-ABI boundary, checked without a GPU: libmemc_b200.so builds (nvcc cross-compiles),
loads, and exports every symbol include/memc_b200.h declares; the ctypes table in
memc_b200/lib.py covers the same set; no compute is invoked here."""
import ctypes
import os
import re

from tests.conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "memc_b200.h")).read()
    return sorted(set(re.findall(r"MEMC_B200_API\s+[\w\s\*]+?\b(\w+)\s*\(", src)))


def test_header_declares_the_reference_launchers():
    names = _declared()
    for n in ["FilterInterpolationLayer_gpu_forward_kernel", "FilterInterpolationLayer_gpu_backward_kernel",
              "FlowProjection_gpu_forward_kernel", "FlowProjection_gpu_backward_kernel",
              "InterpolationLayer_gpu_forward_kernel", "InterpolationLayer_gpu_backward_kernel",
              "InterpolationChLayer_gpu_forward_kernel", "InterpolationChLayer_gpu_backward_kernel",
              "SeparableConvLayer_gpu_forward_kernel", "SeparableConvLayer_gpu_backward_kernel"]:
        assert n in names
    assert len(names) == 52


def test_library_exports_every_declared_symbol(built_lib):
    for n in _declared():
        assert hasattr(built_lib, n), n
        assert ctypes.cast(getattr(built_lib, n), ctypes.c_void_p).value


def test_ctypes_table_matches_header(built_lib):
    from memc_b200 import lib
    assert lib.EXPORTS == _declared()
    assert built_lib.memc_b200_abi_version() == 1
    assert "sm_100a" in lib.build_info()


def test_library_embeds_sm100a_code_only():
    """cuobjdump must list sm_100a cubins and nothing else (no multi-arch fallback)."""
    import shutil
    import subprocess
    from memc_b200 import lib
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        import pytest
        pytest.skip("cuobjdump not available")
    out = subprocess.run([exe, "-lelf", lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_reference_contract_strides_struct():
    from memc_b200 import lib
    assert ctypes.sizeof(lib.Strides) == 24
