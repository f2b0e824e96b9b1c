"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol
include/lego_klt.h declares; struct layouts match; without a GPU the compute calls fail loudly
(there is no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from lego_slam_b200 import _lib, build
    build.build()
    return _lib.load()


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "lego_klt.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lego_klt_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    from lego_slam_b200 import _lib
    assert _declared_functions() == sorted(_lib.EXPORTS)


def test_library_exports_every_declared_symbol(lib):
    for name in _declared_functions():
        assert hasattr(lib, name), f"liblego_klt.so does not export {name}"


def test_abi_version_and_defaults(lib):
    from lego_slam_b200 import _lib
    assert lib.lego_klt_abi_version() == 1
    p = _lib.Params()
    lib.lego_klt_default_params(C.byref(p))
    # the reference's literals: src/algorithm.cpp:40-42,113,135 ; call sites frontend_g2o.cpp:473,515
    assert (p.levels, p.patch_lo, p.patch_hi, p.max_iters, p.inverse, p.has_initial) == (4, -3, 3, 10, 0, 1)
    assert p.eps == 1e-2 and p.kernel == _lib.KERNEL_AUTO


def test_struct_layouts_match_the_header():
    from lego_slam_b200 import _lib
    from oracle import binding
    assert C.sizeof(_lib.Params) == 40 == C.sizeof(binding.Params)
    assert C.sizeof(_lib.Stats) == 4 * 8 + 8 * 8 + 2 * 8 + 4 * 8 + 4 * 4 == C.sizeof(binding.Stats)


def test_no_silent_cpu_fallback_without_a_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from lego_slam_b200 import _lib
    import lego_slam_b200 as klt
    assert lib.lego_klt_device_count() == -3  # LEGO_KLT_ERR_NO_DEVICE
    with pytest.raises(_lib.KltError) as ei:
        klt.Tracker(0)
    assert ei.value.code == -3


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "lego_slam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle/" not in src.replace("klt_oracle", "") or "../oracle" not in src, f
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert '#include "../../oracle' not in src and "klt_oracle.h" not in src, f
